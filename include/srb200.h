/*
 * srb200.h -- C-ABI of libsrb200.so, the B200 (sm_100a) engine behind the conv / deconv /
 * PReLU / PixelShuffle stacks of the reference's base_networks.py.
 *
 * The reference is pure Python: its "FFI" for this path is the implicit ATen dispatch of
 *   torch.nn.Conv2d            base_networks.py:42,112-113,156   (cudnn_convolution fwd/bwd)
 *   torch.nn.ConvTranspose2d   base_networks.py:77, fsrcnn.py:33 (cudnn_convolution_transpose)
 *   torch.nn.PixelShuffle      base_networks.py:157              (pixel_shuffle / pixel_unshuffle)
 *   ReLU/PReLU/LeakyReLU       base_networks.py:51-56, fsrcnn.py:26
 *   torch.add (skip)           base_networks.py:149, vdsr.py:31, edsr.py:42, srgan.py:39
 * Each entry point below names the reference call it replaces.  All functions are plain C:
 * raw device pointers, sizes, strides; no torch types.  Return 0 on success, a negative
 * srb_status otherwise (message via srb_last_error(), thread local).  Nothing here ever falls
 * back to a CPU or library (cuDNN/cuBLAS) path: unsupported == error.
 *
 * Tensors are fp32, logical NCHW with explicit element strides, so NCHW-contiguous and
 * channels_last (NHWC) memory are both accepted.  The tensor-core kernels (math = SRB_MATH_TF32)
 * require channels_last activations (sc == 1) with C % 4 == 0 and C >= 8 (16-byte TMA pixel rows; wgrad: C % 32 == 0),
 * or C <= 4 (packed to NHWC4 internally); everything else runs on the
 * generic fp32 CUDA-core kernels of the same library.
 *
 * Threading: re-entrant, no global mutable state besides a thread-local error string.  All work
 * is enqueued on the given stream; no cudaMalloc / device synchronisation on any hot entry point
 * (CUDA-graph capturable).  Memory is owned by the caller.
 */
#ifndef SRB200_H_
#define SRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRB200_VERSION 200

typedef enum srb_status {
  SRB_OK = 0,
  SRB_EINVAL = -1,       /* bad argument / shape */
  SRB_EUNSUPPORTED = -2, /* valid request this build cannot run (never a silent fallback) */
  SRB_ECUDA = -3,        /* CUDA runtime / driver error */
  SRB_EWORKSPACE = -4    /* workspace too small */
} srb_status;

typedef enum srb_act { SRB_ACT_NONE = 0, SRB_ACT_RELU = 1, SRB_ACT_PRELU = 2, SRB_ACT_LRELU = 3 } srb_act;

typedef enum srb_math {
  SRB_MATH_FP32 = 0, /* CUDA-core fp32 FMA, fp32 accumulate (exact-order independent reference quality) */
  SRB_MATH_TF32 = 1, /* tcgen05 kind::tf32, operands RN-rounded to tf32, fp32 accumulate in TMEM */
  SRB_MATH_AUTO = 2, /* TF32 tensor path when the layer qualifies, FP32 otherwise */
  SRB_MATH_BF16 = 4, /* bf16 STORAGE: activations and activation gradients are bf16 NHWC tensors, operands feed
                        tcgen05 kind::f16 (bf16 x bf16 -> fp32 in TMEM), parameters / parameter gradients / biases stay fp32
                        (weights are converted while being packed).  3-channel network edges stay fp32 and run TF32.
                        BASELINE cfg4 (EDSR 256x32 "bf16"); parity contract: the reference under torch.autocast(bfloat16). */
  SRB_MATH_EXACT = 3 /* fp32-accurate on the tensor cores: operands split hi+lo into tf32 pairs, three partial products
                        (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo) accumulated in fp32 ("3xTF32"); activations stay full fp32.
                        Layers the tensor path cannot run fall to the FP32 CUDA-core kernels.  For the 1e-3 end-to-end
                        contract at the BASELINE depths (VDSR-20, EDSR-256x32), where single-pass TF32 reaches ~1.5e-3. */
} srb_math;

/* SRB_U8: image bytes; accepted only where a function says so (the target of srb_conv_fprop_loss) */
typedef enum srb_dtype { SRB_F32 = 0, SRB_BF16 = 1, SRB_U8 = 2 } srb_dtype;

/* Logical NCHW view with element strides (like torch.Tensor.stride()); dtype = srb_dtype of the elements. */
typedef struct srb_tensor4 {
  void *data;
  int64_t sn, sc, sh, sw;
  int32_t dtype;
} srb_tensor4;

/*
 * One convolution layer = torch.nn.Conv2d(Cin, Cout*ps*ps, (kh,kw), stride, pad) [transposed == 0]
 *                      or torch.nn.ConvTranspose2d(Cin, Cout, (kh,kw), stride, pad, out_pad) [transposed == 1]
 * followed (fused) by bias, activation, residual add and PixelShuffle(ps):
 *     z = conv(x) + bias;  a = act(z);  y = PixelShuffle_ps(a) + residual
 * (activation is elementwise with one scalar slope, so it commutes with the shuffle; this is
 *  PSBlock.forward base_networks.py:177-185, ConvBlock.forward :62-71, ResnetBlock.forward :134-150).
 * N,H,W: input batch / height / width.  Cout: channels of y AFTER the shuffle (conv emits Cout*ps*ps).
 */
typedef struct srb_conv_params {
  int32_t N, Cin, H, W;
  int32_t Cout, kh, kw;
  int32_t stride, pad, out_pad;
  int32_t transposed;
  int32_t ps;   /* PixelShuffle factor r, 1 = none */
  int32_t act;  /* srb_act */
  float slope;  /* LeakyReLU negative slope (0.2 in base_networks.py:56); ignored otherwise */
  int32_t math; /* srb_math */
} srb_conv_params;

int srb_version(void);
const char *srb_last_error(void);

/* Output spatial size of the conv itself (before PixelShuffle).  Returns SRB_EINVAL on bad params. */
int srb_conv_out_hw(const srb_conv_params *p, int32_t *Ho, int32_t *Wo);

/* Which path a call with these params and these layouts will take: 1 = tcgen05 tensor path,
 * 0 = fp32 CUDA-core path.  pass: 0 fprop, 1 dgrad, 2 wgrad.  x_cl / y_cl: channels_last flags. */
int srb_conv_uses_tensor_path(const srb_conv_params *p, int pass, int x_cl, int y_cl);

/* PixelShuffle layers (p->ps > 1): 1 when srb_conv_dgrad and srb_conv_wgrad take dz in y's (shuffled) layout and fold the
 * pixel-un-shuffle into their TMA load addressing (channels_last x and dz, Cout a multiple of the 128-byte channel block:
 * 32 fp32 / 64 bf16 channels); 0 when the caller has to run srb_pixel_unshuffle first and pass ps = 1 geometry. */
int srb_conv_backward_folds_ps(const srb_conv_params *p, int x_cl, int dz_cl);

/* Diagnostic: human-readable tile plan the tensor path would use for this layer (host only, no GPU work). */
int srb_conv_describe_plan(const srb_conv_params *p, int pass, char *buf, size_t n);

/* Bytes of scratch the call needs (may be 0).  pass as above. */
size_t srb_conv_workspace_bytes(const srb_conv_params *p, int pass);

/*
 * Forward.  Replaces Conv2d/ConvTranspose2d forward + bias + act + PixelShuffle + torch.add.
 *   x        (N,Cin,H,W)
 *   w        Conv2d: (Cout*ps*ps, Cin, kh, kw) contiguous; ConvTranspose2d: (Cin, Cout, kh, kw) contiguous
 *   bias     (Cout*ps*ps) or NULL
 *   alpha    device pointer to the single PReLU slope (nn.PReLU() default, base_networks.py:54) or NULL
 *   residual same shape as y, or NULL
 *   y        (N,Cout,Ho*ps,Wo*ps)
 *   preact   optional, same shape as y: receives PixelShuffle(z) (needed by PReLU backward), or NULL
 *   relu_bits optional (NULL = none): receives the packed sign pattern of z, 16 channels per uint16 --
 *            word ((n*Ho+oy)*Wo+ox)*(Cout/16) + c/16, bit c%16 set iff z[n,c,oy,ox] > 0 -- i.e. the ReLU mask of y in
 *            1/32 of y's bytes.  Tensor path only (srb_conv_uses_tensor_path), ps == 1, Cout % 16 == 0; else SRB_EUNSUPPORTED.
 */
int srb_conv_fprop(const srb_conv_params *p, const srb_tensor4 *x, const float *w, const float *bias,
                   const float *alpha, const srb_tensor4 *residual, const srb_tensor4 *y,
                   const srb_tensor4 *preact, uint16_t *relu_bits, void *ws, size_t ws_bytes, void *stream);

/*
 * Forward of the network's LAST convolution fused with the regression loss that follows it in the training loop
 * (`loss = MSELoss(model(x), target)`, espcn.py:128-129, srcnn.py:128-129; L1Loss in edsr.py:152-153):
 *     y = PixelShuffle_ps(conv(x) + bias);   *loss = mean((y-t)^2) [kind 0] or mean(|y-t|) [kind 1];
 *     dz = d loss / d y  (2 (y-t) / n or sign(y-t) / n), written by the same epilogue
 * so the loss forward, the loss backward and (for PixelShuffle layers) the pixel-un-shuffle of the gradient never run as
 * kernels.  y may be NULL (the loss is then the only product).  dz_unshuffled != 0 (requires ps == 4, NCHW y/target):
 * dz is the conv's own dense NHWC tensor (N, Cout*ps*ps, Ho, Wo), ready for srb_conv_dgrad/wgrad with ps = 1 geometry;
 * dz_unshuffled == 0 (requires ps == 1): dz has y's shape and strides of its own.  The loss sum is deterministic
 * (per-warp partials folded in a fixed order).  Layers with an activation or a residual are SRB_EUNSUPPORTED.
 * target->dtype may be SRB_U8: the decoded image itself (strides in bytes; for an (N,H,W,C) HWC buffer sn = H*W*C, sc = 1,
 * sh = W*C, sw = C), read as t = byte / 255 (correctly rounded: bit-identical to torchvision's ToTensor, dataset.py:90, and to
 * srb_image_to_tensor with scale 1/255) -- so the fp32 copy of the HR target, 4x its bytes, never exists on the device.
 */
int srb_conv_fprop_loss(const srb_conv_params *p, const srb_tensor4 *x, const float *w, const float *bias,
                        const srb_tensor4 *target, int loss_kind, const srb_tensor4 *y, const srb_tensor4 *dz, int dz_unshuffled,
                        float *loss, void *ws, size_t ws_bytes, void *stream);

/*
 * torch.optim.Adam.step() (espcn.py:79,131; edsr.py:93,155; amsgrad off) as ONE launch over flat fp32 buffers: p, g, m, v hold all n
 * parameters / gradients / first / second moments of the model back to back (srb200.FlatAdam builds them as views).  state[0] is the
 * step count (float, device resident so that CUDA-graph replays advance it), state[1] scratch (both start at 0).  Same arithmetic
 * as torch's fused Adam: g += wd*p; m += (g-m)(1-b1); v = b2 v + (1-b2) g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 */
int srb_adam_step_flat(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float *state, void *stream);

/*
 * Opt-in packed-weight cache.  By default every srb_conv_fprop / srb_conv_dgrad / srb_conv_fprop_loss call re-packs its filter into
 * the kernel's operand layout (a 3-8 us launch in front of the convolution).  After srb_weight_cache_enable(1) the packed copies
 * persist in buffers the LIBRARY allocates (the one exception to "the caller owns all device memory"; released by
 * srb_weight_cache_enable(0)), conv calls skip their pack launch, and the caller must call srb_weight_cache_repack(stream) after
 * EVERY update of the weights (optimizer step, load_state_dict): one launch re-packs every cached filter.  Entries are created
 * on first use outside stream capture (run one eager step before capturing a CUDA graph); a miss during capture falls back to the
 * per-call pack.  Keyed by the filter's device pointer and layer geometry; math = EXACT never caches (its launches get temporaries).
 */
int srb_weight_cache_enable(int on);
int srb_weight_cache_repack(void *stream);
int srb_weight_cache_entries(void);

/* x[i] *= *g for n contiguous floats unless *g == 1 (device scalar; the upstream gradient of a scalar loss). */
int srb_scale_by_scalar(float *x, int64_t n, const float *g, int round_to_tf32, void *stream);

/*
 * Activation backward (threshold_backward / _prelu_kernel_backward / leaky_relu_backward):
 *   dz = dy * act'(.)   elementwise over the (N,Cout,Ho*ps,Wo*ps) tensor; for PReLU also
 *   *dalpha += sum(dy * z * [z <= 0]).   `ref` is y for RELU/LRELU and the saved preact z for PRELU.
 * dz is written RN-rounded to tf32 when p->math selects the tensor path for the following kernels.
 */
int srb_act_bwd(const srb_conv_params *p, const srb_tensor4 *dy, const srb_tensor4 *ref, const float *alpha,
                const srb_tensor4 *dz, float *dalpha, void *stream);

/*
 * Data gradient.  Replaces cudnn_convolution_backward_input (and, for transposed == 1, the
 * backward of ConvTranspose2d).  dz is the gradient w.r.t. PixelShuffle(z), i.e. in y's layout;
 * the un-shuffle is folded into the load addressing.   dx (N,Cin,H,W).
 *   relu_mask  optional (NULL = none), same shape as dx: dx = relu_mask > 0 ? dx : 0 is applied in the epilogue.
 *              Passing the layer's own input x when x = ReLU(.) of the previous layer folds that layer's
 *              threshold_backward into this kernel (base_networks.py:69 backward), saving one pass over dx.
 *   relu_bits  optional (NULL = none): the same mask in the packed form srb_conv_fprop's relu_bits wrote for the layer
 *              that produced x (geometry of dx: word ((n*H+iy)*W+ix)*(Cin/16) + c/16).  Costs 1/32 of the float mask's
 *              traffic.  Tensor-path dgrad with Cin % 16 == 0 only; else SRB_EUNSUPPORTED (pass relu_mask instead).
 */
int srb_conv_dgrad(const srb_conv_params *p, const srb_tensor4 *dz, const float *w, const srb_tensor4 *relu_mask,
                   const uint16_t *relu_bits, const srb_tensor4 *dx, void *ws, size_t ws_bytes, void *stream);

/*
 * The same with a second gradient folded in: dx = dgrad(dz) + add (then the ReLU mask, if any).  `add` has dx's shape and dtype.
 * This is the fan-out sum autograd performs at the input of a ResnetBlock (`torch.add(out, residual)`, base_networks.py:149:
 * x feeds conv1 AND the skip connection, so dL/dx = dgrad_conv1 + dL/dy) done in the dgrad kernel's epilogue instead of a
 * separate add pass over dx.  Stride-1 Conv2d without PixelShuffle on the tensor path (math auto / tf32 / bf16) only; anything
 * else returns SRB_EUNSUPPORTED before launching (the caller then adds the two tensors itself).  add == NULL: srb_conv_dgrad.
 */
int srb_conv_dgrad_add(const srb_conv_params *p, const srb_tensor4 *dz, const float *w, const srb_tensor4 *relu_mask,
                       const uint16_t *relu_bits, const srb_tensor4 *add, const srb_tensor4 *dx, void *ws, size_t ws_bytes,
                       void *stream);

/*
 * Weight + bias gradient.  Replaces cudnn_convolution_backward_weight + the bias aten::sum.
 *   dw   same shape as w (fp32, contiguous), overwritten (accumulate == 0) or added to (accumulate != 0)
 *   db   (Cout*ps*ps) or NULL, same accumulate rule
 *   scale multiplies the result before it is stored/added (1/world_size for data parallel).
 */
int srb_conv_wgrad(const srb_conv_params *p, const srb_tensor4 *x, const srb_tensor4 *dz, float *dw, float *db,
                   float scale, int accumulate, void *ws, size_t ws_bytes, void *stream);

/*
 * pixel_unshuffle into channels_last (the inverse of the fused PixelShuffle store, ATen pixel_unshuffle):
 *   out[n, c*r*r + i*r + j, h, w] = dz[n, c, h*r+i, w*r+j],   out must be NHWC (sc == 1), values RN-rounded
 * to tf32.  Used by the backward of PixelShuffle layers so that their dgrad/wgrad can run on the tensor
 * path with p->ps = 1, Cout = Cout*r*r.  dz is (N,Cout,Ho*ps,Wo*ps); out is (N,Cout*ps*ps,Ho,Wo).
 */
int srb_pixel_unshuffle(const srb_conv_params *p, const srb_tensor4 *dz, const srb_tensor4 *out, void *stream);

/* Stand-alone PReLU (the raw nn.PReLU() of fsrcnn.py:26): y = x > 0 ? x : alpha*x over n contiguous floats. */
int srb_prelu_fwd(const float *x, const float *alpha, float *y, int64_t n, void *stream);
int srb_prelu_bwd(const float *x, const float *dy, const float *alpha, float *dx, float *dalpha, int64_t n,
                  void *stream);

/*
 * Regression losses with mean reduction over n contiguous floats: kind 0 = MSE (nn.MSELoss, srcnn.py:84, espcn.py:84,
 * vdsr.py:96), kind 1 = L1 (nn.L1Loss, edsr.py:98).  y and t must share one dense layout.
 *   fwd: *loss = mean((y-t)^2) or mean(|y-t|); deterministic two-stage sum; ws >= srb_loss_workspace_bytes().
 *   bwd: dy = *grad_loss * d loss / d y in one pass (mse_backward / l1_loss backward: sign(y-t), 0 at equality).
 */
size_t srb_loss_workspace_bytes(void);
int srb_loss_fwd(int kind, const float *y, const float *t, int64_t n, float *loss, void *ws, size_t ws_bytes, void *stream);
int srb_loss_bwd(int kind, const float *y, const float *t, int64_t n, const float *grad_loss, float *dy, void *stream);

/*
 * ToTensor on the device (torchvision.transforms.ToTensor, dataset.py:90,94,98): uint8 NHWC image batch -> fp32 NCHW,
 * dst = src * scale; with scale == 1/255 (as a float) the result is the correctly rounded src / 255, i.e. bit-identical to
 * ToTensor's `.div(255)` (a plain multiply by 1/255 is 1 ulp off for 126 of the 256 byte values).  Lets the host ship the 1-byte pixels its decoder produced instead of floats
 * (4x fewer PCIe bytes per training step).
 */
int srb_image_to_tensor(const uint8_t *src_nhwc, float *dst_nchw, int32_t N, int32_t H, int32_t W, int32_t C, float scale,
                        void *stream);

/*
 * utils.img_interp(imgs, scale, 'bicubic') (utils.py:242-269; used every step by SRCNN / VDSR / DRCN, srcnn.py:120-121) on the
 * device, optionally followed by utils.shave(., shave) (utils.py:197-205): x fp32 NCHW in [0,1] -> y fp32 NCHW
 * (N, C, TH - 2*shave, TW - 2*shave).  Bit-exact with the reference's per-image PIL loop: ToPILImage truncation to 8 bits,
 * Pillow's two-pass fixed-point bicubic (a = -0.5), ToTensor.  The per-axis tables are Pillow's precompute_coeffs +
 * normalize_coeffs_8bpc results, built by the host: bounds = (first tap, tap count) per output index, coeffs = ksize 22-bit
 * integers per output index.
 */
int srb_img_interp_bicubic(const float *x, float *y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t TH, int32_t TW,
                           const int32_t *bounds_w, const int32_t *coeffs_w, const int32_t *bounds_h, const int32_t *coeffs_h,
                           int32_t ksize, int32_t shave, void *stream);

/* Round n contiguous floats to tf32 (round-to-nearest, ties away) in place or out of place. */
int srb_round_tf32(const float *x, float *y, int64_t n, void *stream);

/*
 * The layers around the conv stacks of SRGAN.  Dense fp32 tensors; reductions are deterministic.
 *
 * BatchNorm2d over a dense channels_last tensor of P = N*H*W pixels x C channels (base_networks.py:46,117,137,145,161), fused
 * with what follows it in the blocks:   y = act(gamma * (x - mean) * invstd + beta) + residual
 *   training != 0: batch statistics (biased variance), running_mean / running_var updated with `momentum` (unbiased variance),
 *   like nn.BatchNorm2d; training == 0: the running statistics.  save_mean / save_invstd (C floats each) feed the backward.
 *   act: srb_act (PReLU: alpha = the single slope); residual: NULL or a tensor like y (ResnetBlock's `bn(conv2(.)) + x`).
 * Backward: dx, dgamma, dbeta (scaled / accumulated like srb_conv_wgrad), *dalpha += PReLU slope gradient; the gradient of
 * the residual is dy itself.  Workspace: srb_bn_workspace_bytes(C).
 */
size_t srb_bn_workspace_bytes(int32_t C);
int srb_bn_fwd(const float *x, float *y, int64_t P, int32_t C, const float *gamma, const float *beta, float *running_mean,
               float *running_var, int training, float momentum, float eps, float *save_mean, float *save_invstd, int act,
               float slope, const float *alpha, const float *residual, int round_to_tf32, void *ws, size_t ws_bytes, void *stream);
int srb_bn_bwd(const float *x, const float *dy, float *dx, int64_t P, int32_t C, const float *gamma, const float *beta,
               const float *save_mean, const float *save_invstd, int act, float slope, const float *alpha, float *dgamma,
               float *dbeta, float *dalpha, float scale, int accumulate, int round_to_tf32, void *ws, size_t ws_bytes, void *stream);

/* nn.Linear (DenseBlock, base_networks.py:7,29-31; srgan.py:66-70): y[B,O] = x[B,I] w[O,I]^T + bias.  The weight matrix is
 * streamed once per 16 batch rows (HBM-bound: 8 flop/byte); I % 4 == 0.  Backward: dx (may be NULL), dw / db (may be NULL)
 * scaled / accumulated like srb_conv_wgrad.  Workspace: srb_linear_workspace_bytes(I, O). */
size_t srb_linear_workspace_bytes(int32_t I, int32_t O);
int srb_linear_fwd(const float *x, const float *w, const float *bias, float *y, int32_t B, int32_t I, int32_t O, void *ws,
                   size_t ws_bytes, void *stream);
int srb_linear_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw, float *db, int32_t B, int32_t I, int32_t O,
                   float scale, int accumulate, void *ws, size_t ws_bytes, void *stream);

/* nn.MaxPool2d(2) on a dense channels_last tensor (VGG19 features[4], srgan.py:84-90); idx: one byte per output (argmax 0..3). */
int srb_maxpool2_fwd(const float *x, float *y, uint8_t *idx, int32_t N, int32_t C, int32_t H, int32_t W, void *stream);
int srb_maxpool2_bwd(const float *dy, const uint8_t *idx, float *dx, int32_t N, int32_t C, int32_t H, int32_t W, void *stream);

/* nn.BCELoss (mean) over n probabilities (srgan.py:157,276-297) and its backward (dy = *grad_loss * d loss / d y). */
int srb_bce_fwd(const float *y, const float *t, int32_t n, float *loss, void *stream);
int srb_bce_bwd(const float *y, const float *t, int32_t n, const float *grad_loss, float *dy, void *stream);

/*
 * Data-parallel gradient exchange (the reference has none; BASELINE.json asks for it): in-place sum of
 * the flat fp32 gradient buffer over `world` GPUs of one NVLink/NVSwitch box, ONE kernel per rank, no NCCL on this path.
 *   bufs_dev  device array of `world` device pointers: every rank's flat buffer (peer-mapped symmetric memory, n floats each)
 *   pads_dev  device array of `world` device pointers: every rank's signal pad (>= 32 zero-initialised uint32 words each)
 *   state2    two zero-initialised uint32 words in LOCAL device memory (epoch, block counter) owned by this bucket
 * Every rank of the group must enqueue the call the same number of times.  Rank r reduces slice r of all buffers in rank order
 * and stores the result into slice r of all buffers, so all replicas end up bit-identical (deterministic).  world <= 16.
 */
int srb_allreduce_inplace(float *const *bufs_dev, uint32_t *const *pads_dev, int32_t rank, int32_t world, int64_t n,
                          uint32_t *state2, void *stream);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t srb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H_ */
