"""Whole-step CUDA-graph replay for the training loop body (espcn.py:126-131 and its siblings).

The small SR nets are host-launch bound when every kernel is launched from Python (ESPCN cfg2: ~35 launches per 0.65 ms
step).  libsrb200 never allocates, synchronises or reads device data on the host, so the whole step -- zeroing of the
non-conv gradients, forward, loss, backward (autograd), gradient clipping, optimizer -- can be captured once per input
slot and replayed.  Tensor maps are encoded at capture time against the graph's private memory pool, which is why the
inputs live in fixed "slots" that the caller refills (device-to-device or pinned-host-to-device copies).

    stepper = srb200.TrainStepGraphs(net, loss_fn, optimizer, bucket, slots=[(x0, t0), (x1, t1)], clip_norm=None)
    loss = stepper.step(i)         # replays slot i; returns the 0-dim loss tensor of that slot (device)

Data parallel: the NCCL all-reduce of the flat gradient buffer sits between the backward graph and the optimizer graph.
With `capture_allreduce=True` it is captured into the backward graph (NCCL supports stream capture); if that capture
raises, the error is kept in `comm_capture_error`, a warning is issued and the all-reduce runs eagerly between two graphs.

Requirements (checked): the optimizer state must already exist (one eager step first).  Limitation: a Python-float `lr` is
baked into the captured optimizer kernel, so `param_group['lr'] = ...` decay (vdsr.py:108-112) has no effect on replays --
construct the optimizer with a tensor lr (`lr=torch.tensor(1e-5, device=...)`) and update it in place.
"""
import torch
import torch.distributed  # noqa: F401  (DistBackendError)

from . import _lib


class TrainStepGraphs:
    def __init__(self, net, loss_fn, optimizer, bucket, slots, clip_norm=None, capture_allreduce=True, forward_loss=None,
                 weight_cache=False):
        self.net, self.loss_fn, self.opt, self.bucket = net, loss_fn, optimizer, bucket
        # weight_cache: libsrb200 keeps the packed filters of every conv layer and this stepper re-packs them all with ONE launch
        # behind optimizer.step() (captured into the optimizer graph) instead of one pack launch in front of every conv.
        # Enable it (srb200.enable_weight_cache()) BEFORE the eager warm-up steps: cache entries cannot be created during capture.
        self.weight_cache = bool(weight_cache)
        if self.weight_cache:
            from . import functional as F
            F.enable_weight_cache(True)
        # forward_loss(x, t) -> loss replaces loss_fn(net(x), t) (e.g. srb200.FusedLoss: criterion inside the last conv's epilogue)
        self.forward_loss = forward_loss if forward_loss is not None else (lambda x, t: loss_fn(net(x), t))
        self.slots = list(slots)
        self.clip_norm = clip_norm
        self.dev = self.slots[0][0].device
        self.fused_comm = False
        self.comm_capture_error = None  # why the NCCL all-reduce could not be captured (None: captured or not requested)
        self.launches_per_step = 0
        self._check_optimizer_state()
        self._capture(capture_allreduce and bucket.world > 1)

    def _check_optimizer_state(self):
        """Adam's exp_avg / step and SGD's momentum buffers are created lazily by the first optimizer.step(); captured into
        the tail graph that initialisation would be replayed (and reset the moments) on every step.  Require it to exist."""
        opt = self.opt
        needs_state = any(g.get("momentum", 0) for g in opt.param_groups) or isinstance(opt, (torch.optim.Adam, torch.optim.AdamW))
        if not needs_state:
            return
        for g in opt.param_groups:
            for p in g["params"]:
                if p.requires_grad and len(opt.state.get(p, {})) == 0:
                    raise RuntimeError(
                        "TrainStepGraphs: optimizer state is missing for a parameter -- run one eager step "
                        "(stepper-less: bucket.begin_step(); loss.backward(); optimizer.step()) before capturing, "
                        "otherwise the lazy state initialisation is captured and replayed every step")
            lr = g.get("lr")
            if not torch.is_tensor(lr):
                self.lr_baked = True  # scalar lr is baked into the captured optimizer kernel; see the class docstring

    # -- eager version of the same step (warm-up, debugging, instrumentation) ---------------------------------------
    def eager_step(self, x, t):
        self.bucket.begin_step()
        loss = self.forward_loss(x, t)
        loss.backward()
        self.bucket.all_reduce()
        if self.clip_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.clip_norm)
        self.opt.step()
        self._repack()
        return loss

    def _repack(self):
        if self.weight_cache:
            from . import functional as F
            F.repack_weights(self.dev)

    def _capture_once(self, with_comm):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        pool = torch.cuda.graph_pool_handle()
        fwd_bwd, losses, counts = [], [], []
        with torch.cuda.stream(side):
            for x, t in self.slots:
                g = torch.cuda.CUDAGraph()
                out = torch.zeros((), device=self.dev)
                c0 = _lib.launch_count()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    self.bucket.begin_step()
                    loss = self.forward_loss(x, t)
                    loss.backward()
                    out.copy_(loss.detach())
                    if with_comm:
                        self.bucket.all_reduce()
                counts.append(_lib.launch_count() - c0)
                fwd_bwd.append(g)
                losses.append(out)
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, pool=pool, stream=side):
                if self.clip_norm is not None:
                    torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.clip_norm)
                self.opt.step()
                self._repack()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.fwd_bwd, self.tail, self.losses = fwd_bwd, tail, losses
        self.launches_per_step = counts[0]
        self.fused_comm = with_comm

    def _capture(self, want_comm):
        if want_comm:
            try:
                self._capture_once(True)
                return
            except (RuntimeError, torch.distributed.DistBackendError) as e:
                # NCCL refused stream capture in this setup: report it, keep the collective eager between the two graphs
                import warnings
                self.comm_capture_error = "%s: %s" % (type(e).__name__, e)
                warnings.warn("TrainStepGraphs: all-reduce not captured (%s); running it eagerly between the graphs"
                              % self.comm_capture_error)
                torch.cuda.synchronize(self.dev)
        self._capture_once(False)

    def step(self, i):
        """Replay the step on slot i (the caller has already put that step's batch into the slot tensors)."""
        self.fwd_bwd[i].replay()
        if not self.fused_comm:
            self.bucket.all_reduce()
        self.tail.replay()
        return self.losses[i]
