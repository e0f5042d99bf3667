#!/usr/bin/env python
"""bench.py -- SR training images/sec on B200 through the libsrb200 hot path.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                   (the reference's CPU PyTorch path, host cores)
    python bench.py --impl cudnn ...                       (the same nets on stock torch.nn -> cuDNN: the GPU baseline
                                                            the north_star names; as-is and tuned variants)

Workload at N=1 (and per rank, weak scaling, at N>1): BASELINE.json configs[1] -- ESPCN x4, 64x64 LR,
batch 128, fwd + MSE + bwd + Adam (espcn.py:126-131), synthetic uniform inputs, weights from the
net's own N(0, 0.02) init under seed 0.  One "step" = one such optimizer step.

Prints ONE JSON line (rank 0).  value = images/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = same metric through the public API with HOST buffers: every step's LR/HR batches are copied from pinned host
memory as the uint8 HWC pixels an image decoder produces (dataset.py:90 applies ToTensor on the host; here
srb200.image_to_tensor does it on the device) and the loss is read back every step;
roofline = the dominant kernel of the step, algorithmic bytes (or flops) per launch / its mean duration,
timed live with CUDA events in a second instrumented pass of the same region;
cpu_baseline = the CPU oracle port of the reference nets (oracle/torch_ref.py) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))

import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

WORKLOADS = {
    # name: (model key, ctor args, per-GPU batch, LR input HxW, loss, optimizer key)
    "espcn_x4_b128_lr64": ("espcn", (3, 64, 4), 128, (64, 64), "mse", "espcn"),
    "vdsr_b64_128": ("vdsr", (3, 64, 18), 64, (128, 128), "mse", "vdsr"),
    "edsr64_x4_b32_lr32": ("edsr", (3, 64, 16), 32, (32, 32), "l1", "edsr"),
    "edsr256_x4_b32_lr32": ("edsr", (3, 256, 32), 32, (32, 32), "l1", "edsr"),  # cfg4 (--math bf16 = its dtype; auto = TF32 on fp32 storage)
    "srcnn_x2_b16": ("srcnn", (3, 64), 16, (64, 64), "mse", "srcnn"),
    # FSRCNN x4 (fsrcnn.py:99 ctor, SURVEY 8a: parity-only net; timed to cover the k9 s4 transposed conv on the tensor path)
    "fsrcnn_x4_b16_lr32": ("fsrcnn", (3, 4, 56, 12, 4), 16, (32, 32), "mse", "fsrcnn"),
    # cfg5: SRGAN adversarial iteration as written (srgan.py:256-310): G(3,64,16), D(3,64,128), VGG19[:9] features
    "srgan_x4_b16": ("srgan", (3, 64, 16), 16, (32, 32), "bce+mse+vgg", "srgan"),
}
DEFAULT_WORKLOAD = "espcn_x4_b128_lr64"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        pk = {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
              "source": "measured"}
    else:
        pk = {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}
    # TF32 peak: measured with the MEASURED_PEAKS method (cuBLAS 8192^3, burst + sustained) by tools/measure_tf32_peak.py;
    # bf16/2 only when that file is missing
    pk["tf32_burst"], pk["tf32_sustained"], pk["tf32_source"] = pk["bf16_burst"] / 2, pk["bf16_sustained"] / 2, "bf16/2 (assumed)"
    tpath = os.path.join(ROOT, "profiles", "tf32_peak.json")
    if os.path.isfile(tpath):
        try:
            t = json.load(open(tpath))
            pk["tf32_burst"], pk["tf32_sustained"] = t["tf32_tflops"], t["tf32_tflops_sustained"]
            pk["tf32_source"] = "measured (profiles/tf32_peak.json: cuBLAS TF32 8192^3)"
        except Exception:
            pass
    return pk


def measure_srgan(ctx, a, workload, steps, warmup):
    """cfg5: one SRGAN adversarial iteration per step (srb200.host.srgan_step), the whole iteration replayed from one CUDA
    graph per input slot (two optimizers, BatchNorm statistics and both gradient exchanges included)."""
    import srb200
    from srb200 import _lib, host
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[workload]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    srb200.set_math(a.math)
    torch.manual_seed(0)
    G = srb200.models.SRGANGenerator(*margs)
    host.init_model("srgan", G)
    torch.manual_seed(1)
    D = srb200.models.SRGANDiscriminator(3, 64, 4 * h)
    host.init_model("srgan", D)
    torch.manual_seed(2)
    FE = srb200.models.FeatureExtractor()
    for m in FE.modules():
        if isinstance(m, torch.nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            torch.nn.init.constant_(m.bias, 0)
    for m in (G, D, FE):
        m.to(dev).train()
    go, do = host.make_srgan_optimizers(G, D, lr=1e-5, capturable=not a.no_graph)
    # one GPU: plain autograd accumulation measured marginally faster than the flat buckets (16.8 vs 17.1 ms / iteration)
    bg = srb200.GradBucket(G, world_size=world) if (world > 1 or a.bucket) else None
    bd = srb200.GradBucket(D, world_size=world) if (world > 1 or a.bucket) else None
    gen = torch.Generator().manual_seed(1 + rank)
    host_x = [synth_images(batch, h, w, gen) for _ in range(3)]
    host_t = [synth_images(batch, 4 * h, 4 * w, gen) for _ in range(3)]
    stage_x = [torch.empty(t.shape, dtype=torch.uint8, device=dev) for t in host_x]
    stage_t = [torch.empty(t.shape, dtype=torch.uint8, device=dev) for t in host_t]
    dev_x = [srb200.image_to_tensor(t.to(dev)) for t in host_x]
    dev_t = [srb200.image_to_tensor(t.to(dev)) for t in host_t]

    use_wcache = not a.no_weight_cache
    if use_wcache:
        srb200.enable_weight_cache(True)
    repack = (lambda: srb200.repack_weights(dev)) if use_wcache else None

    def step(i):
        return host.srgan_step(G, D, FE, go, do, dev_x[i], dev_t[i], bucket_g=bg, bucket_d=bd, after_update=repack)

    for i in range(warmup):
        step(i % 3)
    torch.cuda.synchronize()
    graphs, outs, launches = [], [], 0
    if not a.no_graph:
        # Gradients left over from the eager warm-up live in the ordinary allocator pool.  srgan_step accumulates into G's
        # (D_loss.backward() reaches G, srgan.py:283) before g_opt.zero_grad() drops them, so a captured graph would keep raw
        # pointers to blocks that torch.cuda.graph's own empty_cache() releases when the next capture starts.  Drop them first:
        # every gradient is then created inside the capture, in the graph's private pool.
        if bg is None:
            go.zero_grad(set_to_none=True)
            do.zero_grad(set_to_none=True)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        pool = torch.cuda.graph_pool_handle()
        with torch.cuda.stream(side):
            for i in range(3):
                g = torch.cuda.CUDAGraph()
                c0 = _lib.launch_count()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    dl, gl = step(i)
                    out = (dl + gl).clone()
                launches = _lib.launch_count() - c0
                graphs.append(g)
                outs.append(out)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()

    def run(i):
        if graphs:
            graphs[i].replay()
            return outs[i]
        dl, gl = step(i)
        return dl + gl

    copy_stream = torch.cuda.Stream(device=dev)
    loss_hosts = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_evs = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        s2 = i % 3
        with torch.cuda.stream(copy_stream):
            stage_x[s2].copy_(host_x[s2], non_blocking=True)
            stage_t[s2].copy_(host_t[s2], non_blocking=True)
            srb200.image_to_tensor(stage_x[s2], out=dev_x[s2])
            srb200.image_to_tensor(stage_t[s2], out=dev_t[s2])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def timed(n, e2e):
        ctx.barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            copy_stream.wait_event(e0)
            nxt = prefetch(0)
        for i in range(n):
            if e2e:
                ev = nxt
                if i + 1 < n:
                    nxt = prefetch(i + 1)
                main.wait_event(ev)
                loss = run(i % 3)
                if "loss" not in E2E_SKIP:
                    loss_hosts[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_evs[i % 2].record(main)
                if i > 0 and "sync" not in E2E_SKIP:
                    loss_evs[(i - 1) % 2].synchronize()
            else:
                run(i % 3)
        if e2e:
            loss_evs[(n - 1) % 2].synchronize()
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    for i in range(3):
        run(i)
    l0 = _lib.launch_count()
    with Clocks(ctx.local_rank) as ck:
        ms = timed(steps, False)
        n_launch = launches * steps if graphs else _lib.launch_count() - l0
        ms_e2e = timed(steps, True)
    imgs = batch * world * steps
    # algorithmic flops of one iteration as written: 2 G fwd + 2 G bwd + 3 D fwd + 3 D bwd + 2 VGG fwd + 1 VGG bwd
    flops_iter = 1083.8e9 * batch / 16.0
    pk = peaks()
    tfl = flops_iter / (ms / steps * 1e-3) / 1e12
    res = {"workload": workload, "value": imgs / (ms * 1e-3), "ms_per_step": ms / steps, "clocks": ck.summary(), "steps": steps,
           "gpu_launches": int(n_launch),
           "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                   "h2d_bytes_per_step": int(host_x[0].numel() + host_t[0].numel()), "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_e2e / steps, "host_format": "uint8 HWC pixels, pinned; ToTensor on the device"},
           "roofline": {"bound": "tensor", "achieved": tfl, "peak": pk["tf32_sustained"], "unit": "TFLOP/s",
                        "frac": tfl / pk["tf32_sustained"], "traffic": None, "kernel": "whole adversarial iteration",
                        "layer": "G(3,64,16) + D(3,64,128) + VGG19[:9], 1083.8 GFLOP per 16-image iteration as written (SURVEY.md 8d)",
                        "peak_source": pk["source"] + " sustained; tf32: " + pk["tf32_source"],
                        "algorithmic_flops": flops_iter},
           "kernels": [], "comm": None if world == 1 else (bg.comm if bg is not None else None)}
    graphs.clear()
    return res


def out_shape(model_key, args, n, h, w):
    if model_key == "srcnn":
        return (n, 3, h - 16, w - 16)
    if model_key == "espcn":
        return (n, 3, (h - 8) * args[2], (w - 8) * args[2])
    if model_key == "vdsr":
        return (n, 3, h, w)
    if model_key == "fsrcnn":  # conv5 p0, then ConvTranspose2d(k9, s=r, p3, op1): (h-4)*r
        return (n, 3, (h - 4) * args[1], (w - 4) * args[1])
    return (n, 3, 4 * h, 4 * w)


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of the reference nets on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, steps, warmup, budget_s=20.0, sample_batch=None):
    from oracle import torch_ref as R
    model_key, args, batch, (h, w), loss_kind, opt_key = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = sample_batch or batch
    if model_key == "srgan":
        G, D, FE = R.build("srgan_g", args, seed=0), R.build("srgan_d", (3, 64, 4 * h), seed=1), R.build_feature_extractor()
        go, do = R.make_srgan_optimizers(G, D, lr=1e-5)
        gen = torch.Generator().manual_seed(1)
        lr_img, hr_img = torch.rand((b, 3, h, w), generator=gen), torch.rand((b, 3, 4 * h, 4 * w), generator=gen)
        for _ in range(max(1, warmup)):
            R.srgan_step(G, D, FE, go, do, lr_img, hr_img)
        t0 = time.perf_counter()
        done = 0
        while done < steps and (time.perf_counter() - t0) < budget_s:
            R.srgan_step(G, D, FE, go, do, lr_img, hr_img)
            done += 1
        dt = time.perf_counter() - t0
        return {"value": b * done / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": "%d adversarial iterations of batch %d (%s, srgan.py:256-310 restated, torch %s CPU/oneDNN)" % (
                    done, b, workload, torch.__version__), "ms_per_step": 1e3 * dt / max(done, 1), "steps": done}
    net = R.build(model_key, args, seed=0)
    opt = R.make_optimizer(opt_key, net.parameters(), lr=1e-5)
    gen = torch.Generator().manual_seed(1)
    x = torch.rand((b, 3, h, w), generator=gen)
    t = torch.rand(out_shape(model_key, args, b, h, w), generator=gen)
    for _ in range(max(1, warmup)):
        R.train_step(opt_key, net, opt, x, t)
    t0 = time.perf_counter()
    done = 0
    while done < steps and (time.perf_counter() - t0) < budget_s:
        R.train_step(opt_key, net, opt, x, t)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": b * done / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of batch %d (%s, fwd+loss+bwd+optimizer, torch %s CPU/oneDNN)" % (
                done, b, workload, torch.__version__), "ms_per_step": 1e3 * dt / max(done, 1), "steps": done}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process (a sample per ~5 ms), nvidia-smi
    as the fallback when the NVML binding is unavailable."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples = []  # (sm_mhz, max_mhz, [4 reason flags])
        self.stop = False
        self.index = index
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES remapping when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        self.samples.append((int(sm), int(mx), flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [s.strip() for s in out.strip().split(",")]
        if len(f) >= 6:
            self.samples.append((int(float(f[0])), int(float(f[1])), [f[2 + i].lower().startswith("active") for i in range(4)]))

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                if self.nvml is not None:
                    self.nvml = None  # fall back to nvidia-smi
            time.sleep(0.005 if self.nvml is not None else 0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons, "samples": len(mhz),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# per-kernel instrumentation (second pass): CUDA events around every C-ABI call, on the launching stream
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    def __init__(self, lib_mod):
        self.lib_mod = lib_mod
        self.records = []
        self.names = ["srb_conv_fprop", "srb_conv_dgrad", "srb_conv_wgrad", "srb_act_bwd"]
        self.orig = {}

    def __enter__(self):
        lib = self.lib_mod.lib
        for n in self.names:
            fn = getattr(lib, n)
            self.orig[n] = fn

            def wrapped(*a, _fn=fn, _n=n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = _fn(*a)
                e1.record()
                p = a[0]._obj
                self.records.append((_n, (p.N, p.Cin, p.H, p.W, p.Cout, p.kh, p.stride, p.pad, p.ps, p.transposed),
                                     e0, e1))
                return rc
            setattr(lib, n, wrapped)
        return self

    def __exit__(self, *a):
        for n, fn in self.orig.items():
            setattr(self.lib_mod.lib, n, fn)

    def table(self):
        torch.cuda.synchronize()
        agg = {}
        for n, key, e0, e1 in self.records:
            d = agg.setdefault((n, key), [0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
        return agg


def algorithmic(name, key):
    """(bytes, flops) of one launch: each operand/result touched once, fp32 storage."""
    N, Cin, H, W, Cout, k, st, pad, ps, tr = key
    if tr:
        Ho, Wo = (H - 1) * st - 2 * pad + k, (W - 1) * st - 2 * pad + k
        macs = N * H * W * Cin * Cout * k * k
        xb, yb, wb = N * Cin * H * W, N * Cout * Ho * Wo, Cin * Cout * k * k
    else:
        Ho, Wo = (H + 2 * pad - k) // st + 1, (W + 2 * pad - k) // st + 1
        co = Cout * ps * ps
        macs = N * Ho * Wo * co * Cin * k * k
        xb, yb, wb = N * Cin * H * W, N * co * Ho * Wo, co * Cin * k * k
    if name == "srb_conv_fprop":
        return 4 * (xb + wb + yb), 2 * macs
    if name == "srb_conv_dgrad":
        return 4 * (yb + wb + xb), 2 * macs
    if name == "srb_conv_wgrad":
        return 4 * (xb + yb + 4 * wb), 2 * macs
    return 4 * 3 * yb, 0  # act_bwd: dy, ref, dz


METRIC = "SR training images/sec (device-timed)"


def make_config(workload, world, no_graph, math):
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[workload]
    return {"workload": workload, "net": model_key, "per_gpu_batch": batch, "global_batch": batch * world,
            "lr_hw": [h, w], "loss": loss_kind, "optimizer": opt_key, "parallelism": "dp%d" % world, "math": math,
            "l2": "per-step working set (activations+grads ~1 GB) exceeds the 126 MB L2; inputs rotate over 3 batches",
            "launch": "eager" if no_graph else "cuda-graph replay (srb200.TrainStepGraphs: fwd+loss+bwd[+allreduce] graph per input slot, optimizer graph)"}


def bind_to_gpu_numa(index):
    """Run this process (and therefore first-touch its pinned staging buffers) on the NUMA node the GPU's PCIe root port
    belongs to: the e2e leg's host->device copies then do not cross the inter-socket link.  Returns a small record for the
    JSON line; a no-op (with the reason) where sysfs does not say.  SRB_NO_NUMA_BIND=1 disables it (A/B)."""
    if os.environ.get("SRB_NO_NUMA_BIND") == "1":
        return {"bound": False, "why": "SRB_NO_NUMA_BIND"}
    try:
        pr = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"bound": False, "why": "numa_node -1 for " + bus}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "no allowed cpu on node %d" % node}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus), "pci": bus}
    except Exception as e:  # sysfs layout, permissions: measurement proceeds unbound
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


class Ctx:
    """Process-wide state shared by the measured workloads: rank / world / device / barrier."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = None
        self.dist = None

    def init_cuda(self, need_dist=True):
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device for this --impl: the engine has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa(self.local_rank)
        if self.world > 1 and need_dist:
            import torch.distributed as dist
            # rank 0 prints exactly ONE JSON line on stdout: NCCL's log (its version banner appears from WARN level up) goes
            # to stderr unless the user pointed it somewhere else
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.dist is None:
            return ms
        t = torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()


# diagnostic only (the e2e number of such a run is not an e2e number): SRB_E2E_SKIP=copy,convert,loss,sync drops parts of the e2e leg
E2E_SKIP = set(filter(None, os.environ.get("SRB_E2E_SKIP", "").split(",")))


def synth_images(batch, h, w, gen):
    """Synthetic uint8 HWC image batch (what a decoder hands to dataset.py:90's ToTensor), pinned."""
    return torch.randint(0, 256, (batch, h, w, 3), generator=gen, dtype=torch.uint8).pin_memory()


def measure_srb(ctx, a, workload, steps, warmup, full):
    """Our arm on one workload.  full: also the e2e leg, the instrumented roofline pass and the kernel table."""
    import ctypes
    import srb200
    from srb200 import _lib
    from srb200 import host
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[workload]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    srb200.set_math(a.math)
    torch.manual_seed(0)
    net = srb200.models.MODELS[model_key](*margs)
    host.init_model(model_key, net)
    net.to(dev).train()
    bucket = srb200.GradBucket(net, world_size=world)
    if opt_key in ("espcn", "edsr") and not a.torch_optimizer:
        # Adam (espcn.py:79, edsr.py:93) as ONE libsrb200 launch over the flat parameter / gradient / moment buffers
        opt = srb200.FlatAdam(bucket, lr=1e-5, betas=(0.9, 0.999), eps=1e-8)
    else:
        opt = host.make_optimizer(opt_key, net.parameters(), lr=1e-5, capturable=not a.no_graph)
    lossf = host.loss_for(model_key, fused=True)
    # criterion evaluated inside the last conv's epilogue where the net allows it (ESPCN, SRCNN, EDSR), plain fused kernels else
    fwd_loss = srb200.FusedLoss(net, "l1" if loss_kind == "l1" else "mse") if not a.no_loss_fusion else (lambda x, t: lossf(net(x), t))
    oshape = out_shape(model_key, margs, batch, h, w)

    # host side: uint8 HWC pixels (pinned).  Device side: the LR input is an fp32 NCHW slot filled by srb200.image_to_tensor; the
    # HR target stays the uint8 HWC image -- srb200.FusedLoss reads it as byte/255 inside the last conv's epilogue (nets whose
    # loss cannot be fused convert it inside the step), so the 4x larger fp32 copy of the target is never written or read
    gen = torch.Generator().manual_seed(1 + rank)
    host_x = [synth_images(batch, h, w, gen) for _ in range(3)]
    host_t = [synth_images(batch, oshape[2], oshape[3], gen) for _ in range(3)]
    stage_x = [torch.empty(t.shape, dtype=torch.uint8, device=dev) for t in host_x]
    dev_x = [srb200.image_to_tensor(t.to(dev)) for t in host_x]
    t_u8 = not a.no_loss_fusion
    stage_t = [t.to(dev) for t in host_t]
    dev_t = stage_t if t_u8 else [srb200.image_to_tensor(t) for t in stage_t]
    clip = host.VDSR_CLIP if model_key == "vdsr" else None

    use_wcache = not a.no_weight_cache
    if use_wcache:
        srb200.enable_weight_cache(True)  # before the eager warm-up: that is when the cache entries are created

    def step(x, t):
        bucket.begin_step()
        loss = fwd_loss(x, t)
        loss.backward()
        bucket.all_reduce()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        if use_wcache:
            srb200.repack_weights(dev)  # one launch re-packs every conv filter for the next step
        return loss

    graphs = {}

    def step_slot(i):
        if not graphs:
            return step(dev_x[i], dev_t[i])
        return graphs["stepper"].step(i)

    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        """Step i's batch: pinned uint8 -> device staging (PCIe) on the copy stream, overlapping the previous step's kernels.
        Staging slot i%3 was last read by step i-3 (host-synchronised since)."""
        s = i % 3
        with torch.cuda.stream(copy_stream):
            if "copy" not in E2E_SKIP:
                stage_x[s].copy_(host_x[s], non_blocking=True)
                stage_t[s].copy_(host_t[s], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def to_tensor(i):
        """ToTensor on the device (uint8 HWC staging -> the fp32 NCHW slot the step reads), on the MAIN stream in front of the step:
        beside the training kernels (on the copy stream) it cost 75 us per step -- the one-CTA-per-SM conv kernels own the whole
        register file, so a concurrent elementwise grid and the convs take turns on every SM -- in line it costs its own ~20 us."""
        s = i % 3
        if "convert" not in E2E_SKIP:
            srb200.image_to_tensor(stage_x[s], out=dev_x[s])
            if not t_u8:
                srb200.image_to_tensor(stage_t[s], out=dev_t[s])

    loss_hosts = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_evs = [torch.cuda.Event(), torch.cuda.Event()]
    step_done = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_stream = torch.cuda.Stream(device=dev)

    def timed(nsteps, e2e):
        ctx.barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        seen = 0.0
        if e2e:
            copy_stream.wait_event(e0)
            nxt = prefetch(0)
        for i in range(nsteps):
            if e2e:
                # every step: its batch crosses PCIe (prefetched one step ahead on the copy stream) and its loss is read by the
                # host (like `loss.data[0]`, srcnn.py:134) -- the read of step i-1 happens while step i runs, so the GPU never idles
                ev = nxt
                if i + 1 < nsteps:
                    nxt = prefetch(i + 1)
                main.wait_event(ev)
                to_tensor(i)
                loss = step_slot(i % 3)
                # the 4-byte read-back runs on its own stream behind an event: a memcpy node on the main stream would sit between
                # two steps' kernels (slot i%3's loss tensor is not rewritten before step i+3; the host has read it by then)
                step_done[i % 2].record(main)
                d2h_stream.wait_event(step_done[i % 2])
                with torch.cuda.stream(d2h_stream):
                    if "loss" not in E2E_SKIP:
                        loss_hosts[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)
                    loss_evs[i % 2].record(d2h_stream)
                if i > 0 and "sync" not in E2E_SKIP:
                    loss_evs[(i - 1) % 2].synchronize()
                    seen += float(loss_hosts[(i - 1) % 2][0])
            else:
                step_slot(i % 3)
        if e2e:
            loss_evs[(nsteps - 1) % 2].synchronize()
            seen += float(loss_hosts[(nsteps - 1) % 2][0])
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    for i in range(warmup):
        step(dev_x[i % 3], dev_t[i % 3])
    if not a.no_graph:
        st = srb200.TrainStepGraphs(net, lossf, opt, bucket, slots=list(zip(dev_x, dev_t)), clip_norm=clip, forward_loss=fwd_loss,
                                    weight_cache=use_wcache)
        graphs.update(stepper=st, launches=st.launches_per_step)
        for i in range(3):
            step_slot(i)
    l0 = _lib.launch_count()
    with Clocks(ctx.local_rank) as ck:
        ms = timed(steps, e2e=False)
        launches = graphs["launches"] * steps if graphs else _lib.launch_count() - l0
        ms_e2e = timed(steps, e2e=True) if full else None
    clocks = ck.summary()
    imgs = batch * world * steps
    res = {"workload": workload, "value": imgs / (ms * 1e-3), "ms_per_step": ms / steps, "clocks": clocks,
           "gpu_launches": int(launches), "steps": steps,
           "comm": None if world == 1 else "%s all-reduce of the flat gradient buffer, %s" % (
               {"peer": "libsrb200 one-kernel NVLink peer-memory", "nccl": "NCCL"}.get(bucket.comm, bucket.comm),
               "captured in the backward graph" if graphs and graphs["stepper"].fused_comm else "eager between graphs")}
    if getattr(bucket, "comm_fallback", None):
        res["comm_fallback"] = bucket.comm_fallback
    if graphs and getattr(graphs["stepper"], "comm_capture_error", None):
        res["comm_capture_error"] = graphs["stepper"].comm_capture_error
    if not full:
        graphs.clear()
        bucket.detach()
        return res
    res["e2e"] = {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                  "h2d_bytes_per_step": int(host_x[0].numel() + host_t[0].numel()), "d2h_bytes_per_step": 4,
                  "ms_per_step": ms_e2e / steps, "numa": ctx.numa,
                  "host_format": "uint8 HWC pixels, pinned; ToTensor (x/255, HWC->CHW) runs on the device: a kernel for the LR input, "
                                 "inside the fused loss epilogue for the HR target" if t_u8 else
                                 "uint8 HWC pixels, pinned; ToTensor (x/255, HWC->CHW) runs on the device"}

    # ---- instrumented pass: CUDA events around every C-ABI call on the launching stream --------------------------------
    # The host is slower than the GPU on the small nets, so each instrumented step first parks the GPU on a spin kernel long
    # enough for the host to enqueue the whole step: the spans then measure kernel execution, not launch latency.
    pk = peaks()
    t0 = time.perf_counter()
    step(dev_x[0], dev_t[0])
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    spin_cycles = int(1.5 * min(host_s, 0.02) * (clocks.get("sm_mhz") or 1900) * 1e6) + 200000
    with KernelTimer(_lib) as kt:
        ctx.barrier()
        for i in range(min(steps, 20)):
            torch.cuda._sleep(spin_cycles)
            step(dev_x[i % 3], dev_t[i % 3])
        tab = kt.table()
    total_ms = sum(v[0] for v in tab.values())
    (dname, dkey), (dms, dcnt) = max(tab.items(), key=lambda kv: kv[1][0])
    esize = 2 if a.math == "bf16" else 4
    by, fl = algorithmic(dname, dkey)
    by = by * esize // 4
    dur_s = dms / dcnt * 1e-3
    # kernels timed alone behind a spin kernel: the burst peak applies (B200_PROFILING.md)
    tc_peak = pk["bf16_burst"] if a.math == "bf16" else pk["tf32_burst"]
    t_hbm, t_tc = by / (pk["hbm_gbs"] * 1e9), fl / (tc_peak * 1e12)
    if t_hbm >= t_tc:
        roof = {"bound": "hbm", "achieved": by / dur_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": fl / dur_s / 1e12, "peak": tc_peak, "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None
    layer = "N%d Cin%d %dx%d Cout%d k%d s%d p%d ps%d" % dkey[:9]
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        rec = tr.get(dname + "|" + layer)
        if rec:
            roof["traffic"] = rec["traffic"]
            roof["traffic_source"] = "profiles/ncu_traffic.json (%s, %.1f us under ncu)" % (rec["kernel"], rec["time_us"])
    except Exception:
        pass
    p_ = _lib.ConvParams(dkey[0], dkey[1], dkey[2], dkey[3], dkey[4], dkey[5], dkey[5], dkey[6], dkey[7], 0, dkey[9],
                         dkey[8], 0, 0.2, srb200.functional._state["math"])
    roof.update({"kernel": dname, "layer": layer,
                 "tensor_path": bool(_lib.lib.srb_conv_uses_tensor_path(
                     ctypes.byref(p_), {"srb_conv_fprop": 0, "srb_conv_dgrad": 1, "srb_conv_wgrad": 2}.get(dname, 0), 1, 1)),
                 "us_per_launch": dur_s * 1e6, "share_of_kernel_time": dms / total_ms,
                 "peak_source": pk["source"] + ("" if roof["bound"] == "hbm" else
                                                " burst; " + ("bf16" if a.math == "bf16" else "tf32: " + pk["tf32_source"])),
                 "algorithmic_bytes": by, "algorithmic_flops": fl})
    res["roofline"] = roof
    res["kernels"] = [{"call": n, "layer": "N%d Cin%d %dx%d Cout%d k%d s%d p%d ps%d" % k[:9], "us": round(v[0] / v[1] * 1e3, 2),
                       "share": round(v[0] / total_ms, 4)}
                      for (n, k), v in sorted(tab.items(), key=lambda kv: -kv[1][0])]
    graphs.clear()
    bucket.detach()
    return res


# ------------------------------------------------------------------------------------------------
# --impl cudnn: the same nets on stock torch.nn (ATen -> cuDNN), the GPU baseline the north_star names
# ------------------------------------------------------------------------------------------------
def reference_net(model_key, margs):
    """The imported, unmodified reference class when /root/reference is present, else its line-by-line restatement
    (oracle/torch_ref.py, bit-identical: tests/test_oracle.py::test_oracle_equals_live_reference)."""
    from oracle import ref_import
    torch.manual_seed(0)
    if ref_import.available():
        mods = ref_import.load()
        cls = {"srcnn": ("srcnn", "Net"), "espcn": ("espcn", "Net"), "vdsr": ("vdsr", "Net"), "edsr": ("edsr", "Net")}[model_key]
        net = getattr(mods[cls[0]], cls[1])(*margs)
        net.weight_init()
        return net, "reference classes (/root/reference)"
    from oracle import torch_ref as R
    return R.build(model_key, margs, seed=0), "restated reference nets (oracle/torch_ref.py)"


def measure_cudnn(ctx, a, workload, variant, steps, warmup):
    import torch.nn.functional as TF
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[workload]
    dev, world = ctx.dev, ctx.world
    torch.backends.cudnn.benchmark = variant != "as-is"
    net, src = reference_net(model_key, margs)
    net = net.to(dev).train()
    cl = "cl" in variant
    gen = torch.Generator().manual_seed(1 + ctx.rank)
    oshape = out_shape(model_key, margs, batch, h, w)
    host_x = [torch.rand((batch, 3, h, w), generator=gen).pin_memory() for _ in range(3)]
    host_t = [torch.rand(oshape, generator=gen).pin_memory() for _ in range(3)]
    x = [t.to(dev) for t in host_x]
    t = [v.to(dev) for v in host_t]
    if cl:
        net = net.to(memory_format=torch.channels_last)
        x = [v.contiguous(memory_format=torch.channels_last) for v in x]
    graph = variant.endswith("-graph") and world == 1
    fused = {"fused": variant != "as-is"}
    if opt_key in ("espcn", "edsr"):
        opt = torch.optim.Adam(net.parameters(), lr=1e-5, capturable=graph, **fused)
    elif opt_key == "vdsr":
        opt = torch.optim.SGD(net.parameters(), lr=1e-5, momentum=0.9, weight_decay=1e-4, **fused)
    else:
        opt = torch.optim.SGD(net.parameters(), lr=1e-5, **fused)
    model = net
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[ctx.local_rank])
    lf = TF.l1_loss if loss_kind == "l1" else TF.mse_loss
    bf16 = "bf16" in variant

    def fwd_loss(xi, ti):
        if bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return lf(model(xi).float(), ti)
        return lf(model(xi), ti)

    def step(i, set_none=True):
        opt.zero_grad(set_to_none=set_none)
        loss = fwd_loss(x[i % 3], t[i % 3])
        loss.backward()
        if opt_key == "vdsr":
            torch.nn.utils.clip_grad_norm_(net.parameters(), 0.4)
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    losses = [None] * 3
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        pool = torch.cuda.graph_pool_handle()
        gs = []
        with torch.cuda.stream(side):
            for i in range(3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    losses[i] = step(i, set_none=False).detach()
                gs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def run(i):
            gs[i % 3].replay()
            return losses[i % 3]
        for i in range(3):
            run(i)
    else:
        def run(i):
            return step(i)
    loss_host = torch.zeros(1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        s = i % 3
        with torch.cuda.stream(copy_stream):
            x[s].copy_(host_x[s], non_blocking=True)
            t[s].copy_(host_t[s], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def timed(n, e2e):
        ctx.barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            copy_stream.wait_event(e0)
            nxt = prefetch(0)
        for i in range(n):
            if e2e:
                ev = nxt
                if i + 1 < n:
                    nxt = prefetch(i + 1)
                main.wait_event(ev)
                loss = run(i)
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
                main.synchronize()
            else:
                run(i)
        e1.record()
        ctx.barrier()
        return ctx.max_over_ranks(e0.elapsed_time(e1))

    with Clocks(ctx.local_rank) as ck:
        ms = timed(steps, False)
        ms_e2e = timed(steps, True)
    imgs = batch * world * steps
    return {"variant": variant, "value": imgs / (ms * 1e-3), "ms_per_step": ms / steps,
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(host_x[0].numel() + host_t[0].numel()) * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / steps, "host_format": "fp32 CHW tensors, pinned (ToTensor on the host, as dataset.py:90)"},
            "clocks": ck.summary(), "net_source": src, "allow_tf32": torch.backends.cudnn.allow_tf32,
            "cudnn": torch.backends.cudnn.version()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="srb200", choices=["srb200", "reference", "cudnn"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="auto", choices=["auto", "fp32", "exact", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-loss-fusion", action="store_true", help="evaluate the criterion with the stand-alone loss kernels")
    ap.add_argument("--bucket", action="store_true", help="SRGAN workload on one GPU: GradBuckets instead of ordinary autograd gradient accumulation")
    ap.add_argument("--no-weight-cache", action="store_true", help="re-pack every filter in front of its conv instead of one repack launch per step")
    ap.add_argument("--debug-flags", type=int, default=0, help="libsrb200 debug / A-B knobs (srb_debug_set_flags; see csrc/tc_conv_sl.cu)")
    ap.add_argument("--torch-optimizer", action="store_true", help="torch.optim's fused Adam instead of srb200.FlatAdam (ESPCN / EDSR)")
    ap.add_argument("--no-sub", action="store_true", help="skip the VDSR cfg3 sub-result (extra key of the default run)")
    ap.add_argument("--variants", default=None, help="--impl cudnn: comma list of as-is,tuned,tuned-cl,tuned-cl-graph,tuned-cl-bf16-graph")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[a.workload]
    config = make_config(a.workload, world, a.no_graph, a.math)

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(a.workload, a.steps, a.warmup, budget_s=120.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s",
                "n_gpus": a.gpus, "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    ctx.init_cuda()
    if a.debug_flags:
        import ctypes
        from srb200 import _lib as _l
        _l.lib.srb_debug_set_flags.argtypes = [ctypes.c_int]
        _l.lib.srb_debug_set_flags.restype = None
        _l.lib.srb_debug_set_flags(a.debug_flags)
    if os.environ.get("SRB_BENCH_DEBUG"):
        torch._C._set_print_stack_traces_on_fatal_signal(True)  # C++ backtrace on SIGSEGV/SIGABRT (debugging aid)
    if a.impl == "cudnn":
        variants = (a.variants.split(",") if a.variants else
                    ["as-is", "tuned-cl-graph"] + (["tuned-cl-bf16-graph"] if a.workload.startswith("edsr256") else []))
        runs = []
        for v in variants:
            try:
                runs.append(measure_cudnn(ctx, a, a.workload, v, a.steps, a.warmup))
            except Exception as e:  # report, keep going
                runs.append({"variant": v, "error": repr(e)[:300]})
        ok = [r for r in runs if "value" in r]
        best = max(ok, key=lambda r: r["value"])
        config.update({"launch": "stock torch.nn eager / cuda-graph per variant", "math": "cudnn (allow_tf32=%s)" % best["allow_tf32"]})
        line = {"impl": "cudnn", "metric": METRIC, "value": best["value"], "unit": "images/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if "bf16" in best["variant"] else "tf32",
                "data": "synthetic", "config": config, "clocks": best["clocks"], "e2e": best["e2e"], "gpu_launches": 0,
                "best_variant": best["variant"], "variants": runs}
        if rank == 0:
            print(json.dumps(line))
            sys.stdout.flush()
        finish(ctx)
        return 0

    if model_key == "srgan":
        res = measure_srgan(ctx, a, a.workload, a.steps, a.warmup)
    else:
        res = measure_srb(ctx, a, a.workload, a.steps, a.warmup, full=True)
    line = {"metric": METRIC, "value": res["value"], "unit": "images/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"auto": "tf32", "fp32": "f32", "exact": "3xtf32", "bf16": "bf16"}[a.math],
            "data": "synthetic", "config": config, "clocks": res["clocks"], "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "kernels": res["kernels"]}
    if res.get("comm"):
        line["comm"] = res["comm"]
    if res.get("comm_capture_error"):
        line["comm_capture_error"] = res["comm_capture_error"]
    if res.get("comm_fallback"):
        line["comm_fallback"] = res["comm_fallback"]
    if a.workload == DEFAULT_WORKLOAD and not a.no_sub and a.math == "auto":
        # BASELINE.json names VDSR cfg3 for the 1->8 GPU curve: a short device-timed sub-result rides along in the same line
        try:
            sub = measure_srb(ctx, a, "vdsr_b64_128", 12, 3, full=False)
            line["sub_results"] = [{"workload": sub["workload"], "value": sub["value"], "unit": "images/s",
                                    "ms_per_step": sub["ms_per_step"], "steps": sub["steps"], "n_gpus": world,
                                    "comm": sub.get("comm"), "clocks": sub["clocks"]}]
        except Exception as e:
            line["sub_results"] = [{"workload": "vdsr_b64_128", "error": repr(e)[:300]}]
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_reference_run(a.workload, 1000, 1, budget_s=15.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        sys.stdout.flush()
    finish(ctx)
    return 0


def finish(ctx):
    if ctx.dist is not None:
        # Teardown order matters: CUDA graphs that captured NCCL kernels must be gone before the communicator is, and a
        # communicator abort can block in this torch/NCCL build.  All ranks rendezvous, then leave without running the
        # interpreter's (and NCCL's) destructors -- every result has been printed and flushed by now.
        torch.cuda.synchronize()
        ctx.dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    sys.exit(main())
