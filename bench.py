#!/usr/bin/env python
"""bench.py -- SR training images/sec on B200 through the libsrb200 hot path.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                   (the reference's CPU PyTorch path, host cores)

Workload at N=1 (and per rank, weak scaling, at N>1): BASELINE.json configs[1] -- ESPCN x4, 64x64 LR,
batch 128, fwd + MSE + bwd + Adam (espcn.py:126-131), synthetic uniform inputs, weights from the
net's own N(0, 0.02) init under seed 0.  One "step" = one such optimizer step.

Prints ONE JSON line (rank 0).  value = images/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = same metric with the step's inputs copied from pinned host memory and the loss read back every step;
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
    "edsr256_x4_b32_lr32": ("edsr", (3, 256, 32), 32, (32, 32), "l1", "edsr"),  # cfg4 shapes, TF32 on fp32 storage (bf16: next round)
    "srcnn_x2_b16": ("srcnn", (3, 64), 16, (64, 64), "mse", "srcnn"),
}
DEFAULT_WORKLOAD = "espcn_x4_b128_lr64"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def out_shape(model_key, args, n, h, w):
    if model_key == "srcnn":
        return (n, 3, h - 16, w - 16)
    if model_key == "espcn":
        return (n, 3, (h - 8) * args[2], (w - 8) * args[2])
    if model_key == "vdsr":
        return (n, 3, h, w)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="srb200", choices=["srb200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="auto", choices=["auto", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model_key, margs, batch, (h, w), loss_kind, opt_key = WORKLOADS[a.workload]
    config = {"workload": a.workload, "net": model_key, "per_gpu_batch": batch, "global_batch": batch * world,
              "lr_hw": [h, w], "loss": loss_kind, "optimizer": opt_key, "parallelism": "dp%d" % world,
              "l2": "per-step working set (activations+grads ~1 GB) exceeds the 126 MB L2; inputs rotate over 3 batches",
              "launch": "eager" if a.no_graph else "cuda-graph replay (srb200.TrainStepGraphs: fwd+loss+bwd[+allreduce] graph per input slot, optimizer graph)"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(a.workload, a.steps, a.warmup, budget_s=120.0)
        line = {"impl": "reference", "metric": "SR training images/sec (device-timed)", "value": r["value"], "unit": "images/s",
                "n_gpus": a.gpus, "steps": r["steps"], "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import srb200
    from srb200 import _lib
    from srb200 import host

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl srb200) needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    srb200.set_math(a.math)

    torch.manual_seed(0)
    net = srb200.models.MODELS[model_key](*margs)
    host.init_model(model_key, net)
    net.to(dev).train()
    opt = host.make_optimizer(opt_key, net.parameters(), lr=1e-5, capturable=not a.no_graph)
    bucket = srb200.GradBucket(net, world_size=world)
    lossf = host.loss_for(model_key, fused=True)
    oshape = out_shape(model_key, margs, batch, h, w)

    gen = torch.Generator().manual_seed(1 + rank)
    host_x = [torch.rand((batch, 3, h, w), generator=gen).pin_memory() for _ in range(3)]
    host_t = [torch.rand(oshape, generator=gen).pin_memory() for _ in range(3)]
    dev_x = [t.to(dev) for t in host_x]
    dev_t = [t.to(dev) for t in host_t]
    loss_host = torch.zeros(1).pin_memory()

    def step(x, t):
        bucket.begin_step()
        y = net(x)
        loss = lossf(y, t)
        loss.backward()
        bucket.all_reduce()
        if model_key == "vdsr":
            torch.nn.utils.clip_grad_norm_(net.parameters(), host.VDSR_CLIP)
        opt.step()
        return loss

    # ---- whole-step CUDA graphs (srb200.TrainStepGraphs) -----------------------------------------------------------------
    # The small nets are host-launch bound when every kernel is launched from Python, so the step is captured once per
    # resident input slot: backward graph (zero non-direct grads + forward + loss + backward [+ NCCL all-reduce]) and an
    # optimizer graph (clipping + step).
    graphs = {}

    def capture_graphs():
        st = srb200.TrainStepGraphs(net, lossf, opt, bucket, slots=list(zip(dev_x, dev_t)),
                                    clip_norm=host.VDSR_CLIP if model_key == "vdsr" else None)
        graphs.update(stepper=st, launches=st.launches_per_step)

    def step_slot(i):
        """One training step on resident batch slot i (graph replay when captured, eager otherwise)."""
        if not graphs:
            return step(dev_x[i], dev_t[i])
        return graphs["stepper"].step(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        """H2D copy of step i's batch from pinned host memory on the copy stream (overlaps the previous step's kernels)."""
        with torch.cuda.stream(copy_stream):
            # into the resident slot (the graphs read fixed addresses); slot i%3 was last read by step i-3, long finished
            dev_x[i % 3].copy_(host_x[i % 3], non_blocking=True)
            dev_t[i % 3].copy_(host_t[i % 3], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def timed(nsteps, e2e):
        barrier()
        main = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            # every step's inputs cross PCIe inside the timed region (double buffered: batch i+1 is in flight while
            # step i computes) and every step's loss is read back by the host, like `loss.data[0]` in srcnn.py:134
            copy_stream.wait_event(e0)
            nxt = prefetch(0)
        for i in range(nsteps):
            if e2e:
                ev = nxt
                if i + 1 < nsteps:
                    nxt = prefetch(i + 1)
                main.wait_event(ev)
                loss = step_slot(i % 3)
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
                main.synchronize()
            else:
                step_slot(i % 3)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = tms.item()
        return ms

    for i in range(a.warmup):
        step(dev_x[i % 3], dev_t[i % 3])
    if not a.no_graph:
        capture_graphs()
        for i in range(3):
            step_slot(i)
    l0 = _lib.launch_count()
    with Clocks(local_rank) as ck:
        ms = timed(a.steps, e2e=False)
        launches = graphs["launches"] * a.steps if graphs else _lib.launch_count() - l0
        ms_e2e = timed(a.steps, e2e=True)
    clocks = ck.summary()

    # instrumented pass for the roofline of the dominant kernel
    pk = peaks()
    # The events bracket each C-ABI call on the launching stream.  The host is slower than the GPU on the small nets, so each
    # instrumented step first parks the GPU on a spin kernel long enough for the host to enqueue the whole step: the spans then
    # measure kernel execution, not launch latency.
    t0 = time.perf_counter()
    step(dev_x[0], dev_t[0])
    host_s = time.perf_counter() - t0
    torch.cuda.synchronize()
    spin_cycles = int(1.5 * min(host_s, 0.02) * (clocks.get("sm_mhz") or 1900) * 1e6) + 200000  # <= 30 ms even under a profiler
    with KernelTimer(_lib) as kt:
        barrier()
        for i in range(min(a.steps, 20)):
            torch.cuda._sleep(spin_cycles)
            step(dev_x[i % 3], dev_t[i % 3])
        tab = kt.table()
    total_ms = sum(v[0] for v in tab.values())
    (dname, dkey), (dms, dcnt) = max(tab.items(), key=lambda kv: kv[1][0])
    by, fl = algorithmic(dname, dkey)
    dur_s = dms / dcnt * 1e-3
    tf32_peak = pk["bf16_sustained"] / 2.0
    t_hbm, t_tc = by / (pk["hbm_gbs"] * 1e9), fl / (tf32_peak * 1e12)
    if t_hbm >= t_tc:
        roof = {"bound": "hbm", "achieved": by / dur_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": fl / dur_s / 1e12, "peak": tf32_peak, "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = None  # DRAM read+write bytes of this kernel from the committed `ncu --set full` capture, if it is in there
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        rec = tr.get(dname + "|" + "N%d Cin%d %dx%d Cout%d k%d s%d p%d ps%d" % dkey[:9])
        if rec:
            roof["traffic"] = rec["traffic"]
            roof["traffic_source"] = "profiles/ncu_traffic.json (%s, %.1f us under ncu)" % (rec["kernel"], rec["time_us"])
    except Exception:
        pass
    p_ = _lib.ConvParams(dkey[0], dkey[1], dkey[2], dkey[3], dkey[4], dkey[5], dkey[5], dkey[6], dkey[7], 0, dkey[9],
                         dkey[8], 0, 0.2, _lib.MATH_AUTO if a.math == "auto" else _lib.MATH_FP32)
    import ctypes
    roof.update({"kernel": dname, "layer": "N%d Cin%d %dx%d Cout%d k%d s%d p%d ps%d" % dkey[:9],
                 "tensor_path": bool(_lib.lib.srb_conv_uses_tensor_path(
                     ctypes.byref(p_), {"srb_conv_fprop": 0, "srb_conv_dgrad": 1, "srb_conv_wgrad": 2}.get(dname, 0), 1, 1)),
                 "us_per_launch": dur_s * 1e6, "share_of_kernel_time": dms / total_ms,
                 "peak_source": pk["source"] + (" (tf32 = bf16_sustained/2)" if roof["bound"] == "tensor" else ""),
                 "algorithmic_bytes": by, "algorithmic_flops": fl})
    kernels = sorted(((n, "N%d Cin%d %dx%d Cout%d k%d s%d p%d ps%d" % k[:9], v[0] / v[1] * 1e3, v[0] / total_ms)
                      for (n, k), v in tab.items()), key=lambda r: -r[3])

    imgs = batch * world * a.steps
    line = {"metric": "SR training images/sec (device-timed)", "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if a.math == "auto" else "f32",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(host_x[0].numel() + host_t[0].numel()) * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches), "roofline": roof,
            "kernels": [{"call": n, "layer": l, "us": round(us, 2), "share": round(sh, 4)} for n, l, us, sh in kernels]}
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            r = cpu_reference_run(a.workload, 1000, 1, budget_s=15.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # Teardown order matters: CUDA graphs that captured NCCL kernels must be gone before the communicator is, and a
        # communicator abort can block in this torch/NCCL build.  All ranks rendezvous, then leave without running the
        # interpreter's (and NCCL's) destructors -- every result has been printed and flushed by now.
        graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
