#!/usr/bin/env python
"""TF32 tensor peak by the MEASURED_PEAKS.json method (the driver measured bf16 only): torch.matmul fp32 inputs with
allow_tf32 on 8192^3, best of 10 (burst) and back to back for 4 s (sustained), CUDA events.  Library GEMM = the roofline
denominator, not a product path.  Writes gpurun_out/tf32_peak.json (copied to profiles/tf32_peak.json, read by bench.py)."""
import json
import os
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
flops = 2.0 * n ** 3
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
cnt = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(20):
        a @ b
    cnt += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / cnt
# same thing in bf16 as a cross-check against MEASURED_PEAKS.json
ah, bh = a.bfloat16(), b.bfloat16()
for _ in range(3):
    ah @ bh
torch.cuda.synchronize()
bb = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ah @ bh
    e1.record()
    torch.cuda.synchronize()
    bb = min(bb, e0.elapsed_time(e1))
out = {"tf32_tflops": flops / best / 1e9, "tf32_tflops_sustained": flops / sus / 1e9, "bf16_tflops_crosscheck": flops / bb / 1e9,
       "how": "torch.matmul fp32 with allow_tf32, 8192^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained), CUDA events",
       "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("TF32_PEAK_OUT", "gpurun_out/tf32_peak.json"), "w"), indent=1)
print(json.dumps(out))
