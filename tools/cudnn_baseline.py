#!/usr/bin/env python
"""GPU baseline the north_star names: the same nets on stock torch.nn (ATen -> cuDNN), eager, fp32 NCHW,
fwd + loss + bwd + optimizer, device-timed.  Measurement tool only (never imported by the product or bench.py).

Nets are restated with plain torch.nn layers following srcnn.py:13-25, espcn.py:13-25, vdsr.py:13-32,
edsr.py:13-45 of the reference (same layer shapes, init irrelevant for timing).  Two variants per net:
"as-is" (torch defaults: cudnn TF32 convs allowed, NCHW) and "tuned" (cudnn.benchmark + channels_last).
Prints one JSON line per (workload, variant).
"""
import argparse
import json
import torch
import torch.nn as nn
import torch.nn.functional as F


class ESPCN(nn.Module):
    def __init__(self, r=4):
        super().__init__()
        self.c1, self.c2, self.c3 = nn.Conv2d(3, 64, 5), nn.Conv2d(64, 32, 3), nn.Conv2d(32, 3 * r * r, 3)
        self.ps = nn.PixelShuffle(r)

    def forward(self, x):
        return self.ps(self.c3(F.relu(self.c2(F.relu(self.c1(x))))))


class SRCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1, self.c2, self.c3 = nn.Conv2d(3, 64, 9), nn.Conv2d(64, 32, 5), nn.Conv2d(32, 3, 5)

    def forward(self, x):
        return self.c3(F.relu(self.c2(F.relu(self.c1(x)))))


class VDSR(nn.Module):
    def __init__(self, F_=64, n=18):
        super().__init__()
        self.inp = nn.Conv2d(3, F_, 3, 1, 1, bias=False)
        self.body = nn.ModuleList([nn.Conv2d(F_, F_, 3, 1, 1, bias=False) for _ in range(n)])
        self.out = nn.Conv2d(F_, 3, 3, 1, 1, bias=False)

    def forward(self, x):
        h = F.relu(self.inp(x))
        for c in self.body:
            h = F.relu(c(h))
        return self.out(h) + x


class EDSR(nn.Module):
    def __init__(self, F_=64, n=16):
        super().__init__()
        self.inp = nn.Conv2d(3, F_, 3, 1, 1)
        self.b1 = nn.ModuleList([nn.Conv2d(F_, F_, 3, 1, 1) for _ in range(n)])
        self.b2 = nn.ModuleList([nn.Conv2d(F_, F_, 3, 1, 1) for _ in range(n)])
        self.mid = nn.Conv2d(F_, F_, 3, 1, 1)
        self.u1, self.u2 = nn.Conv2d(F_, 4 * F_, 3, 1, 1), nn.Conv2d(F_, 4 * F_, 3, 1, 1)
        self.out = nn.Conv2d(F_, 3, 3, 1, 1)

    def forward(self, x):
        h = self.inp(x)
        r = h
        for a, b in zip(self.b1, self.b2):
            h = h + b(F.relu(a(h)))
        h = self.mid(h) + r
        h = F.pixel_shuffle(self.u1(h), 2)
        h = F.pixel_shuffle(self.u2(h), 2)
        return self.out(h)


WORK = {
    "espcn_x4_b128_lr64": (lambda: ESPCN(4), 128, (64, 64), (224, 224), "mse", "adam"),
    "srcnn_x2_b16": (SRCNN, 16, (64, 64), (48, 48), "mse", "sgd"),
    "vdsr_b64_128": (lambda: VDSR(), 64, (128, 128), (128, 128), "mse", "vdsr"),
    "edsr64_x4_b32_lr32": (lambda: EDSR(64, 16), 32, (32, 32), (128, 128), "l1", "adam"),
    "edsr256_x4_b32_lr32": (lambda: EDSR(256, 32), 32, (32, 32), (128, 128), "l1", "adam"),
}


def run(name, variant, steps, warmup):
    mk, b, (h, w), (oh, ow), loss, optk = WORK[name]
    torch.backends.cudnn.benchmark = variant != "as-is"
    net = mk().cuda().train()
    x = [torch.rand(b, 3, h, w, device="cuda") for _ in range(3)]
    t = [torch.rand(b, 3, oh, ow, device="cuda") for _ in range(3)]
    if variant.startswith("tuned-cl"):
        net = net.to(memory_format=torch.channels_last)
        x = [v.contiguous(memory_format=torch.channels_last) for v in x]
    fused = {"fused": variant != "as-is"}  # the tuned variants get the same single-kernel optimizer as our arm
    graph = variant.endswith("-graph")     # ... and, like our arm, whole-step CUDA-graph replay
    if optk == "adam":
        opt = torch.optim.Adam(net.parameters(), lr=1e-5, capturable=graph, **fused)
    elif optk == "vdsr":
        opt = torch.optim.SGD(net.parameters(), lr=1e-5, momentum=0.9, weight_decay=1e-4, **fused)
    else:
        opt = torch.optim.SGD(net.parameters(), lr=1e-5, **fused)
    lf = F.l1_loss if loss == "l1" else F.mse_loss
    bf16 = "bf16" in variant  # autocast: what "the reference's cuDNN build in bf16" means for cfg4

    def fwd_loss(i):
        if bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return lf(net(x[i % 3]).float(), t[i % 3])
        return lf(net(x[i % 3]), t[i % 3])

    def step(i):
        opt.zero_grad(set_to_none=True)
        l = fwd_loss(i)
        l.backward()
        if optk == "vdsr":
            torch.nn.utils.clip_grad_norm_(net.parameters(), 0.4)
        opt.step()

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        pool = torch.cuda.graph_pool_handle()
        gs = []
        with torch.cuda.stream(side):
            for i in range(3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    opt.zero_grad(set_to_none=False)
                    fwd_loss(i).backward()
                    if optk == "vdsr":
                        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.4)
                    opt.step()
                gs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eager_step = step

        def step(i):  # noqa: F811
            gs[i % 3].replay()
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"impl": "torch-cudnn", "workload": name, "variant": variant, "ms_per_step": ms,
                      "images_per_s": b / ms * 1e3, "allow_tf32": torch.backends.cudnn.allow_tf32,
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="espcn_x4_b128_lr64,vdsr_b64_128,edsr64_x4_b32_lr32")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--variants", default="as-is,tuned,tuned-cl,tuned-cl-graph")
    a = ap.parse_args()
    for wl in a.workloads.split(","):
        for v in a.variants.split(","):
            try:
                run(wl, v, a.steps, a.warmup)
            except Exception as e:  # report, keep going
                print(json.dumps({"impl": "torch-cudnn", "workload": wl, "variant": v, "error": repr(e)[:200]}), flush=True)
