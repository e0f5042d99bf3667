"""Optimizer of the hot loop as one kernel: torch.optim.Adam (espcn.py:79,131; edsr.py:93,155) over FLAT buffers.

`GradBucket` already makes every parameter's .grad a view of one flat fp32 array; `FlatAdam` does the same for the parameters
and the two moments, so `optimizer.step()` is a single elementwise launch of libsrb200 (srb_adam_step_flat) instead of torch's
multi-tensor kernel (16 us for ESPCN's six tensors: one 512-thread block per tensor).  The step counter lives on the device, so
the call is CUDA-graph capturable and replays keep counting.  Same update rule, hyper-parameters and state as torch.optim.Adam
(amsgrad off); `state_dict()` is keyed like torch's (`exp_avg`, `exp_avg_sq`, `step` per parameter index)."""
import ctypes

import torch

from ._lib import lib, check


class FlatAdam:
    def __init__(self, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.bucket = bucket
        self.params = bucket.params
        dev = bucket.flat.device
        if dev.type != "cuda":
            raise RuntimeError("srb200.FlatAdam needs CUDA parameters; there is no CPU path")
        n = bucket.flat.numel()
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:  # re-seat every parameter as a view of the flat array (same values, same Parameter objects)
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            off += k
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._state = torch.zeros(2, dtype=torch.float32, device=dev)  # [step count, scratch]
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=self.params)]

    def zero_grad(self, set_to_none=False):
        self.bucket.begin_step()

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        flat_g = self.bucket.flat
        check(lib.srb_adam_step_flat(ctypes.c_void_p(self.flat_p.data_ptr()), ctypes.c_void_p(flat_g.data_ptr()),
                                     ctypes.c_void_p(self.exp_avg.data_ptr()), ctypes.c_void_p(self.exp_avg_sq.data_ptr()),
                                     self.flat_p.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                     float(g["weight_decay"]), ctypes.c_void_p(self._state.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream(self.flat_p.device).cuda_stream)))

    @property
    def step_count(self):
        return int(self._state[0].item())

    def state_dict(self):
        st, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            st[i] = {"step": self._state[0].clone(), "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                     "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
            off += k
        return {"state": st, "param_groups": [dict(self.defaults, params=list(range(len(self.params))))]}

    def load_state_dict(self, sd):
        off = 0
        for i, p in enumerate(self.params):
            k = p.numel()
            s = sd["state"][i]
            self.exp_avg[off:off + k].copy_(s["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + k].copy_(s["exp_avg_sq"].reshape(-1))
            self._state[0] = float(s["step"])
            off += k
        pg = sd["param_groups"][0]
        for key in ("lr", "betas", "eps", "weight_decay"):
            self.param_groups[0][key] = pg[key]
