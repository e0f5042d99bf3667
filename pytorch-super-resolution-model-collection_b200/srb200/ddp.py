"""Data parallelism for the SR nets: one process per GPU, the batch sharded by rank, ONE flat fp32 gradient
buffer per model all-reduced over NCCL (NVLink 5 / NVSwitch) after backward.  The reference has no
multi-device code at all; this is the exchange step BASELINE.json asks for.

* Every parameter's .grad is a view into the flat buffer, so the collective is a single call on one
  tensor (ESPCN 149 KB ... EDSR-256 172 MB) and the optimizer reads the reduced values in place.
* Conv weights/biases are written into their slot directly by the wgrad kernel (pre-scaled by 1/world via
  srb200.set_grad_scale): the first write of a step overwrites, further writes in the same step (a module used twice
  before one backward, srgan.py:275-286) accumulate, slots nobody wrote are zeroed before the reduction --
  no zero_grad pass, no autograd accumulation kernel, no flatten copy.
* Everything else (PReLU slopes, BatchNorm, Linear) arrives through autograd and is copied into its slot.
torch.distributed is plumbing only (process group + the all_reduce call).
"""
import torch
import torch.distributed as dist

from . import functional as F


class GradBucket:
    def __init__(self, model, world_size=None, direct=True):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views = []
        self.direct_ids = set()
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            self.views.append(v)
        if direct and dev.type == "cuda":
            for m in model.modules():
                if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                    for p in (m.weight, m.bias):
                        if p is not None and p.requires_grad:
                            self.direct_ids.add(id(p))
        self._attach()
        F.set_grad_scale(1.0 / self.world)

    def _attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v
            p._srb_direct = id(p) in self.direct_ids
            p._srb_written = False

    def begin_step(self):
        """Replaces optimizer.zero_grad(): the first wgrad of the step overwrites a direct slot, later ones of the same
        step accumulate (`_srb_written`); the non-direct slots are cleared here."""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
            if id(p) in self.direct_ids:
                p._srb_written = False
            else:
                v.zero_()

    def finish_backward(self):
        """A direct parameter no wgrad reached in this step (a branch not taken) must read as zero gradient, not as the
        previous step's value.  Called by all_reduce(); call it yourself if you step without all_reduce()."""
        for p, v in zip(self.params, self.views):
            if id(p) in self.direct_ids and not getattr(p, "_srb_written", False):
                v.zero_()
                p._srb_written = True

    def all_reduce(self):
        """Sum over ranks (values are already scaled by 1/world where the kernels produced them)."""
        self.finish_backward()
        if self.world > 1:
            if len(self.direct_ids) < len(self.params):
                for p, v in zip(self.params, self.views):
                    if id(p) not in self.direct_ids:
                        v.mul_(1.0 / self.world)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat

    def detach(self):
        for p in self.params:
            p.grad = None
            p._srb_direct = False
        F.set_grad_scale(1.0)
