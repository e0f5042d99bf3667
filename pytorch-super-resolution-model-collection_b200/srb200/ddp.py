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
* The exchange itself is libsrb200's own kernel when the ranks share an NVLink/NVSwitch box: the flat buffer lives in
  symmetric memory and srb_allreduce_inplace (csrc/comm.cu) reduces it in place over peer loads/stores, one launch per rank,
  CUDA-graph capturable, deterministic.  torch.distributed is plumbing only (process group, handle exchange); NCCL's
  all_reduce remains the fallback (SRB_ALLREDUCE=nccl forces it, CPU/gloo tests use it).
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import functional as F
from ._lib import lib, check


class GradBucket:
    def __init__(self, model, world_size=None, direct=True):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = world_size if world_size is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.symm = None       # symmetric-memory handle when the one-kernel NVLink all-reduce is in use
        self.comm = "none" if self.world == 1 else "nccl"
        self.flat = None
        if self.world > 1 and dev.type == "cuda" and os.environ.get("SRB_ALLREDUCE", "peer") != "nccl":
            self._try_symmetric(total, dev)
        if self.flat is None:
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

    def _try_symmetric(self, total, dev):
        """Put the flat buffer into symmetric memory (peer-mapped over NVLink) so that libsrb200's own kernel can reduce it in
        place (srb_allreduce_inplace).  torch only allocates and exchanges the handles; any failure leaves the NCCL path."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            n = (total + 3) // 4 * 4
            buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            if hdl.world_size != self.world or self.world > 16:
                raise RuntimeError("unexpected symmetric-memory group size")
            # device arrays of per-rank pointers to this tensor (allocation base + the tensor's offset) and to the signal pads
            off = int(getattr(hdl, "offset", 0))
            ptrs = [int(b) + off for b in hdl.buffer_ptrs]
            if ptrs[hdl.rank] != buf.data_ptr():
                raise RuntimeError("symmetric-memory pointer table does not match the local tensor")
            buf.zero_()
            self.flat = buf[:total]
            self._symm_buf = buf
            self.symm = hdl
            self._peer_bufs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
            self._peer_pads = torch.tensor([int(b) for b in hdl.signal_pad_ptrs], dtype=torch.int64, device=dev)
            self._state2 = torch.zeros(2, dtype=torch.int32, device=dev)
            self.comm = "peer"
            torch.cuda.synchronize(dev)
            dist.barrier()  # every rank's buffer is zeroed and mapped before the first exchange
        except Exception as e:  # no P2P / fabric support in this setup: NCCL
            self.symm = None
            self.flat = None
            self.comm_fallback = "%s: %s" % (type(e).__name__, e)

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
            if self.symm is not None:
                dev = self.flat.device
                check(lib.srb_allreduce_inplace(ctypes.c_void_p(self._peer_bufs.data_ptr()), ctypes.c_void_p(self._peer_pads.data_ptr()),
                                                self.symm.rank, self.world, self.flat.numel(), ctypes.c_void_p(self._state2.data_ptr()),
                                                ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat

    def detach(self):
        for p in self.params:
            p.grad = None
            p._srb_direct = False
        F.set_grad_scale(1.0)
