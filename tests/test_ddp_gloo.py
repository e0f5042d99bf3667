"""CPU suite, world_size 2 over gloo: the data-parallel exchange step (srb200.GradBucket).

The batch is sharded by rank, every rank back-propagates its shard, ONE all_reduce of the flat fp32 gradient
buffer follows, and the result must equal the full-batch gradient of a single process.  The nets
here are the CPU oracle's torch.nn restatements: the bucket is host logic (flat views, pre-scaling, one
collective), independent of which kernels produced the gradients -- the CUDA direct-write path is covered by the
-m gpu suite and by bench.py --gpus N.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as TF

import srb200
from oracle import torch_ref as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, args, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = R.build(name, args, seed=0)
        bucket = srb200.GradBucket(net, world_size=world)
        gen = torch.Generator().manual_seed(7)
        x = torch.rand((4, 3, 20, 20), generator=gen)
        with torch.no_grad():
            oshape = net(x[:1]).shape[1:]
        t = torch.rand((4,) + tuple(oshape), generator=gen)
        per = 4 // world
        xs, ts = x[rank * per:(rank + 1) * per], t[rank * per:(rank + 1) * per]
        for _ in range(2):  # second pass proves begin_step() clears what autograd accumulated
            bucket.begin_step()
            TF.mse_loss(net(xs), ts).backward()
            flat = bucket.all_reduce()
        # every parameter's .grad must still alias the flat buffer
        off = 0
        for p in bucket.params:
            assert p.grad.data_ptr() == flat[off:off + p.numel()].data_ptr()
            off += p.numel()
        q.put((rank, flat.clone().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,args", [("espcn", (3, 16, 2)), ("vdsr", (3, 8, 2)), ("fsrcnn", (3, 2, 8, 4, 1))])
def test_sharded_batch_allreduce_equals_full_batch_gradient(name, args):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth on the whole batch
    net = R.build(name, args, seed=0)
    gen = torch.Generator().manual_seed(7)
    x = torch.rand((4, 3, 20, 20), generator=gen)
    with torch.no_grad():
        oshape = net(x[:1]).shape[1:]
    t = torch.rand((4,) + tuple(oshape), generator=gen)
    TF.mse_loss(net(x), t).backward()
    truth = torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.requires_grad])
    assert torch.equal(torch.from_numpy(got[0]), torch.from_numpy(got[1]))  # ranks agree bit for bit
    err = (torch.from_numpy(got[0]) - truth).norm() / truth.norm()
    assert err < 1e-5, err


def test_bucket_layout_and_detach():
    net = R.build("espcn", (3, 8, 2), seed=0)
    b = srb200.GradBucket(net, world_size=1)
    assert b.flat.numel() == sum(p.numel() for p in net.parameters())
    assert all(p.grad is v for p, v in zip(b.params, b.views))
    assert not b.direct_ids  # direct wgrad writes are a CUDA-only contract
    b.detach()
    assert all(p.grad is None for p in net.parameters())
