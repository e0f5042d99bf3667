"""Concurrent pinned host->device bandwidth of all ranks of one node (torchrun), unbound vs bound to the GPU's NUMA node.
The 8-GPU e2e leg of bench.py is PCIe-limited; this says what the platform can deliver when all ranks copy at once."""
import os, sys, json
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
bench = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0]]
spec.loader.exec_module(bench)

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))


def probe(tag):
    host = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    devb = torch.empty_like(host, device="cuda")
    for _ in range(3):
        devb.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        devb.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = 20 * host.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device="cuda")
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        vals = [float(v) for v in allv]
    else:
        vals = [gbs]
    if rank == 0:
        print(json.dumps({"probe": tag, "per_rank_GBps": [round(v, 1) for v in vals], "sum_GBps": round(sum(vals), 1)}), flush=True)


probe("unbound (affinity: %d cpus)" % len(os.sched_getaffinity(0)))
rec = bench.bind_to_gpu_numa(lr)
probe("bound %s" % json.dumps(rec))
if world > 1:
    dist.destroy_process_group()
