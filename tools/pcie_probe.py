"""Concurrent host <-> device copy probe, one process per GPU (the ceiling of the end-to-end path at N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
    python tools/pcie_probe.py            # one GPU

Every rank copies the c2 step's buffers (691 MB in, 562 MB out) between pinned host memory and its GPU, all ranks at
once: H2D alone, D2H alone, both directions.  With --bind (default) a rank first binds itself to its GPU's CPU set
(nvmlDeviceSetCpuAffinity) so that its pinned pages are first touched on the GPU's own NUMA node.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
bind = "--no-bind" not in sys.argv
note = "unbound"
if bind:
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        note = "bound to %d cpus" % len(os.sched_getaffinity(0))
    except Exception as e:
        note = "not bound: %s" % e
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a = torch.empty(691_200_000, dtype=torch.uint8).pin_memory()
b = torch.empty(561_588_352, dtype=torch.uint8).pin_memory()
a.zero_(); b.zero_()
da = torch.empty_like(a, device="cuda")
db = torch.empty_like(b, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(h2d, d2h, K=5):
    sync()
    t0 = time.perf_counter()
    for _ in range(K):
        if h2d:
            with torch.cuda.stream(s1):
                da.copy_(a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                b.copy_(db, non_blocking=True)
    sync()
    t = (time.perf_counter() - t0) / K
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return t


for _ in range(2):
    run(True, True)
rows = [("H2D alone", run(True, False), a.numel()), ("D2H alone", run(False, True), b.numel()),
        ("both directions", run(True, True), a.numel() + b.numel())]
if rank == 0:
    print("N=%d ranks, %s" % (world, note))
    for name, t, nbytes in rows:
        print("  %-16s %8.2f ms per step   %7.1f GB/s per GPU   %7.1f GB/s aggregate" % (name, t * 1e3, nbytes / t / 1e9, world * nbytes / t / 1e9))
if world > 1:
    dist.destroy_process_group()
