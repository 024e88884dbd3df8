#!/usr/bin/env python
"""Host<->device copy ceiling of a box, one process per GPU (run under torchrun with N = 1, 2, 4, 8):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank copies 1 GiB pinned -> device, device -> pinned, and both at once, all ranks at the same time (barrier before
each timed loop); rank 0 prints per-rank and aggregate GB/s.  This is the ceiling the end-to-end (host buffer) numbers
of bench.py scale against: the e2e step moves 1.93 GB up and 1.43 GB down per scene."""
import json, os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world)
dev = torch.device("cuda", local)
N = 1 << 30
up = torch.empty(N, dtype=torch.uint8).pin_memory(); dn = torch.empty(N, dtype=torch.uint8).pin_memory()
d_up = torch.empty(N, dtype=torch.uint8, device=dev); d_dn = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def timed(fn, n=4):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    barrier()
    return dt

def h2d():
    with torch.cuda.stream(s1): d_up.copy_(up, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): dn.copy_(d_dn, non_blocking=True)
def both():
    h2d(); d2h()

res = torch.tensor([N / timed(h2d) / 1e9, N / timed(d2h) / 1e9, 2 * N / timed(both) / 1e9], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    rows = [[round(float(v), 1) for v in r.cpu()] for r in allr]
    agg = [round(sum(r[i] for r in rows), 1) for i in range(3)]
    numa = None
    try:
        numa = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except OSError:
        pass
    print(json.dumps({"n_gpus": world, "per_rank_GBps[h2d,d2h,both]": rows, "aggregate_GBps[h2d,d2h,both]": agg,
                      "host_cpus": len(os.sched_getaffinity(0)), "numa_nodes": numa}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
