"""Host<->device copy bandwidth probe: which host allocation lets an 8-GPU box feed all its GPUs at once?

    python scripts/pcie_probe.py                                   # one GPU
    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py     # all N concurrently (aggregate GB/s)

Allocation kinds: `torch` = torch pin_memory (cudaHostAlloc from an unpinned thread, pages wherever the kernel puts them),
`numa` = pvgpu_host_alloc on the GPU's own NUMA node (mbind + first touch from a thread bound to that node,
cudaHostRegister), `numa-huge` = the same on explicit 2 MB pages when the box has a hugetlb pool (else THP).
For every kind: H2D alone, D2H alone, both directions at once; first every rank alone in turn (per-GPU matrix), then all
ranks concurrently.  One JSON line per measurement on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audiomod_b200 as A  # noqa: E402
from audiomod_b200 import _lib  # noqa: E402

rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
GB = float(os.environ.get("PROBE_GB", "2.0"))
n = int(GB * 1e9 / 4)
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def alloc(kind):
    if kind == "torch":
        a, b = torch.empty(n, dtype=torch.float32, pin_memory=True), torch.empty(n, dtype=torch.float32, pin_memory=True)
        a.fill_(1.0); b.fill_(0.0)
        return a, b, None, {"numa_node": None, "hugepages": None}
    ha = A.HostBuffer(4 * n, rank, True, kind == "numa-huge")
    hb = A.HostBuffer(4 * n, rank, True, kind == "numa-huge")
    x, y = ha.array(np.float32, (n,)), hb.array(np.float32, (n,))
    x[:] = 1.0
    y[:] = 0.0
    return torch.from_numpy(x), torch.from_numpy(y), (ha, hb), ha.info()


def run(kind, which, h_in, h_out, reps=3):
    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    fn = {"h2d": h2d, "d2h": d2h, "both": lambda: (h2d(), d2h())}[which]
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (2 if which == "both" else 1) * GB * reps / (time.perf_counter() - t0)


node = _lib.lib().pvgpu_device_numa_node(rank)
rows = []
for kind in ("torch", "numa", "numa-huge"):
    h_in, h_out, keep, info = alloc(kind)
    sync_all()
    # every rank alone in turn
    for r in range(world):
        sync_all()
        if r == rank:
            for which in ("h2d", "d2h", "both"):
                rows.append({"alloc": kind, "mode": "alone", "gpu": rank, "gpu_numa_node": node, "buffer": info, "dir": which,
                             "gbs": round(run(kind, which, h_in, h_out), 1)})
        sync_all()
    # all ranks at once
    for which in ("h2d", "d2h", "both"):
        sync_all()
        v = run(kind, which, h_in, h_out)
        sync_all()
        if world > 1:
            t = torch.tensor([v], device="cuda", dtype=torch.float64)
            dist.all_reduce(t)
            agg = float(t.item())
        else:
            agg = v
        rows.append({"alloc": kind, "mode": f"all {world} concurrently", "gpu": rank, "dir": which, "gbs": round(v, 1), "aggregate_gbs": round(agg, 1)})
    del h_in, h_out
    if keep:
        for k in keep:
            k.close()
if world > 1:
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(rows, gathered, dst=0)
    if rank == 0:
        for rr in gathered:
            for row in rr:
                if row["mode"] == "alone" or row["gpu"] == 0:
                    print(json.dumps(row))
    dist.destroy_process_group()
else:
    for row in rows:
        print(json.dumps(row))
