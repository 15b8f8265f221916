"""Host<->device copy bandwidth probe (pinned memory): one and both directions.
Single process: python scripts/pcie_probe.py; all GPUs concurrently: torchrun --nproc-per-node N scripts/pcie_probe.py"""
import os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rows, cols = 4096, 441000
h_in = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
h_out = torch.empty((rows, cols), dtype=torch.float32, pin_memory=True)
h_in.fill_(1.0); h_out.fill_(0.0)
d_in = torch.empty((rows, cols), dtype=torch.float32, device="cuda")
d_out = torch.ones((rows, cols), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = rows * cols * 4 / 1e9


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=3):
    fn(); sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync_all()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


res = {}
for name, fn, vol in (("h2d", h2d, gb), ("d2h", d2h, gb), ("both", both, 2 * gb)):
    t = timed(fn)
    res[name] = vol / t
if world > 1:
    tt = torch.tensor([res[k] for k in res], device="cuda")
    dist.all_reduce(tt)
    if rank == 0:
        print(f"{world} GPUs concurrently, aggregate GB/s:", {k: round(float(v), 1) for k, v in zip(res, tt)})
    dist.destroy_process_group()
else:
    print("1 GPU, GB/s:", {k: round(v, 1) for k, v in res.items()})
