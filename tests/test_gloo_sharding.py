"""world_size-2 gloo test of the host-side sharding / gather logic (the N > 1 path has no data-path collective)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from audiomod_b200.shard import run_sharded
    from audiomod_b200.synth import synth
    from oracle import pv_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xs = [synth(100 + i, 44100, 0.15 + 0.03 * i, 1) for i in range(5)]

    # stand-in for the per-GPU batch in this CPU test: the oracle (the GPU tests check the batch itself)
    def local(items):
        return [O.run_offline(x, 44100, semitones=7.0) for x in items]

    out = run_sharded(xs, local)
    if rank == 0:
        ref = local(xs)
        q.put(all(np.array_equal(a, b) for a, b in zip(out, ref)) and len(out) == len(ref))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_preserves_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_partitions_cover_and_balance():
    from audiomod_b200.shard import balanced_partition, block_partition
    for n, w in ((4096, 8), (4097, 8), (5, 8), (0, 2), (7, 2)):
        parts = [list(block_partition(n, w, r)) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    rng = np.random.default_rng(1)
    lens = rng.integers(1000, 500000, size=101)
    shards = balanced_partition(lens, 4)
    assert sorted(np.concatenate(shards).tolist()) == list(range(101))
    tot = [int(lens[s].sum()) for s in shards]
    assert max(tot) / min(tot) < 1.1


def test_library_partition_matches_python_rule(pvlib):
    """pvgpu_shard_streams (what pvgpu_mbatch uses on the device side) == the documented rule of shard.py."""
    from audiomod_b200.shard import balanced_partition, block_partition, library_partition
    rng = np.random.default_rng(2)
    for n, w in ((4096, 8), (4097, 8), (5, 8), (7, 2), (1, 1)):
        got = library_partition([441000] * n, w)
        assert [g.tolist() for g in got] == [list(block_partition(n, w, r)) for r in range(w)]
    for n, w in ((101, 4), (9, 8), (64, 3)):
        lens = rng.integers(1000, 500000, size=n)
        lens[::7] = lens[0]                      # ties keep the stable order
        got = library_partition(lens, w)
        want = balanced_partition(lens, w)
        assert [g.tolist() for g in got] == [x.tolist() for x in want]
