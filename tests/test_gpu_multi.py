"""In-library multi-GPU batch (pvgpu_mbatch_*) and NUMA-placed host buffers (pvgpu_host_*), through the C ABI.

The sharded run must be bit-identical to the single-device batch: a stream's result does not depend on which device or
which neighbours it had (no collective).  With one visible GPU the same code runs with one device thread; the
two-device case needs `gpurun --gpus 2` (profiles/ holds the log of that run).
"""
import numpy as np
import pytest

from cases import make_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A(pvlib):
    import audiomod_b200
    if pvlib.pvgpu_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return audiomod_b200


def _single(A, xs, sr, ch, st, device=0):
    b = A.PhaseVocoderBatch(len(xs), max(x.shape[1] for x in xs), sr, ch, 1.0, st, device=device)
    ys = b.run(xs)
    b.close()
    return ys


@pytest.mark.parametrize("ragged", [False, True], ids=["equal", "ragged"])
def test_mbatch_matches_single_device_bitwise(A, pvlib, oracle, ragged):
    sr, ch, S = 44100, 2, 11
    nd = min(pvlib.pvgpu_device_count(), 4)
    xs = [make_input("x", sr, ch, (0.25 + 0.07 * (i % 5)) if ragged else 0.4, 1300 + i) for i in range(S)]
    want = _single(A, xs, sr, ch, 4.0)
    m = A.PhaseVocoderMultiBatch(S, max(x.shape[1] for x in xs), sr, ch, 1.0, 4.0, devices=list(range(nd)))
    got = m.run(xs)
    owner = m.owner()
    st = m.stats()
    m.close()
    assert st["devices_used"] == nd and st["kernel_launches"] > 0
    assert sorted(set(owner.tolist())) == list(range(nd))
    assert np.array_equal(owner, A.shard_streams([x.shape[1] for x in xs], nd))
    for i in range(S):
        assert np.array_equal(got[i], want[i]), f"stream {i} (device {owner[i]}) differs from the single-device run"
    # and the single-device run is the reference's result
    ref = oracle.run_offline(xs[3], sr, semitones=4.0)
    err = np.max(np.abs(got[3].astype(np.float64) - ref))
    assert got[3].shape == ref.shape and err <= 1e-4


def test_mbatch_more_devices_than_streams(A, pvlib):
    sr = 44100
    nd = pvlib.pvgpu_device_count()
    xs = [make_input("x", sr, 1, 0.3, 1400)]
    want = _single(A, xs, sr, 1, 7.0)
    m = A.PhaseVocoderMultiBatch(1, xs[0].shape[1], sr, 1, 1.0, 7.0)      # devices=None: every visible device
    got = m.run(xs)
    assert m.stats()["devices_used"] == 1
    m.close()
    assert np.array_equal(got[0], want[0])
    assert nd >= 1


def test_mbatch_every_device_gives_the_same_bits(A, pvlib):
    """Each visible GPU on its own produces the same samples (tables are built per device from the same host values)."""
    sr = 44100
    xs = [make_input("x", sr, 1, 0.3, 1500 + i) for i in range(2)]
    want = _single(A, xs, sr, 1, 7.0, device=0)
    for d in range(1, pvlib.pvgpu_device_count()):
        got = _single(A, xs, sr, 1, 7.0, device=d)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), f"device {d}"


def test_mbatch_rejects_bad_device_lists(A, pvlib):
    from audiomod_b200 import _lib
    with pytest.raises(A.PvgpuError) as e:
        A.PhaseVocoderMultiBatch(2, 1000, 44100, 1, 1.0, 7.0, devices=[0, 0])
    assert e.value.code == _lib.EINVAL
    with pytest.raises(A.PvgpuError):
        A.PhaseVocoderMultiBatch(2, 1000, 44100, 1, 1.0, 7.0, devices=[pvlib.pvgpu_device_count()])


@pytest.mark.parametrize("huge", [False, True], ids=["thp", "hugetlb"])
def test_host_buffers_feed_the_batch(A, pvlib, oracle, huge):
    """pvgpu_host_alloc: page-locked, NUMA-placed I/O buffers; a batch run from them equals a run from ordinary numpy rows."""
    sr, S = 44100, 4
    xs = [make_input("x", sr, 1, 0.35, 1600 + i) for i in range(S)]
    n = xs[0].shape[1]
    want = _single(A, xs, sr, 1, 7.0)
    b = A.PhaseVocoderBatch(S, n, sr, 1, 1.0, 7.0)
    n_out = int(b.plan(n)[0])
    hin, hout = A.HostBuffer(4 * S * n, 0, True, huge), A.HostBuffer(4 * S * n_out, 0, True, huge)
    info = hin.info()
    assert info["bytes"] >= 4 * S * n and info["bytes"] % (2 << 20) == 0
    X, Y = hin.array(np.float32, (S, n)), hout.array(np.float32, (S, n_out))
    for i in range(S):
        X[i] = xs[i][0]
    Y[:] = 5.0
    b.run_host_rows([X[i] for i in range(S)], [Y[i] for i in range(S)])
    b.close()
    for i in range(S):
        assert np.array_equal(Y[i], want[i][0])
    del X, Y
    hin.close(); hout.close()
    assert pvlib.pvgpu_device_numa_node(0) >= -1
