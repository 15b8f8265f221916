"""Post-chain of FFT-free effects on the batch's output rows (pvgpu_batch_set_postchain; audiomod_b200/csrc/pv_post.cu) against
the unmodified reference's gain / compressor / limiter objects (oracle/_ref/fxref_drv) applied to the unmodified reference's
phase-vocoder output -- the chain an SDK user builds from those objects (README.md:78-94)."""
import numpy as np
import pytest

from cases import make_input
from test_gpu_parity import assert_parity

pytestmark = pytest.mark.gpu

CHAINS = [
    [("gain", 1.7)],
    [("compressor", -10.0, 6.0, 6.0, 10.0, 100.0)],
    [("limiter", -10.0, 6.0, 0.0, 100.0)],
    [("gain", 0.6), ("compressor", -20.0, 4.0, 3.0, 5.0, 50.0), ("limiter", -6.0, 2.0, 1.0, 80.0)],
    [("limiter", -3.0, 0.0, 0.5, 20.0), ("limiter", -12.0, 9.0, 0.0, 200.0)],
]


@pytest.fixture(scope="module")
def A(pvlib):
    import audiomod_b200
    if pvlib.pvgpu_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return audiomod_b200


@pytest.mark.parametrize("chain", CHAINS, ids=["gain", "compressor", "limiter", "all-three", "two-limiters"])
@pytest.mark.parametrize("ch,st,fpc", [(1, 7.0, 64), (2, -3.0, 8)])
def test_postchain_matches_reference_objects(A, oracle, chain, ch, st, fpc):
    if not (oracle.have_ref() and oracle.have_ref_fx()):
        pytest.skip("oracle/_ref binaries were not built (need /root/reference at build time)")
    sr = 44100
    xs = [make_input("x", sr, ch, 0.9 - 0.25 * i, 2700 + i) for i in range(3)]
    ref = [oracle.run_ref_fx(oracle.run_ref(x, sr, semitones=st), sr, chain) for x in xs]
    outs = {}
    for fused in (False, True):
        b = A.PhaseVocoderBatch(len(xs), xs[0].shape[1], sr, ch, 1.0, st)
        b.set_fused(fused)
        b.tune(frames_per_chunk=fpc)          # the effect state crosses chunk boundaries
        b.set_postchain(chain)
        ys = b.run(xs)
        b.close()
        for i, (y, r) in enumerate(zip(ys, ref)):
            assert_parity(y, r, f"{[c[0] for c in chain]} fused={fused} [{i}]")
        outs[fused] = ys
    for a, c in zip(outs[False], outs[True]):
        assert np.array_equal(a, c)


def test_postchain_time_sliced_host_rows_and_clear(A, oracle):
    """Equal-length streams in one 2-D host array take the time-sliced pipeline: the chain runs on every chunk's columns before
    their D2H copy.  Clearing the chain gives the plain output again."""
    if not (oracle.have_ref() and oracle.have_ref_fx()):
        pytest.skip("oracle/_ref binaries were not built")
    sr, S = 44100, 5
    xs = [make_input("x", sr, 1, 0.6, 2800 + i) for i in range(S)]
    X = np.ascontiguousarray(np.concatenate(xs, axis=0))
    chain = [("compressor", -15.0, 3.0, 4.0, 10.0, 100.0), ("limiter", -8.0, 3.0, 0.0, 100.0)]
    b = A.PhaseVocoderBatch(S, X.shape[1], sr, 1, 1.0, 7.0)
    b.tune(frames_per_chunk=16)
    n_out = int(b.plan(X.shape[1])[0])
    Y = np.zeros((S, n_out), np.float32)
    b.set_postchain(chain)
    b.run_host_rows([X[r] for r in range(S)], [Y[r] for r in range(S)])
    for i in range(S):
        assert_parity(Y[i:i + 1], oracle.run_ref_fx(oracle.run_ref(xs[i], sr, semitones=7.0), sr, chain), f"ts[{i}]")
    b.set_postchain([])
    b.run_host_rows([X[r] for r in range(S)], [Y[r] for r in range(S)])
    assert_parity(Y[2:3], oracle.run_ref(xs[2], sr, semitones=7.0), "cleared")
    from audiomod_b200 import _lib
    q = np.zeros((S, n_out), np.int16)
    b.set_postchain(chain)
    with pytest.raises(A.PvgpuError) as e:
        b.run_host_rows([X[r].astype(np.int16) for r in range(S)], [q[r] for r in range(S)], A.S16)
    assert e.value.code == _lib.EINVAL
    b.close()


EQ_PARAMS = [1, 120, 0.7, 1.0,  1, 300, 0.5, -3.0,  1, 900, 1.2, 4.0,  0, 2000, 0.3, 1.5,
             1, 3100, 2.0, -5.0,  1, 4500, 0.9, 2.5,  1, 7000, 0.6, 3.0,  1, 12000, 0.8, 1.0]


@pytest.mark.parametrize("spec", ["default", "all-bands", "sections"])
def test_equalizer_and_biquads_match_reference_objects(A, oracle, spec):
    """The 8-band equalizer object (src/equalizer/equalizer.cc) and single biquad sections of every type
    (src/common/filters/biquadfilter.cc), expanded into the post-chain, against the reference's own objects."""
    if not (oracle.have_ref() and oracle.have_ref_fx()):
        pytest.skip("oracle/_ref binaries were not built")
    sr, ch = 44100, 2
    if spec == "default":
        chain, ref_chain = A.equalizer_chain(), [("equalizer",)]
        assert [c[1] for c in chain] == [0]                       # only the 200 Hz high-pass is on by default
    elif spec == "all-bands":
        chain, ref_chain = A.equalizer_chain(EQ_PARAMS), [tuple(["equalizer"] + EQ_PARAMS)]
        assert [c[1] for c in chain] == [0, 1, 2, 2, 2, 4, 5]
    else:
        chain = [("biquad", t, 500.0 + 400.0 * t, 0.8, 3.0) for t in (3, 6, 7, 8)] + [("gain", 0.9)]
        ref_chain = chain
    xs = [make_input("x", sr, ch, 0.8 - 0.3 * i, 3100 + i) for i in range(2)]
    ref = [oracle.run_ref_fx(oracle.run_ref(x, sr, semitones=4.0), sr, ref_chain) for x in xs]
    b = A.PhaseVocoderBatch(len(xs), xs[0].shape[1], sr, ch, 1.0, 4.0)
    b.tune(frames_per_chunk=16)
    b.set_postchain(chain)
    ys = b.run(xs)
    b.close()
    for i, (y, r) in enumerate(zip(ys, ref)):
        assert_parity(y, r, f"{spec}[{i}]")
