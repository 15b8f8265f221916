"""The optional cepstral-envelope modes (PVGPU_GENDER_CEPSTRAL = 8, PVGPU_FORMANT_CEPSTRAL = 9; audiomod_b200/csrc/pv_cepstral.cu)
against the reference built with its commented-out formantShiftSlice calls switched on (oracle/_ref/pvref_drv_cep, see
oracle/Makefile; phasevocoderprocess.cc:824-840, 925-999).  Same bars as the parity modes: counts exact, >= 90 dB, <= 1e-4."""
import numpy as np
import pytest

from cases import make_input
from test_gpu_parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A(pvlib):
    import audiomod_b200
    if pvlib.pvgpu_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return audiomod_b200


@pytest.mark.parametrize("mode,semitones,ch,fft,sr", [
    (8, 4.0, 1, 2048, 44100),     # male -> female: envelope warp 0.85
    (8, -4.0, 1, 2048, 44100),    # female -> male: 1.17
    (8, 7.0, 2, 1024, 44100),
    (8, -3.0, 2, 4096, 48000),
    (9, 4.0, 1, 2048, 44100),     # formant "preservation": identity warp (whiten and re-colour with the same envelope)
    (9, -5.0, 2, 512, 22050),
    (8, 5.0, 1, 8192, 44100),
])
def test_cepstral_modes_match_patched_reference(A, oracle, mode, semitones, ch, fft, sr):
    if not oracle.have_ref_cepstral():
        pytest.skip("oracle/_ref/pvref_drv_cep was not built (needs /root/reference at build time)")
    xs = [make_input("x", sr, ch, 1.0 - 0.35 * i, 2500 + i) for i in range(2)]
    ref = [oracle.run_ref(x, sr, semitones=semitones, mode=mode - 7, fftsize=fft, cepstral=True) for x in xs]
    for fused in (False, True):
        b = A.PhaseVocoderBatch(len(xs), xs[0].shape[1], sr, ch, 1.0, semitones, mode, 1, fft)
        b.set_fused(fused)
        ys = b.run(xs)
        kt = b.stats()
        b.close()
        assert kt["kernel_launches"] > 0
        for i, (y, r) in enumerate(zip(ys, ref)):
            assert_parity(y, r, f"cepstral mode {mode} {semitones:+g} st fft {fft} fused={fused} [{i}]")
    # the streaming instance
    from test_gpu_parity import _cli_protocol
    pv = A.phasevocoder(sr, ch, 1.0, semitones, mode, 1, fft)
    y = _cli_protocol(pv, xs[1], sr, mode)
    pv.close()
    assert_parity(y, ref[1], f"cepstral mode {mode} stream")


def test_cepstral_gender_differs_from_bin_warp(A):
    """Mode 8 is not mode 1: the envelope warp changes the timbre differently from the nearest-bin warp."""
    sr = 44100
    x = make_input("x", sr, 1, 0.5, 2600)
    outs = []
    for mode in (1, 8):
        b = A.PhaseVocoderBatch(1, x.shape[1], sr, 1, 1.0, 4.0, mode, 1, 2048)
        outs.append(b.run([x])[0])
        b.close()
    assert outs[0].shape == outs[1].shape and np.max(np.abs(outs[0] - outs[1])) > 1e-2


def test_cepstral_modes_need_a_tiled_fft_size(A):
    from audiomod_b200 import _lib
    with pytest.raises(A.PvgpuError) as e:
        A.PhaseVocoderBatch(1, 1000, 44100, 1, 1.0, 4.0, 8, 1, 256)
    assert e.value.code == _lib.EINVAL
