"""Host side of the product on CPU: the data-independent slice schedule against the oracle's counts, the derived
sizes, the C-ABI surface, and the loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIGS = [
    (dict(semitones=4, mode=0, fftsize=2048), 44100, 2),
    (dict(timeratio=1.5, mode=5, fftsize=4096), 48000, 2),
    (dict(semitones=7, mode=0, fftsize=2048), 44100, 1),
    (dict(semitones=-4, mode=2, fftsize=2048), 44100, 1),
    (dict(semitones=0, mode=1, fftsize=2048), 44100, 1),
    (dict(semitones=12, mode=0, coremode=2, fftsize=2048), 44100, 1),
    (dict(semitones=-12, mode=0, fftsize=1024), 22050, 1),
    (dict(timeratio=0.7, mode=5, fftsize=1024), 44100, 1),
    (dict(timeratio=2.0, mode=5, fftsize=512), 44100, 1),
    (dict(mode=6, fftsize=8192), 44100, 2),
    (dict(mode=3, fftsize=512), 44100, 2),
    (dict(mode=-1, semitones=3, fftsize=1024), 44100, 1),
    (dict(semitones=5, timeratio=1.3, mode=0, fftsize=2048), 44100, 1),
    (dict(semitones=7, mode=0, fftsize=1000), 44100, 1),   # non power of two -> rounded up like the reference
]


@pytest.mark.parametrize("kw,sr,ch", CONFIGS)
def test_plan_counts_match_oracle(pvlib, oracle, kw, sr, ch):
    import audiomod_b200 as A
    for n in (0, 1, 479, 480, 5000, 44100, 100003):
        x = np.zeros((ch, n), np.float32)
        y, slices = oracle.run_offline(x, sr, return_slices=True, **kw)
        c = A.plan_counts(n, sr, ch, kw.get("timeratio", 1.0), kw.get("semitones", 0.0), kw.get("mode", 0), kw.get("coremode", 1),
                          kw.get("fftsize", 2048))
        assert (c["n_out"], c["n_slices"], c["n_dropped"]) == (y.shape[1], slices, 0), (kw, n)


def test_survey_probe_counts(pvlib):
    """Sample counts the survey measured on the unmodified reference."""
    import audiomod_b200 as A
    assert A.plan_counts(441000, 44100, 1, 1.0, 7.0)["n_out"] == 441000
    c = A.plan_counts(472320, 48000, 2, 1.5, 0.0, A.NORMAL_STRETCH, 1, 4096)
    assert c["n_out"] > 0 and c["n_dropped"] == 0


def test_describe_matches_oracle_sizes(pvlib, oracle):
    import audiomod_b200 as A
    for kw, sr, ch in CONFIGS:
        d = A.describe(sr, ch, kw.get("timeratio", 1.0), kw.get("semitones", 0.0), kw.get("mode", 0), kw.get("coremode", 1),
                       kw.get("fftsize", 2048))
        st = oracle.OracleStream(sr, ch, kw.get("timeratio", 1.0), kw.get("semitones", 0.0), kw.get("mode", 0), kw.get("coremode", 1),
                                 kw.get("fftsize", 2048))
        assert d["hop"] == st.hop and d["fftsize"] == oracle.lib().pvo_fftsize(st.h)
        assert d["pitch_scale"] == oracle.lib().pvo_pitch_scale(st.h)
        st.close()
    # values the survey read from the reference's own log lines
    assert A.describe(44100, 2, 1.0, 4.0)["hop"] == 203
    assert A.describe(48000, 2, 1.5, 0.0, A.NORMAL_STRETCH, 1, 4096)["hop"] == 341
    assert A.describe(44100, 1, 1.0, 7.0)["hop"] == 170
    assert A.describe(44100, 1, 1.0, 7.0)["resampler_filt_len"] == 96
    assert A.describe(44100, 1, 1.0, 4.0)["resampler_filt_len"] == 80


def test_resampler_table_matches_oracle(pvlib, oracle):
    """Same fraction and filter length as the oracle's Speex restatement."""
    import audiomod_b200 as A
    for st in (4.0, 7.0, 12.0, -4.0, -12.0, 0.5):
        d = A.describe(44100, 1, 1.0, st)
        r = oracle.resampler_params(d["pitch_scale"])
        assert (d["resampler_num"], d["resampler_den"], d["resampler_filt_len"]) == (r["num"], r["den"], r["filt_len"])


def test_c_abi_exports_every_declared_symbol(pvlib):
    hdr = open(os.path.join(ROOT, "include", "pvgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(pvgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    so = C.CDLL(os.path.join(ROOT, "audiomod_b200", "libpvgpu.so"))
    missing = [n for n in sorted(names) if not hasattr(so, n)]
    assert not missing, missing
    from audiomod_b200 import _lib
    assert set(_lib.SYMBOLS) == names


def test_invalid_arguments_return_codes(pvlib):
    import audiomod_b200 as A
    from audiomod_b200 import _lib
    with pytest.raises(A.PvgpuError) as e:
        A.describe(44100, 0, 1.0, 0.0)
    assert e.value.code == _lib.EINVAL
    with pytest.raises(A.PvgpuError):
        A.describe(44100, 1, 1.0, 0.0, fftsize=1 << 20)
    assert pvlib.pvgpu_version() >= 100


def test_no_gpu_fails_loudly(pvlib):
    """There is no CPU fallback: without a device every create call reports PVGPU_ECUDA."""
    import audiomod_b200 as A
    from audiomod_b200 import _lib
    if pvlib.pvgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(A.PvgpuError) as e:
        A.phasevocoder(44100, 1, 1.0, 7.0)
    assert e.value.code == _lib.ECUDA
    with pytest.raises(A.PvgpuError) as e:
        A.PhaseVocoderBatch(4, 1000, 44100, 1, 1.0, 7.0)
    assert e.value.code == _lib.ECUDA
    with pytest.raises(A.PvgpuError) as e:
        A.PhaseVocoderMultiBatch(4, 1000, 44100, 1, 1.0, 7.0)
    assert e.value.code == _lib.ECUDA
    with pytest.raises(A.PvgpuError) as e:
        A.run_wav_files([("/nonexistent/in.wav", "/nonexistent/out.wav")], 1.0, 7.0)
    assert e.value.code == _lib.ECUDA
    with pytest.raises(A.PvgpuError) as e:
        A.HostBuffer(1 << 20, 0)
    assert e.value.code == _lib.ECUDA


def test_extension_modes_derive_like_their_parity_twins(pvlib):
    """Modes 8 / 9 (cepstral gender / formant) share every size, hop and resampler setting with modes 1 / 2; they only swap the
    spectral-envelope routine.  Unknown modes stay invalid."""
    import audiomod_b200 as A
    for st in (4.0, -4.0, 7.0):
        for twin, ext in ((A.GENDER_CHANGE, A.GENDER_CEPSTRAL), (A.FORMANT_PRESERVE, A.FORMANT_CEPSTRAL)):
            assert A.describe(44100, 1, 1.0, st, twin, 1, 2048) == A.describe(44100, 1, 1.0, st, ext, 1, 2048)
            assert A.plan_counts(44100, 44100, 1, 1.0, st, twin, 1, 2048) == A.plan_counts(44100, 44100, 1, 1.0, st, ext, 1, 2048)
    assert A.plan_counts(44100, 44100, 1, 1.0, 4.0, 10, 1, 2048)["n_out"] == 0


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "audiomod_b200")):
        if "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "pv_oracle" not in txt and "libpv_oracle" not in txt and "oracle/" not in txt.replace("the CPU oracle", ""), os.path.join(base, f)


def test_biquad_design_matches_the_cookbook(pvlib):
    """pvgpu_biquad_design restates biquadfilter::computeCoeffs (the RBJ audio-EQ-cookbook sections): checked against the
    closed forms in double precision, and structurally (unity DC gain of a low-pass, unity Nyquist gain of a high-pass)."""
    import math
    import audiomod_b200 as A
    sr = 44100
    for f0, q, g in ((200.0, 0.3, 1.0), (1000.0, 0.707, -6.0), (8000.0, 2.0, 4.5)):
        w = 2 * math.pi * f0 / sr
        al = math.sin(w) / 2 / q
        a = 10 ** (g / 40)
        want = {5: [(1 - math.cos(w)) / 2, 1 - math.cos(w), (1 - math.cos(w)) / 2, 1 + al, -2 * math.cos(w), 1 - al],
                0: [(1 + math.cos(w)) / 2, -(1 + math.cos(w)), (1 + math.cos(w)) / 2, 1 + al, -2 * math.cos(w), 1 - al],
                2: [1 + al * a, -2 * math.cos(w), 1 - al * a, 1 + al / a, -2 * math.cos(w), 1 - al / a],
                8: [1 - al, -2 * math.cos(w), 1 + al, 1 + al, -2 * math.cos(w), 1 - al]}
        for t, c in want.items():
            got = A.biquad_design(t, sr, f0, q, g)
            assert np.allclose(got, c, rtol=2e-6, atol=1e-7), (t, got, c)
        lp, hp = A.biquad_design(5, sr, f0, q, g), A.biquad_design(0, sr, f0, q, g)
        # float32 coefficients: the sums cancel down to 2(1 - cos w) ~ 8e-4 at 200 Hz, so a few 1e-4 relative is rounding
        assert abs(sum(lp[:3]) / sum(lp[3:]) - 1) < 1e-3
        assert abs((hp[0] - hp[1] + hp[2]) / (hp[3] - hp[4] + hp[5]) - 1) < 1e-3
    with pytest.raises(A.PvgpuError):
        A.biquad_design(9, sr, 100.0, 1.0, 0.0)
    assert len(A.equalizer_chain()) == 1 and len(A.equalizer_chain([1, 100, 1, 0] * 8)) == 8


def test_live_batch_host_containers(pvlib):
    """Ring FIFO (wrap-around, growth, partial pops), row-copy pool and non-temporal copies of the live batch: the library's own
    host-only self-test (pv_engine.cu: host_structs_selftest); needs no device."""
    assert pvlib.pvgpu_test_host_structs() == 0
