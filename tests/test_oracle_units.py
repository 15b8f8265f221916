"""Unit pins of the third-party arithmetic the path depends on (glibc 2.39, not in /root/reference)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_glibc_rand_restatement(oracle):
    """whisperSlice calls the unseeded libc rand() (phasevocoderprocess.cc:820); the restatement must reproduce the
    sequence of a fresh process."""
    code = r'''
#include <stdio.h>
#include <stdlib.h>
int main(void){ for(int i=0;i<2000;i++) printf("%d\n", rand()); return 0; }
'''
    d = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(d, exist_ok=True)
    src, exe = os.path.join(d, "rand_seq.c"), os.path.join(d, "rand_seq")
    with open(src, "w") as f:
        f.write(code)
    subprocess.check_call(["gcc", "-O1", "-o", exe, src])
    host = np.array([int(v) for v in subprocess.check_output([exe]).split()], dtype=np.int64)
    mine = np.zeros(2000, dtype=np.int32)
    import ctypes as C
    oracle.lib().pvo_rand_sequence(mine.ctypes.data_as(C.POINTER(C.c_int)), 2000)
    assert np.array_equal(host, mine.astype(np.int64))


def test_atan2f_restatement_bit_exact_on_host():
    """audiomod_b200/csrc/pv_math.cuh compiled for the host == this image's libm atan2f, bit for bit."""
    exe = os.path.join(ROOT, "oracle", "_ref", "host_atan2f_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "host_atan2f_check.cc")])
    out = subprocess.run([exe, "8000000", "3"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]


def test_princarg_range(oracle):
    a = np.linspace(-50, 50, 10001)
    p = np.array([oracle.lib().pvo_princarg(float(v)) for v in a])
    assert np.all(p > -np.pi - 1e-12) and np.all(p <= np.pi + 1e-12)
    assert np.allclose(np.exp(1j * p), np.exp(1j * a), atol=1e-9)


def test_forward_inverse_polar_roundtrip(oracle):
    """Hann-windowed forward + inverse of the restated KissFFT path reproduces window^2 * frame * N."""
    rng = np.random.default_rng(0)
    for n in (512, 1024, 2048, 4096, 8192):
        x = rng.standard_normal(n).astype(np.float32)
        mag, ph, _ = oracle.forward_polar(x)
        y = oracle.inverse_polar(mag, ph, n)
        w, _ = oracle.hann(n)
        assert np.allclose(y / n, x * w * w, atol=2e-5)


def test_squared_magnitude_peak_shortcut_is_exact():
    """k_lock_peaks decides the reference's strict `mag[b] > mag[n]` on squared magnitudes (audiomod_b200/csrc/pv_lock.cuh,
    lock_is_peak): q_b <= q_n -> false, q_b > q_n * (1 + 2^-21) and q_b > 1e-30 -> true, otherwise the exactly rounded
    square roots decide.  The two shortcuts must agree with comparing sqrtf values for every pair, in particular near ties."""
    import numpy as np
    rng = np.random.default_rng(11)
    n = 2_000_000
    qn = (rng.random(n).astype(np.float32) + np.float32(1e-3)) * np.float32(10.0) ** rng.integers(-28, 8, n).astype(np.float32)
    # neighbours a few ulps to a few 2^-21 apart, on both sides, plus unrelated pairs
    steps = rng.integers(-40, 41, n).astype(np.int32)
    qb = qn.copy()
    qb[: n // 2] = (qn[: n // 2].view(np.int32) + steps[: n // 2]).view(np.float32)
    qb[n // 2:] = qn[n // 2:] * (np.float32(1.0) + rng.random(n - n // 2).astype(np.float32) * np.float32(4e-6) - np.float32(1e-6)).astype(np.float32)
    qb = np.abs(qb).astype(np.float32)
    truth = np.sqrt(qb, dtype=np.float32) > np.sqrt(qn, dtype=np.float32)
    not_greater = ~(qb > qn)
    sure = (qb > qn * np.float32(1.00000048)) & (qb > np.float32(1e-30))
    assert not np.any(truth & not_greater), "q_b <= q_n must imply sqrtf(q_b) <= sqrtf(q_n)"
    assert np.all(truth[sure]), "q_b > q_n (1 + 2^-21) must imply sqrtf(q_b) > sqrtf(q_n)"
    assert np.any(~truth & ~not_greater), "the sample must contain rounded-square-root ties (the case the fallback exists for)"
