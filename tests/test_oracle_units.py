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
