"""Generate tests/golden/pv_golden.npz from the UNMODIFIED reference (oracle/_ref/pvref_drv, built by oracle/Makefile
from /root/reference).  Run in the build container:  python tests/golden/make_golden.py

Stored per case: the int16 input PCM and the reference's float32 output (each one OS process, the reference CLI's
block protocol).  The reference has no golden vectors of its own, so these pin the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES, make_input  # noqa: E402
from oracle import pv_oracle as O  # noqa: E402


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    data = {}
    for name, kw, sr, ch, secs, seed in CASES:
        x = make_input(name, sr, ch, secs, seed)
        y = O.run_ref(x, sr, **kw)
        pcm = np.round(x.astype(np.float64) * 32768.0).astype(np.int16)
        assert np.array_equal((pcm.astype(np.float64) / 32768.0).astype(np.float32), x)
        data[name + "__in"] = pcm
        data[name + "__out"] = y.astype(np.float32)
        print(name, x.shape, "->", y.shape)
    # real-time (processBlock / outputReady) protocol of the SDK loop for two cases
    for name in ("cfg4_shift_p7_mono", "cfg5_robotic_512"):
        kw, sr, ch, secs, seed = next((c[1], c[2], c[3], c[4], c[5]) for c in CASES if c[0] == name)
        x = make_input(name, sr, ch, secs, seed)
        data[name + "__rt"] = O.run_ref(x, sr, protocol="rt", **kw).astype(np.float32)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pv_golden.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
