"""Developer helper: locate a host-heap corruption by forcing a large malloc after each step."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import audiomod_b200 as A
from audiomod_b200.synth import synth
from oracle import pv_oracle as O
libc = ctypes.CDLL(None)
libc.malloc.restype = ctypes.c_void_p
libc.malloc.argtypes = [ctypes.c_size_t]
libc.free.argtypes = [ctypes.c_void_p]
def probe(tag):
    ps = [libc.malloc(sz) for sz in (100, 5000, 130000, 100000, 3000000)]
    for p in ps: libc.free(p)
    print("heap ok after", tag, flush=True)
CASE = os.environ.get("CASE", "0")
sr, ch = (22050, 2) if CASE == "0" else (48000, 1)
KW = dict(semitones=5.0, mode=0, coremode=1, fftsize=256) if CASE == "0" else dict(timeratio=1.25, mode=5, coremode=1, fftsize=16384)
TR, ST, MODE, FFT = KW.get("timeratio", 1.0), KW.get("semitones", 0.0), KW["mode"], KW["fftsize"]
xs = [synth(800 + i, sr, 0.7 - 0.2 * i, ch) for i in range(2)]
probe("synth")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "oracle"):
    ref = [O.run_offline(x, sr, **KW) for x in xs]
    probe("oracle")
if which in ("all", "batch"):
    b = A.PhaseVocoderBatch(2, xs[0].shape[1], sr, ch, TR, ST, MODE, 1, FFT, 0)
    probe("batch create")
    ys = b.run(xs)
    probe("batch run")
    b.close()
    probe("batch close")
if which in ("all", "stream"):
    pv = A.phasevocoder(sr, ch, TR, ST, MODE, 1, FFT, 0)
    probe("stream create")
    x = xs[1]
    B, n, produced = 480, x.shape[1], 0
    for i in range(0, n, B):
        pv.processInData(x[:, i:i + B]); y = pv.getOutData(pv.getOutSamples()); produced += y.shape[1]
        if (i // B) % 5 == 0: probe(f"stream block {i//B}")
    pv.close()
    probe("stream close")
