"""Developer helper: run one configuration through an AddressSanitizer build of the host code (see audiomod_b200/csrc/build_asan)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from audiomod_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "build_asan", "libpvgpu_asan.so")
import audiomod_b200 as A
from audiomod_b200.synth import synth
CASE = os.environ.get("CASE", "0")
sr, ch = {"0": (22050, 2), "1": (48000, 1), "2": (44100, 1)}[CASE]
TR, ST, MODE, FFT = {"0": (1.0, 5.0, 0, 256), "1": (1.25, 0.0, 5, 16384), "2": (1.0, -3.0, 2, 1000)}[CASE]
xs = [synth(800 + i, sr, 0.7 - 0.2 * i, ch) for i in range(2)]
b = A.PhaseVocoderBatch(2, xs[0].shape[1], sr, ch, TR, ST, MODE, 1, FFT, 0)
ys = b.run(xs)
b.close()
print("batch ok", ys[0].shape, flush=True)
pv = A.phasevocoder(sr, ch, TR, ST, MODE, 1, FFT, 0)
x = xs[1]
B, n, produced = 480, x.shape[1], 0
for i in range(0, n, B):
    pv.processInData(x[:, i:i + B]); y = pv.getOutData(pv.getOutSamples()); produced += y.shape[1]
z = np.zeros((ch, B), np.float32)
while MODE != 5 and produced < n:
    pv.processInData(z); y = pv.getOutData(pv.getOutSamples()); produced += y.shape[1]
pv.close()
print("stream ok", produced, flush=True)
