"""Per-call latency of the streaming drop-in API (one audiomod::phasevocoder instance, the reference CLI's block protocol):
processBlock on blocks of max(480, sr/100) samples, wall time per call.  Usage on the GPU box: python scripts/stream_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import audiomod_b200 as A
from audiomod_b200.synth import synth

for name, ch, tr, st, mode, fft in (("cfg4 mono +7 st", 1, 1.0, 7.0, 0, 2048), ("cfg1 stereo +4 st", 2, 1.0, 4.0, 0, 2048),
                                     ("formant mono +4 st", 1, 1.0, 4.0, 2, 2048), ("robotic stereo", 2, 1.0, 0.0, 6, 2048)):
    sr = 48000 if fft == 4096 else 44100
    x = synth(1, sr, 4.0, ch)
    B = max(480, sr // 100)
    pv = A.phasevocoder(sr, ch, tr, st, mode, 1, fft)
    times = []
    for i in range(0, x.shape[1] - B, B):
        blk = np.ascontiguousarray(x[:, i:i + B])
        t0 = time.perf_counter()
        pv.processBlock(blk)
        times.append(time.perf_counter() - t0)
    pv.close()
    t = np.array(times[20:]) * 1e3
    print(f"{name}: block {B} samples = {1e3 * B / sr:.2f} ms of audio; per call median {np.median(t):.3f} ms, p99 {np.percentile(t, 99):.3f} ms, "
          f"max {t.max():.3f} ms; real-time factor {1e3 * B / sr / np.mean(t):.1f}x")
