"""Developer check (GPU): batch + streaming output vs the CPU oracle for a list of configurations."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pv_oracle as O
from audiomod_b200.synth import synth
import audiomod_b200 as A


def metrics(a, b):
    if a.shape != b.shape:
        return f"SHAPE {a.shape} vs {b.shape}"
    out = []
    for c in range(a.shape[0]):
        err = a[c].astype(np.float64) - b[c].astype(np.float64)
        p = np.sum(b[c].astype(np.float64) ** 2)
        e = np.sum(err ** 2)
        snr = 10 * np.log10(p / e) if e > 0 else float("inf")
        out.append(f"ch{c}: snr={snr:.1f}dB maxabs={np.max(np.abs(err)):.2e} bitexact={np.array_equal(a[c].view(np.uint32), b[c].view(np.uint32))}")
    return "; ".join(out)


CASES = [
    ("robotic", dict(mode=6, fftsize=2048), 44100, 2, 1.0, 5000),
    ("cfg4", dict(semitones=7, mode=0, fftsize=2048), 44100, 1, 2.0, 4000),
    ("cfg1", dict(semitones=4, mode=0, coremode=1, fftsize=2048), 44100, 2, 2.0, 1001),
    ("cfg2", dict(timeratio=1.5, mode=5, coremode=1, fftsize=4096), 48000, 2, 2.0, 1002),
    ("formant+4", dict(semitones=4, mode=2, fftsize=2048), 44100, 1, 2.0, 1003),
    ("formant-4", dict(semitones=-4, mode=2, fftsize=2048), 44100, 1, 2.0, 1003),
    ("gender+4", dict(semitones=4, mode=1, fftsize=2048), 44100, 1, 2.0, 1004),
    ("gender-4", dict(semitones=-4, mode=1, fftsize=2048), 44100, 1, 2.0, 1004),
    ("gender0", dict(semitones=0, mode=1, fftsize=2048), 44100, 1, 2.0, 1004),
    ("whisper1024", dict(mode=7, fftsize=1024), 44100, 2, 1.0, 5000),
    ("vocoder", dict(mode=3, fftsize=2048), 44100, 2, 1.0, 5001),
    ("chord512", dict(mode=4, fftsize=512), 44100, 2, 1.0, 5001),
    ("chord8192", dict(mode=4, fftsize=8192), 44100, 2, 1.0, 5001),
    ("core0", dict(semitones=3, mode=0, coremode=0, fftsize=2048), 44100, 2, 2.0, 5002),
    ("core2", dict(semitones=12, mode=0, coremode=2, fftsize=2048), 44100, 1, 2.0, 5003),
    ("constant", dict(mode=-1, fftsize=1024), 44100, 1, 1.0, 5004),
    ("robotic8192", dict(mode=6, fftsize=8192), 44100, 2, 1.0, 5000),
    ("robotic512", dict(mode=6, fftsize=512), 44100, 2, 1.0, 5000),
]

only = sys.argv[1:]
for name, kw, sr, ch, secs, seed in CASES:
    if only and name not in only:
        continue
    xs = [synth(seed + i, sr, secs * (1.0 - 0.13 * i), ch) for i in range(3)]
    ref = [O.run_offline(x, sr, **kw) for x in xs]
    t = time.time()
    b = A.PhaseVocoderBatch(len(xs), max(x.shape[1] for x in xs), sr, ch, kw.get("timeratio", 1.0), kw.get("semitones", 0.0),
                            kw.get("mode", 0), kw.get("coremode", 1), kw.get("fftsize", 2048))
    try:
        outs = b.run(xs)
        dt = time.time() - t
        for i in range(len(xs)):
            print(f"{name}[{i}] batch: {metrics(outs[i], ref[i])}  ({dt:.2f}s)", flush=True)
    except Exception as e:
        print(f"{name} batch FAILED: {e}", flush=True)
    b.close()
    # streaming instance with the CLI block protocol
    try:
        x = xs[0]
        pv = A.phasevocoder(sr, ch, kw.get("timeratio", 1.0), kw.get("semitones", 0.0), kw.get("mode", 0), kw.get("coremode", 1),
                            kw.get("fftsize", 2048))
        B = max(480, sr // 100)
        chunks = []
        n = x.shape[1]
        produced = 0
        for i in range(0, n, B):
            pv.processInData(x[:, i:i + B])
            y = pv.getOutData(pv.getOutSamples())
            chunks.append(y.copy()); produced += y.shape[1]
        if kw.get("mode", 0) != 5:
            z = np.zeros((ch, B), np.float32)
            while produced < n:
                pv.processInData(z)
                y = pv.getOutData(pv.getOutSamples())
                if n - produced <= y.shape[1]:
                    y = y[:, :n - produced]
                chunks.append(y.copy()); produced += y.shape[1]
        ys = np.concatenate(chunks, axis=1)
        print(f"{name} stream: {metrics(ys, ref[0])}", flush=True)
        pv.close()
    except Exception as e:
        print(f"{name} stream FAILED: {e}", flush=True)
