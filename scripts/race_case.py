"""Small batch (mono and stereo, phase-locked core) for compute-sanitizer racecheck / memcheck runs, plus a run-to-run
determinism check.  Usage on the GPU box: compute-sanitizer --tool racecheck python scripts/race_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import audiomod_b200 as A
from audiomod_b200.synth import synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
for ch, st in ((1, 7.0), (2, 4.0)):
    xs = [synth(100 + i, 44100, secs, ch) for i in range(3)]
    xs[1][:, 3000:9000] = 0.0
    outs = []
    for rep in range(3):
        b = A.PhaseVocoderBatch(len(xs), xs[0].shape[1], 44100, ch, 1.0, st, 0, 1, 2048)
        b.tune(frames_per_chunk=(64, 7, 64)[rep])
        outs.append(b.run(xs))
        b.close()
    same = all(np.array_equal(a, c) for o in outs[1:] for a, c in zip(outs[0], o))
    print("channels", ch, "deterministic across runs and chunk sizes:", same)
