import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, audiomod_b200 as A
from audiomod_b200.synth import synth
B=480
for S in (1024, 4096):
    x = np.ascontiguousarray(np.tile(synth(2, 44100, 1.5, 1), (S, 1)))
    pv = A.phasevocoder(44100, 1, 1.0, 7.0, 0, 1, 2048, streams=S)
    tp=[]; 
    for i in range(0, x.shape[1]-B, B):
        blk = np.ascontiguousarray(x[:, i:i+B])
        t0=time.perf_counter(); pv.processBlock(blk); tp.append(time.perf_counter()-t0)
    pv.close()
    print(S, 'p50 ms', np.percentile(np.array(tp[20:])*1e3, 50))
