"""Sustained time per 480-sample block of a live batch fed device rows (pvgpu_process_block_device), for A/B runs of the
resynthesis back end (PVGPU_FUSED=0/1) and block sizes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audiomod_b200 as A
from audiomod_b200.synth import synth
for S, B in ((4096, 480), (4096, 256), (4096, 960), (1024, 480)):
    x = np.ascontiguousarray(np.tile(synth(2, 44100, 2.0, 1), (S, 1)))
    d = torch.from_numpy(x).cuda()
    pv = A.phasevocoder(44100, 1, 1.0, 7.0, 0, 1, 2048, streams=S)
    st = torch.cuda.Stream()
    n = x.shape[1]
    calls = (n - B) // B
    with torch.cuda.stream(st):
        for i in range(20):
            pv.processBlockDevice(d.data_ptr() + 4 * i * B, n, B, st.cuda_stream)
        st.synchronize()
        t0 = time.perf_counter()
        for i in range(20, calls):
            pv.processBlockDevice(d.data_ptr() + 4 * i * B, n, B, st.cuda_stream)
        st.synchronize()
        total = time.perf_counter() - t0
    pv.close()
    per = 1e3 * total / (calls - 20)
    print(json.dumps({"fused_env": os.environ.get("PVGPU_FUSED"), "streams": S, "block": B, "ms_per_call": round(per, 4), "audio_s_per_s": round(S * B / 44100 / per * 1e3)}))
