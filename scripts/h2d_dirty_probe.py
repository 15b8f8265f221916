"""Does a host-to-device copy run slower when the page-locked source was just written by the CPU (lines still in the caches)?
Times an 8 MB H2D with CUDA events: source untouched since the last copy / rewritten by one thread / by numpy from 8 threads."""
import json, time
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
n = 2 * 1024 * 1024
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
src = np.random.rand(n).astype(np.float32)
hv = h.numpy()
d = torch.empty(n, dtype=torch.float32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pool = ThreadPoolExecutor(8)
def copy_ms():
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
def rewrite1(): hv[:] = src
def rewrite8():
    list(pool.map(lambda i: np.copyto(hv[i * n // 8:(i + 1) * n // 8], src[i * n // 8:(i + 1) * n // 8]), range(8)))
for name, prep in (("untouched", lambda: None), ("rewritten by 1 thread", rewrite1), ("rewritten by 8 threads", rewrite8), ("rewritten, then 2 ms pause", lambda: (rewrite8(), time.sleep(0.002)))):
    ts = []
    for _ in range(30):
        prep(); ts.append(copy_ms())
    t = float(np.median(ts[5:]))
    print(json.dumps({"source": name, "h2d_ms_8MB": round(t, 3), "GBps": round(4 * n / t / 1e6, 1)}))
