#!/bin/bash
# final bench line of round 2 (plain run) with the live-batch latency leg
out=gpurun_out
python bench.py --steps 5 --warmup 3 > $out/r02_final3_bench.log 2> $out/r02_final3_bench.err; echo "== bench rc $?"
tail -3 $out/r02_final3_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_final3_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("== value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "f32", round(d["e2e"]["f32"]["value"]), "launches", d["gpu_launches"], "cpu", round(d["cpu_baseline"]["value"]), d["clocks"], "issue_frac", round(d["roofline"]["issue_frac"], 3))
        print("== live", {k: (round(v["p50"], 3), round(v["p99"], 3)) for k, v in d["stream_latency_ms"]["live_batch_cfg4"].items()})
PY
