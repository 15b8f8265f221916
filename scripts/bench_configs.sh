#!/bin/bash
# Throughput of the other BASELINE.json configurations (parity-test configurations, timed for the record):
#   scripts/bench_configs.sh   (under gpurun, from the repo root) -> gpurun_out/bench_<workload>.json
for w in cfg1 cfg2 cfg3-formant cfg3-gender cfg5-robotic-2048 cfg5-whisper-2048 cfg5-vocoder-2048 cfg5-robotic-512 cfg5-robotic-8192; do
  python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "FAILED $w"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w", round(d["value"]), round(d["ms_per_step"], 1), round(d["e2e"]["value"]), round(d["e2e"]["s16"]["value"]), {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
PY
done
