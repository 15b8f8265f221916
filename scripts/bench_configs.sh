#!/bin/bash
# Throughput of the other BASELINE.json configurations (parity-test configurations, timed for the record), the optional
# cepstral mode and the generic-size kernels:  scripts/bench_configs.sh  (under gpurun, from the repo root)
# -> gpurun_out/r02_bench_<workload>.json and one markdown table row per workload on stdout
for w in cfg1 cfg2 cfg3-formant cfg3-gender cfg5-robotic-2048 cfg5-whisper-2048 cfg5-vocoder-2048 cfg5-robotic-512 cfg5-robotic-8192 cepstral-gender generic-256 generic-16384; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-latency --parity-rows 2 > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err || echo "FAILED $w"
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_$w.json").read().strip().splitlines()[-1])
p = d["parity"] or {}
dr = p.get("device_resident_f32", {})
print("| $w |", d["config"]["workload"].split(": ", 1)[1], "|", f'{round(d["value"]):,}', "|", round(d["ms_per_step"], 1), "|", f'{round(d["e2e"]["value"]):,}', "|", f'{round(d["e2e"]["f32"]["value"]):,}', "|",
      ", ".join(f"{k} {v:.1f}" for k, v in d["roofline"]["kernel_ms_per_step"].items()), "|",
      ("counts equal, %.1f dB, %.1e" % (dr.get("min_snr_db") or float("inf"), dr.get("max_abs", 0))) if dr else "-", "|")
PY
done
