#!/bin/bash
# frames per chunk with the final kernels (the default is 64)
out=gpurun_out
short="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for fpc in 64 32 48 96 128 64; do
  timeout 600 python bench.py $short --frames-per-chunk $fpc > $out/r02am_bench.log 2> $out/r02am_bench.err
  python - "$out/r02am_bench.log" "frames_per_chunk=$fpc" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
  tail -1 $out/r02am_bench.err
done
