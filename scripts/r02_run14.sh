#!/bin/bash
# row-group / chunk sizes small enough for the frame ring to stay in L2 between synthesis and overlap-add
out=gpurun_out
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for v in "0 0 0" "128 64 4" "256 64 4" "512 64 3" "256 32 4" "1024 16 3" "1024 32 3"; do
  set -- $v
  timeout 600 python bench.py $short --rows-per-group $1 --frames-per-chunk $2 --contexts $3 > $out/r02y_bench.log 2> $out/r02y_bench.err
  python - "$out/r02y_bench.log" "rows_per_group=$1 frames_per_chunk=$2 contexts=$3" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
  tail -1 $out/r02y_bench.err
done
