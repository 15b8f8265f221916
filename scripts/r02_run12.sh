#!/bin/bash
# OLA phase: 8 covering frames in flight + early window-sum loads; run length sweep
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -q -x > $out/r02s_pytest.log 2>&1
echo "== parity + fused tests: $(tail -1 $out/r02s_pytest.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
line() {
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
}
for v in "batch8:" "batch8_run32:PVGPU_OLA_RUN=32" "batch8_run8:PVGPU_OLA_RUN=8" $EXTRA; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02s_bench_$tag.log 2> $out/r02s_bench_$tag.err
  line $out/r02s_bench_$tag.log $tag; tail -2 $out/r02s_bench_$tag.err
done
