#!/bin/bash
# do the pipeline stages overlap better when every kernel leaves room on the SM for its neighbours?  (PVGPU_OCC_CAP experiment)
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_postchain.py -q -x > $out/r02l_pytest.log 2>&1
echo "== fused + postchain tests: $(tail -1 $out/r02l_pytest.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for v in "default:" "cap2:PVGPU_OCC_CAP=2" "cap3:PVGPU_OCC_CAP=3"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02l_bench_$tag.log 2> $out/r02l_bench_$tag.err
  python - "$out/r02l_bench_$tag.log" "$tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
  tail -2 $out/r02l_bench_$tag.err
done
