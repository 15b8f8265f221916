#!/bin/bash
# persistent grids (PVGPU_PERSIST=k CTAs per SM per kernel) so that neighbouring pipeline stages share the SMs
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -q -x > $out/r02m_pytest.log 2>&1
echo "== parity + fused tests (default grids): $(tail -1 $out/r02m_pytest.log)"
PVGPU_PERSIST=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "golden or ragged or full_length or silence" > $out/r02m_pytest_persist.log 2>&1
echo "== parity subset, PVGPU_PERSIST=2: $(tail -1 $out/r02m_pytest_persist.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for v in "default:" "persist2:PVGPU_PERSIST=2" "persist3:PVGPU_PERSIST=3" "persist1:PVGPU_PERSIST=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02m_bench_$tag.log 2> $out/r02m_bench_$tag.err
  python - "$out/r02m_bench_$tag.log" "$tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
  tail -2 $out/r02m_bench_$tag.err
done
