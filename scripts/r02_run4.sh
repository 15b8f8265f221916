#!/bin/bash
# after the gather-batching fix: tests, fused / split / bulk A/B, ncu of the fused path
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/r02d_pytest.log 2>&1
echo "== full suite: $(tail -1 $out/r02d_pytest.log)"
grep -E "^(FAILED|ERROR)" $out/r02d_pytest.log | head -20
PVGPU_ANALYSE_BULK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or ragged or device_resident" > $out/r02d_pytest_bulk.log 2>&1
echo "== bulk-copy analysis, parity subset: $(tail -1 $out/r02d_pytest_bulk.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for v in "default:" "split:PVGPU_FUSED=0" "bulk:PVGPU_ANALYSE_BULK=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02d_bench_$tag.log 2> $out/r02d_bench_$tag.err
  python - "$out/r02d_bench_$tag.log" "$tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
done
small="--streams 888 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency --no-parity"
python bench.py $small > $out/r02d_bench_888.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_synth -s 60 -c 1 -o $out/r02d_ncu_fused python bench.py $small > $out/r02d_ncu2.log 2>&1
ncu -i $out/r02d_ncu_fused.ncu-rep --page raw --csv > $out/r02d_ncu_fused_raw.csv 2>/dev/null
ls -la $out | grep r02d_ncu
