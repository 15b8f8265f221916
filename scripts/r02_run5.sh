#!/bin/bash
# warp-specialised fused kernel: bitwise tests against the split kernels, then bench A/B and ncu
out=gpurun_out
PVGPU_FUSED_WS=1 timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fullsize.py -q > $out/r02f_pytest_ws.log 2>&1
echo "== WS fused vs split + full size: $(tail -1 $out/r02f_pytest_ws.log)"
grep -E "^(FAILED|ERROR)" $out/r02f_pytest_ws.log | head -20
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --no-parity"
for v in "ws:PVGPU_FUSED_WS=1" "fused:PVGPU_FUSED_WS=0" "split:PVGPU_FUSED=0"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02f_bench_$tag.log 2> $out/r02f_bench_$tag.err
  python - "$out/r02f_bench_$tag.log" "$tag" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
  tail -2 $out/r02f_bench_$tag.err
done
small="--streams 888 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency --no-parity"
PVGPU_FUSED_WS=1 python bench.py $small > $out/r02f_bench_888.log 2>&1 && \
PVGPU_FUSED_WS=1 ncu --set full --clock-control none --import-source on -k regex:k_synth -s 60 -c 1 -o $out/r02f_ncu_ws python bench.py $small > $out/r02f_ncu2.log 2>&1
ncu -i $out/r02f_ncu_ws.ncu-rep --page raw --csv > $out/r02f_ncu_ws_raw.csv 2>/dev/null
ls -la $out | grep r02f_ncu
