#!/bin/bash
# packed FFMA2 resampler inner loop + whole-bucket bank-aware ordering, A/B against the previous build (ab/libpvgpu_base.so)
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py -q -x > $out/r02q_pytest.log 2>&1
echo "== parity + fused + fullsize tests (FFMA2): $(tail -1 $out/r02q_pytest.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency"
line() {
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1), "parity", (d.get("parity") or {}))
PY
}
for v in "ffma2:" "ffma2_win256:PVGPU_RS_WIN=256" "ffma2_fused:PVGPU_FUSED=1"; do
  tag=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py $short > $out/r02q_bench_$tag.log 2> $out/r02q_bench_$tag.err
  line $out/r02q_bench_$tag.log $tag; tail -2 $out/r02q_bench_$tag.err
done
cp audiomod_b200/libpvgpu.so /tmp/new.so && cp ab/libpvgpu_base.so audiomod_b200/libpvgpu.so
timeout 600 python bench.py $short > $out/r02q_bench_base.log 2> $out/r02q_bench_base.err
line $out/r02q_bench_base.log base; tail -2 $out/r02q_bench_base.err
cp /tmp/new.so audiomod_b200/libpvgpu.so
