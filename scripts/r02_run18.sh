#!/bin/bash
# host-computed cubic coefficients of the resampler: tests, then A/B against the previous build (ab/libpvgpu_base.so)
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py -q -x > $out/r02ah_pytest.log 2>&1
echo "== all gpu tests: $(tail -1 $out/r02ah_pytest.log)"; grep -E "^(FAILED|ERROR)" $out/r02ah_pytest.log | head -5
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency"
line() {
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        par = (d.get("parity") or {}).get("device_resident_f32") or {}
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1), "snr", par.get("min_snr_db"), "maxabs", par.get("max_abs"))
PY
}
timeout 600 python bench.py $short > $out/r02ah_bench_new.log 2> $out/r02ah_bench_new.err; line $out/r02ah_bench_new.log new; tail -1 $out/r02ah_bench_new.err
cp audiomod_b200/libpvgpu.so /tmp/new.so && cp ab/libpvgpu_base.so audiomod_b200/libpvgpu.so
timeout 600 python bench.py $short > $out/r02ah_bench_base.log 2> $out/r02ah_bench_base.err; line $out/r02ah_bench_base.log base; tail -1 $out/r02ah_bench_base.err
cp /tmp/new.so audiomod_b200/libpvgpu.so
timeout 600 python bench.py $short > $out/r02ah_bench_new2.log 2> $out/r02ah_bench_new2.err; line $out/r02ah_bench_new2.log new_again
