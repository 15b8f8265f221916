#!/bin/bash
# persistent warp-specialised overlap-add + resampler kernel (PVGPU_OLA_WS=1) against the default k_ola_resample
out=gpurun_out
if [ -z "$SKIPTEST" ]; then PVGPU_OLA_WS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x > $out/r02t_pytest_ws.log 2>&1; fi
echo "== parity + fullsize tests, PVGPU_OLA_WS=1: $(tail -1 $out/r02t_pytest_ws.log)"
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency $BENCHX"
line() {
  python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        par = (d.get("parity") or {}).get("device_resident_f32") or {}
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1), "snr", par.get("min_snr_db"), "checksum", d.get("result_checksum"))
PY
}
for v in $VARIANTS; do
  tag=${v%%:*}; envs=${v#*:}; envs=${envs//,/ }
  env $envs timeout 600 python bench.py $short > $out/r02t_bench_$tag.log 2> $out/r02t_bench_$tag.err
  line $out/r02t_bench_$tag.log $tag; tail -2 $out/r02t_bench_$tag.err
done
