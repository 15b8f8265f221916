#!/bin/bash
# post-chain + faster cepstral kernel: full suite, then the cepstral workload once more
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $out/r02k_pytest.log 2>&1
echo "== full suite: $(tail -1 $out/r02k_pytest.log)"
grep -E "^(FAILED|ERROR)" $out/r02k_pytest.log | head -20
python bench.py --workload cepstral-gender --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-latency --parity-rows 2 > $out/r02k_bench_cepstral.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r02k_bench_cepstral.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("== cepstral-gender value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, d["parity"]["device_resident_f32"])
PY
