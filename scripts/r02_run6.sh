#!/bin/bash
# full suite with the final back-end selection + cepstral modes, then memcheck on the small cases
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $out/r02h_pytest.log 2>&1
echo "== full suite: $(tail -1 $out/r02h_pytest.log)"
grep -E "^(FAILED|ERROR)" $out/r02h_pytest.log | head -20
python bench.py --workload cfg3-gender --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-latency --no-parity > $out/r02h_bench_gender.log 2>&1
bash scripts/r02_sanitize.sh memcheck
