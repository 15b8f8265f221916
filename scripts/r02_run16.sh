#!/bin/bash
# live batch (pvgpu_create_multi) + device-resident input window of the streaming instance
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_live_batch.py tests/test_gpu_parity.py tests/test_gpu_dropin_cli.py -q -x --durations=8 > $out/r02aa_pytest.log 2>&1
echo "== live batch + streaming + drop-in tests: $(tail -1 $out/r02aa_pytest.log)"; grep -E "^(FAILED|ERROR)|Error" $out/r02aa_pytest.log | head -5; grep -A10 "slowest" $out/r02aa_pytest.log | head -12
timeout 600 python - > $out/r02aa_latency.json 2> $out/r02aa_latency.err <<'PY'
import json, bench
print(json.dumps(bench.stream_latency(), indent=1))
PY
echo "== latency rc $?"; cat $out/r02aa_latency.json; tail -3 $out/r02aa_latency.err
