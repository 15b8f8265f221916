#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the bench (both arms)
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $out/r02n_pytest.log 2>&1
echo "== pytest -m gpu: $(tail -1 $out/r02n_pytest.log)"
grep -E "^(FAILED|ERROR)" $out/r02n_pytest.log | head
python -c "import __graft_entry__ as g; g.smoke()" > $out/r02n_smoke.log 2>&1; echo "== smoke rc $? $(tail -1 $out/r02n_smoke.log)"
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02n_bench_ref.log 2>&1; echo "== reference arm rc $?"; tail -c 400 $out/r02n_bench_ref.log
python bench.py > $out/r02n_bench.log 2> $out/r02n_bench.err; echo "== bench rc $?"
python - <<'PY'
import json
for l in open("gpurun_out/r02n_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("== value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "f32", round(d["e2e"]["f32"]["value"]), "launches", d["gpu_launches"], "cpu", round(d["cpu_baseline"]["value"]), d["clocks"])
PY
