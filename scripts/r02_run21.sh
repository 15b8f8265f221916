#!/bin/bash
# two GPUs: in-library sharding tests and the torchrun bench line with the final build
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $out/r02aq_pytest_multi.log 2>&1; echo "== multi tests on $(nvidia-smi -L | wc -l) GPUs: $(tail -1 $out/r02aq_pytest_multi.log)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/r02aq_bench_2gpu.log 2> $out/r02aq_bench_2gpu.err; echo "== bench 2 gpus rc $?"
python - <<'PY'
import json
for l in open("gpurun_out/r02aq_bench_2gpu.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("== N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "f32", round(d["e2e"]["f32"]["value"]), "strong", d.get("strong_scaling"), "parity", (d.get("parity") or {}).get("device_resident_f32"))
PY
tail -2 $out/r02aq_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $out/r02aq_ref_2gpu.log 2>&1; echo "== reference arm under torchrun rc $?"; grep -c '"impl": "reference"' $out/r02aq_ref_2gpu.log
