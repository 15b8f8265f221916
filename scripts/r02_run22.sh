#!/bin/bash
# eight GPUs, final build: the torchrun bench line only
out=gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $NG --steps 3 --warmup 3 > $out/r02_bench_${NG}gpu_final.log 2> $out/r02_bench_${NG}gpu_final.err; echo "== bench $NG gpus rc $?"
python - "$out/r02_bench_${NG}gpu_final.log" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("== N", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e s16", round(d["e2e"]["value"]), "f32", round(d["e2e"]["f32"]["value"]),
              "strong", d["strong_scaling"] and (round(d["strong_scaling"]["value"]), round(d["strong_scaling"].get("e2e", {}).get("value", 0))), "parity", d["parity"] and d["parity"]["device_resident_f32"], d["clocks"])
PY
tail -2 $out/r02_bench_${NG}gpu_final.err
