#!/bin/bash
# where did the split synthesis kernel's 22 ms go?  round-1 tree vs this tree on the same box, ncu of the split path
out=gpurun_out
short="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
(cd _old_r01 && timeout 600 python bench.py $short > ../$out/r02c_bench_oldtree.log 2> ../$out/r02c_bench_oldtree.err)
python - $out/r02c_bench_oldtree.log oldtree <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
PVGPU_FUSED=0 timeout 600 python bench.py $short --no-latency --no-parity > $out/r02c_bench_split.log 2>&1
python - $out/r02c_bench_split.log split <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); r = d["roofline"]
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()}, "serial", round(r["serialised_ms_per_step"], 1))
PY
small="--streams 592 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency --no-parity"
PVGPU_FUSED=0 python bench.py $small > $out/r02c_bench_592.log 2>&1 && \
PVGPU_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:k_synthesise -s 50 -c 1 -o $out/r02c_ncu_synth python bench.py $small > $out/r02c_ncu.log 2>&1
ncu -i $out/r02c_ncu_synth.ncu-rep --page raw --csv > $out/r02c_ncu_synth_raw.csv 2>/dev/null
ls -la $out | grep r02c
