#!/bin/bash
# source-level ncu capture of the two resynthesis kernels after the load-chain work (one launch each)
out=gpurun_out
small="--streams 592 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency --no-parity"
python bench.py $small > $out/r02az_bench_592.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_ola_resample|k_synthesise" -s 100 -c 2 -o $out/r02az_ncu python bench.py $small > $out/r02az_ncu.log 2>&1
ls -la $out/r02az_ncu.ncu-rep
