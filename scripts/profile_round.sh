#!/bin/bash
# Round profile of the cfg4 hot path on one B200 (run under gpurun from the repo root):
#   scripts/profile_round.sh r01_final
# writes gpurun_out/<tag>_bench.log (plain run, the judged line), <tag>_launches.csv (ncu launch list, serialised and
# cold-cache: shares only), <tag>_ncu_full.ncu-rep + <tag>_ncu_full_raw.csv (ncu --set full, one launch of each kernel at
# 592 streams x 64 frames).  scripts/profile_digest.py turns them into profiles/<tag>_*.{csv,md} and profiles/traffic.json.
set -u
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench.log 2>&1 || exit 1
tail -c 600 $out/${tag}_bench.log
small="--streams 592 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-latency --no-parity"
python bench.py $small > $out/${tag}_bench_592.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 420 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $small > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 250 -c 5 -o $out/${tag}_ncu_full \
    python bench.py $small > $out/${tag}_ncu2.log 2>&1
ncu -i $out/${tag}_ncu_full.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw.csv 2>/dev/null
ls -la $out | grep ${tag}
