#!/bin/bash
# first GPU pass of round 2: new tests on the split kernels, fused kernel against them, then the bench both ways
out=gpurun_out
mkdir -p $out
rm -f $out/parity_fullsize.jsonl
PVGPU_FUSED=0 timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fused.py > $out/r02a_pytest_split.log 2>&1
echo "== split kernels: $(tail -1 $out/r02a_pytest_split.log)"
grep -E "^(FAILED|ERROR)" $out/r02a_pytest_split.log | head -20
cp $out/parity_fullsize.jsonl $out/r02a_parity_fullsize_split.jsonl 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_fused.py -q > $out/r02a_pytest_fused.log 2>&1
echo "== fused vs split: $(tail -1 $out/r02a_pytest_fused.log)"
grep -E "^(FAILED|ERROR)" $out/r02a_pytest_fused.log | head -40
rm -f $out/parity_fullsize.jsonl
timeout 900 python -m pytest tests -m gpu -q > $out/r02a_pytest.log 2>&1
echo "== full suite, fused default: $(tail -1 $out/r02a_pytest.log)"
grep -E "^(FAILED|ERROR)" $out/r02a_pytest.log | head -40
timeout 900 python bench.py --steps 3 --warmup 3 > $out/r02a_bench.log 2> $out/r02a_bench.err
echo "== bench (fused): rc $?"; tail -c 3000 $out/r02a_bench.log; tail -5 $out/r02a_bench.err
PVGPU_FUSED=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency --no-parity > $out/r02a_bench_split.log 2> $out/r02a_bench_split.err
echo "== bench (split): rc $?"; tail -c 1500 $out/r02a_bench_split.log; tail -5 $out/r02a_bench_split.err
# shuffle-vs-shared-memory exchange microbenchmark (north_star "warp shuffles"), plain then under ncu
scripts/microbench/exchange_bench > $out/r02_exchange_bench.json 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:k_exchange --csv --log-file $out/r02_exchange_ncu.csv scripts/microbench/exchange_bench > $out/r02_exchange_ncu.log 2>&1
echo "== exchange microbench:"; cat $out/r02_exchange_bench.json; grep -c k_exchange $out/r02_exchange_ncu.csv
