#!/bin/bash
# 8-GPU pass: host-copy probe matrix (which allocation feeds all GPUs), in-library multi-GPU batch tests, scaling bench
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/r02_topo.txt 2>&1
(nproc; numactl -H 2>/dev/null || ls /sys/devices/system/node/) >> $out/r02_topo.txt 2>&1
for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist)"; done >> $out/r02_topo.txt 2>&1
grep -i -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status >> $out/r02_topo.txt
grep -E "HugePages_Total|HugePages_Free|Hugepagesize" /proc/meminfo >> $out/r02_topo.txt
NG=$(nvidia-smi -L | wc -l)
echo "== $NG GPUs"
PROBE_GB=1.5 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 scripts/pcie_probe.py > $out/r02_pcie_probe_${NG}gpu.jsonl 2> $out/r02_pcie_probe.err
echo "== probe rc $?"; grep concurrently $out/r02_pcie_probe_${NG}gpu.jsonl | head -12; tail -3 $out/r02_pcie_probe.err
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $out/r02_pytest_multi_${NG}gpu.log 2>&1
echo "== multi-GPU tests: $(tail -1 $out/r02_pytest_multi_${NG}gpu.log)"
for alloc in numa torch; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $NG --steps 3 --warmup 3 --host-alloc $alloc --no-latency > $out/r02_bench_${NG}gpu_$alloc.log 2> $out/r02_bench_${NG}gpu_$alloc.err
python - "$out/r02_bench_${NG}gpu_$alloc.log" "$alloc" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print("==", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e s16", round(d["e2e"]["value"]), "f32", round(d["e2e"]["f32"]["value"]),
              "strong", d["strong_scaling"] and (round(d["strong_scaling"]["value"]), round(d["strong_scaling"].get("e2e", {}).get("value", 0))), "parity", d["parity"] and d["parity"]["device_resident_f32"])
PY
tail -2 $out/r02_bench_${NG}gpu_$alloc.err
done
