#!/usr/bin/env python
"""Digest of scripts/profile_round.sh's output: copies the judged artifacts into profiles/ and derives
profiles/traffic.json (measured DRAM bytes per frame of every kernel, from the ncu --set full capture) plus a markdown
table of the per-kernel counters.  Usage: python scripts/profile_digest.py r01_final [rows frames]"""
import csv, json, os, shutil, sys

tag = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 592
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 64
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
for name in (f"{tag}_bench.log", f"{tag}_launches.csv", f"{tag}_ncu_full_raw.csv"):
    shutil.copy(os.path.join(src, name), os.path.join(dst, name))
KIND = {"k_synth_ola": "synth_ola", "k_analyse": "analyse", "k_lock_peaks": "lock_peaks", "k_lock_chain": "lock_chain", "k_synthesise": "synthesise",
        "k_ola_resample": "ola_resample", "k_phase_lock": "phase_core", "k_phase_core": "phase_core"}
data = list(csv.reader(open(os.path.join(src, f"{tag}_ncu_full_raw.csv"))))
hdr = data[0]
traffic, table = {}, []
for r in data[2:]:
    d = dict(zip(hdr, r))
    kind = next((v for k, v in KIND.items() if k in d["Kernel Name"]), None)
    if kind is None or kind in traffic:
        continue
    f = lambda k: float(d[k]) if d.get(k) not in (None, "") else float("nan")
    unit = dict(zip(hdr, data[1]))
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = f("dram__bytes_read.sum") * scale.get(unit.get("dram__bytes_read.sum"), 1.0)
    wr = f("dram__bytes_write.sum") * scale.get(unit.get("dram__bytes_write.sum"), 1.0)
    n = rows * frames
    issue, pipe = f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
    traffic[kind] = {"dram_bytes_per_frame": (rd + wr) / n, "warp_instr_per_frame": f("smsp__inst_executed.sum") / n,
                     "issue_active_pct": issue, "l1_data_pipe_pct": pipe,
                     "limiter": "latency (serial chain)" if kind == "lock_chain" else ("l1_data_pipe" if pipe == pipe and pipe > issue else "issue"),
                     "source": f"profiles/{tag}_ncu_full_raw.csv (ncu --set full, {rows} rows x {frames} frames per launch)"}
    table.append((kind, f("gpu__time_duration.sum"), f("smsp__inst_executed.sum") / n, f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  f("sm__warps_active.avg.pct_of_peak_sustained_active"), (rd + wr) / n, f("launch__registers_per_thread"),
                  f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / n))
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
tot = sum(t[1] for t in table)
print("| kernel | us / launch (ncu) | share | warp-instr / frame | issue active % | warps active % | DRAM B / frame | regs | smem wavefronts / frame |")
print("|---|---|---|---|---|---|---|---|---|")
for t in table:
    print(f"| {t[0]} | {t[1]:.0f} | {100 * t[1] / tot:.0f} % | {t[2]:.0f} | {t[3]:.0f} | {t[4]:.0f} | {t[5]:.0f} | {t[6]:.0f} | {t[7]:.0f} |")
# launch list shares
import collections
sh = collections.Counter()
for r in csv.reader(open(os.path.join(src, f"{tag}_launches.csv"))):
    if len(r) > 5 and r[0].isdigit():
        kind = next((v for k, v in KIND.items() if k in r[4]), None)
        try:
            val = float(r[-1].replace(",", ""))
        except ValueError:
            continue
        if kind:
            sh[kind] += val
tl = sum(sh.values()) or 1
print("\nlaunch-list shares (ncu gpu__time_duration.sum, serialised):", {k: f"{100 * v / tl:.0f} %" for k, v in sh.items()})
