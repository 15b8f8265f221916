#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel of an ncu report.

ncu's CSV export of the source page has no line numbers, so the per-instruction counters (SASS view) are joined, by
instruction index, with the line table nvdisasm prints for the same kernel of the library that was profiled.

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep k_lock_peaks [min_share_percent]
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = [b for b in blocks if kern in b["name"]][0]
hdr = b["rows"][0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) > ix["Instructions Executed"]]
# mangled-name fragment: take the template arguments into account through the demangled name's order of appearance
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "audiomod_b200", "libpvgpu.so")], cwd=tmp, capture_output=True)
dis = "".join(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, c)], capture_output=True, text=True).stdout
              for c in sorted(os.listdir(tmp)) if c.endswith(".cubin"))
funcs, name, line = defaultdict(list), None, None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        name, line = m.group(1), None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        funcs[name].append(line)
demangled = {n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() for n in funcs}
norm = lambda s: re.sub(r"\(int\)|\(bool\)|pvgpu::|\s", "", s)
cands = [n for n in funcs if norm(demangled[n]).startswith(norm(b["name"]).split("(")[0]) and len(funcs[n]) == len(data)]
if not cands:
    cands = [n for n in funcs if kern in n and len(funcs[n]) == len(data)]
if not cands:
    sys.exit("no nvdisasm function with %d instructions matches %s (rebuild the library that was profiled?)" % (len(data), b["name"]))
lines = funcs[cands[0]]
ie, ist = ix["Instructions Executed"], ix["Warp Stall Sampling (All Samples)"]
tot = sum(int(r[ie]) for r in data)
tst = sum(int(r[ist]) for r in data) or 1
per = defaultdict(lambda: [0, 0])
for r, ln in zip(data, lines):
    per[ln][0] += int(r[ie])
    per[ln][1] += int(r[ist])
src = {}
print("%s: %d SASS instructions, %d warp-instructions executed, %d stall samples" % (b["name"][:70], len(data), tot, tst))
for ln in sorted(per, key=lambda k: (k is None, k)):
    n, s = per[ln]
    if 100.0 * n / tot < min_share and 100.0 * s / tst < 2 * min_share:
        continue
    text = ""
    if ln:
        if ln[0] not in src:
            p = os.path.join(root, "audiomod_b200", "csrc", ln[0])
            src[ln[0]] = open(p).read().splitlines() if os.path.exists(p) else []
        if ln[1] - 1 < len(src[ln[0]]):
            text = src[ln[0]][ln[1] - 1].strip()[:100]
    print("%5.1f%% instr %5.1f%% stall  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / tst, ln[0] if ln else "?", ln[1] if ln else "?", text))
