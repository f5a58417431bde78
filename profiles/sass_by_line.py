#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the same
cubin and aggregate executed warp instructions / stall samples per CUDA source line.
usage: sass_by_line.py <ncu_source.csv> <nvdisasm -g -c output> <kernel mangled-name substring> [topN]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = rows[2:]
# nvdisasm: collect (line marker) per instruction within the kernel section
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//---") and kname in l)
cur = ("?", 0)
inline = ""
seq = []
for l in lines[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
n = min(len(seq), len(sass))
print(f"sass rows {len(sass)} disasm instrs {len(seq)}", file=sys.stderr)
agg = defaultdict(lambda: [0, 0])
tot = [0, 0]
for i in range(n):
    inst = int(float(sass[i][ci["Instructions Executed"]] or 0))
    smp = int(float(sass[i][ci["# Samples"]] or 0))
    agg[seq[i]][0] += inst
    agg[seq[i]][1] += smp
    tot[0] += inst
    tot[1] += smp
print(f"total warp-instr {tot[0]}  samples {tot[1]}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<5d} inst {v[0]:>12d} ({100*v[0]/tot[0]:5.1f}%)  samples {v[1]:>7d} ({100*v[1]/max(tot[1],1):5.1f}%)")
