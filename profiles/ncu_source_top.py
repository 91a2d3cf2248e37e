#!/usr/bin/env python
"""Top source lines by executed instructions / stall samples from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv`.
usage: ncu_source_top.py src.csv <kernel-substring> [N]"""
import csv, sys
path, kern, n = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
cur, rows, fname = None, {}, None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur = r[1]
    elif r[0] == "Line No":
        continue
    elif cur and kern in cur and r[0].isdigit():
        key = (fname, int(r[0]))
        try:
            inst, samp = int(r[7]), int(r[6])
        except ValueError:
            continue
        src = r[1].strip()[:100]
        o = rows.setdefault(key, [0, 0, src])
        o[0] += inst
        o[1] += samp
tot_i = sum(v[0] for v in rows.values()) or 1
tot_s = sum(v[1] for v in rows.values()) or 1
print(f"kernel ~ {kern}: {tot_i} warp instructions, {tot_s} samples")
for (f, ln), (i, s, src) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*i/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% smp  {f}:{ln:<4d} {src}")
