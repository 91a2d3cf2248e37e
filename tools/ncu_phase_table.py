#!/usr/bin/env python
"""Per-phase instruction table of a tile kernel from an ncu source-page export:
    ncu -i X.ncu-rep --page source --print-source sass --csv > src.csv
    python tools/ncu_phase_table.py src.csv <kernel-substring> [warps]
Phases are the SASS ranges between BAR.SYNCs.  Prints executed warp instructions per phase (per warp if the number
of warps is given), split by pipe, and the stall samples."""
import csv
import sys

ALU = ("LOP3", "PRMT", "SHF", "IADD3", "VIADD", "VIMNMX", "ISETP", "LEA", "SEL", "IABS", "HSET", "IMNMX", "FMNMX", "LOP", "SGXT", "BMSK", "POPC", "FLO", "MOV", "PLOP3", "CS2R", "S2R", "FSEL")
FMA = ("IMAD", "HFMA2", "HMUL2", "HADD2", "FFMA", "FMUL", "FADD", "IDP")
LSU = ("LDS", "STS", "LDG", "STG", "LD.", "ST.", "ATOM", "RED", "LDC")


def pipe(op):
    for p, names in (("alu", ALU), ("fma", FMA), ("lsu", LSU)):
        if any(op.startswith(n) for n in names):
            return p
    return "other"


def main():
    path, kern = sys.argv[1], sys.argv[2]
    warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = list(csv.reader(open(path)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name" and kern in rows[i][1]:
            break
        i += 1
    hdr = rows[i + 1]
    ci, cs, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    phases, cur = [], {"alu": 0, "fma": 0, "lsu": 0, "other": 0, "samples": 0, "imadhi": 0}
    for r in rows[i + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        src = r[csrc].strip()
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        n = int(r[ci] or 0)
        cur[pipe(op)] += n
        cur["samples"] += int(r[cs] or 0)
        if op.startswith("IMAD.HI"):
            cur["imadhi"] += n
        if op.startswith("BAR"):
            phases.append(cur)
            cur = {"alu": 0, "fma": 0, "lsu": 0, "other": 0, "samples": 0, "imadhi": 0}
    phases.append(cur)
    tot = sum(p["alu"] + p["fma"] + p["lsu"] + p["other"] for p in phases)
    tots = sum(p["samples"] for p in phases) or 1
    print(f"{kern}: {tot} warp instructions ({tot / warps:.1f} per warp)")
    print("phase   total   share    alu     fma     lsu   other  imad.hi  samples%")
    for k, p in enumerate(phases):
        t = p["alu"] + p["fma"] + p["lsu"] + p["other"]
        print(f"{k:5d} {t / warps:7.1f} {100 * t / tot:6.1f}% {p['alu'] / warps:7.1f} {p['fma'] / warps:7.1f} {p['lsu'] / warps:7.1f} "
              f"{p['other'] / warps:7.1f} {p['imadhi'] / warps:7.1f} {100 * p['samples'] / tots:8.1f}")


if __name__ == "__main__":
    main()
