#!/usr/bin/env python
"""A/B harness for kernel variants: runs bench.py (device-resident leg only) with each library given on the command
line, interleaved ABAB... on the same box so that box-to-box and thermal drift cancel, and prints the median
per-kernel milliseconds.    python tools/ab_bench.py [--rounds 3] [--frames 2048] name=path.so[,ENV=VALUE...] ..."""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    libs = [l.split("=", 1) for l in a.libs]
    res = {n: [] for n, _ in libs}
    for rnd in range(a.rounds):
        order = libs[rnd % len(libs):] + libs[:rnd % len(libs)]      # rotate: the first run after a pause sees a cooler, faster GPU
        for name, path in order:
            path, *extra = path.split(",")
            env = dict(os.environ, HGI_B200_LIB=os.path.abspath(path), **dict(e.split("=", 1) for e in extra))
            out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(a.steps), "--no-cpu", "--no-e2e",
                                  "--frames", str(a.frames)], env=env, capture_output=True, text=True)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")]
            if not line:
                print(name, "FAILED", out.stderr[-300:])
                continue
            d = json.loads(line[-1])
            res[name].append({k: v["ms"] for k, v in d["roofline"]["per_kernel"].items()} | {"step": d["ms_per_step"]}
                             | {"sm_mhz": d.get("clocks", {}).get("sm_mhz") or 0.0})
    keys = ["step", "encode_lossless", "decode_lossless", "encode_medium", "decode_medium", "sm_mhz"]
    print(f"{'variant':24s} " + " ".join(f"{k:>16s}" for k in keys))
    for name, runs in res.items():
        if runs:
            print(f"{name:24s} " + " ".join(f"{statistics.median(r[k] for r in runs):16.3f}" for k in keys))


if __name__ == "__main__":
    main()
