#!/usr/bin/env python
"""PCIe ceiling of the box for the e2e leg of bench.py: pinned host <-> device copies, one direction at a time and
both directions at once (two streams), in GB/s.  The e2e figure moves 1 byte per pixel each way per operation."""
import sys
import torch

n = 1 << 30                                   # 1 GiB per buffer
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a); s2.wait_event(a)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    torch.cuda.current_stream().wait_event(e1); torch.cuda.current_stream().wait_event(e2)
    b.record(); b.synchronize()
    return reps * n / (a.elapsed_time(b) * 1e-3) / 1e9


for _ in range(2):
    run(True, True, 2)
print(f"H2D alone  {run(True, False):6.1f} GB/s")
print(f"D2H alone  {run(False, True):6.1f} GB/s")
print(f"both       {run(True, True):6.1f} GB/s per direction")
