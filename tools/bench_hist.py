#!/usr/bin/env python
"""Cost of the residual histogram: encode with hist_out (the library chains hgi_hist_kernel behind the encode
kernel) vs encode alone vs hgi_histogram_dev on its own."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rustyhgi_b200 as hgi

def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b) / n

n, h, w = 1024, 1080, 1920
yy = torch.arange(h, device="cuda", dtype=torch.int32)[:, None]; xx = torch.arange(w, device="cuda", dtype=torch.int32)[None, :]
frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
for k0 in range(0, n, 256):
    k = torch.arange(k0, k0 + 256, device="cuda", dtype=torch.int32)[:, None, None]
    frames[k0:k0 + 256] = ((xx * yy + 31 * k) & 255).to(torch.uint8)
grids = torch.empty_like(frames); hist = torch.empty((n, 256), dtype=torch.int32, device="cuda")
ctx = hgi.Context(0); L = hgi.lib()
for q in (hgi.QuantizationLevel.Lossless, hgi.QuantizationLevel.Medium):
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(q), 4, ctx=ctx)
    plain = t(lambda: enc.encode_device(frames, grids_out=grids))
    fused = t(lambda: enc.encode_device(frames, grids_out=grids, hist_out=hist))
    st = torch.cuda.current_stream().cuda_stream or 1
    sep = t(lambda: ctx.check(L.hgi_histogram_dev(ctx._h, grids.data_ptr(), h * w, n, hist.data_ptr(), st), "hist"))
    print(f"{q.name}: encode {plain:.3f} ms, encode with hist_out {fused:.3f} ms (+{fused - plain:.3f}), hgi_histogram_dev alone {sep:.3f} ms")
