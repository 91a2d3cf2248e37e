#!/usr/bin/env python
"""Bandwidth of the reduction / colour kernels: error metrics (2 B/pixel), RGB->luma (4 B/pixel), histogram (1 B/byte)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rustyhgi_b200 as hgi

n = 1 << 30
a = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
b = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
ctx = hgi.Context(0); L = hgi.lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")


def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize(); return e0.elapsed_time(e1) / reps


m = hgi.error_metrics(a[:1 << 20].cpu().numpy(), b[:1 << 20].cpu().numpy(), ctx=ctx)
d = a[:1 << 20].to(torch.int64) - b[:1 << 20].to(torch.int64)
assert m["sum_sq"] == int((d * d).sum().item()) and m["max_abs"] == int(d.abs().max().item()), m
st = 1
ms = t(lambda: ctx.check(L.hgi_error_metrics_dev(ctx._h, a.data_ptr(), b.data_ptr(), n, out.data_ptr(), st), "err"))
print(f"error metrics : {ms:.3f} ms per GiB pair  -> {2 * n / ms / 1e6:.0f} GB/s")
ms = t(lambda: ctx.check(L.hgi_error_metrics_dev(ctx._h, a.data_ptr() + 1, b.data_ptr() + 1, n - 1, out.data_ptr(), st), "err"))
print(f"error metrics, unaligned planes : {ms:.3f} ms -> {2 * n / ms / 1e6:.0f} GB/s")
npx = n // 4
luma = torch.empty(npx, dtype=torch.uint8, device="cuda")
ms = t(lambda: ctx.check(L.hgi_rgb_to_luma_dev(ctx._h, a.data_ptr(), npx, luma.data_ptr(), st), "luma"))
print(f"rgb -> luma   : {ms:.3f} ms per {npx >> 20} Mpixel -> {4 * npx / ms / 1e6:.0f} GB/s")
hist = torch.empty((1024, 256), dtype=torch.int32, device="cuda")
ms = t(lambda: ctx.check(L.hgi_histogram_dev(ctx._h, a.data_ptr(), n // 1024, 1024, hist.data_ptr(), st), "hist"))
print(f"histogram     : {ms:.3f} ms per GiB (random bytes) -> {n / ms / 1e6:.0f} GB/s")
