#!/usr/bin/env python
"""Times encode (Lossless, Medium) and decode on 2048 resident 1080p frames with no result checks -- for kernel
experiments whose output is deliberately incomplete (e.g. HGI_VAR_STOP_AFTER phase cuts).  HGI_B200_LIB selects the
library.   python tools/time_encode.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rustyhgi_b200 as hgi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h, w = 1080, 1920
yy = torch.arange(h, device="cuda", dtype=torch.int32)[:, None]; xx = torch.arange(w, device="cuda", dtype=torch.int32)[None, :]
frames = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
for k0 in range(0, n, 256):
    k = torch.arange(k0, min(n, k0 + 256), device="cuda", dtype=torch.int32)[:, None, None]
    frames[k0:k0 + 256] = ((xx * yy + 31 * k) & 255).to(torch.uint8)
grids = torch.empty_like(frames); out = torch.empty_like(frames)
ctx = hgi.Context(0)


def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); b.synchronize(); return a.elapsed_time(b) / reps


res = []
for q in (hgi.QuantizationLevel.Lossless, hgi.QuantizationLevel.Medium):
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(q), 4, ctx=ctx)
    res.append(f"encode {q.name} {t(lambda: enc.encode_device(frames, grids_out=grids)):.3f} ms")
dec = hgi.Decoder(hgi.Crossed, ctx=ctx)
res.append(f"decode {t(lambda: dec.decode_device(4, grids, images_out=out)):.3f} ms")
print(os.environ.get("HGI_B200_LIB", "default"), "|", " | ".join(res))
