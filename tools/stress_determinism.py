#!/usr/bin/env python
"""Race stand-in for `compute-sanitizer --tool racecheck` (closed on this pool): the same batch is encoded and decoded
`reps` times while a second stream keeps the memory system busy with copies of varying size, and every result must equal
the first one BYTE FOR BYTE (and the first one equals the oracle on sampled frames).  A shared-memory race between the
owner-computed s = 2 words and their neighbours' reads, or between the levels, would show up as a run that differs.
    python tools/stress_determinism.py [reps] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import rustyhgi_b200 as hgi
from oracle import c as oc

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
h, w = 1080, 1920
g = torch.Generator(device="cuda").manual_seed(7)
yy = torch.arange(h, device="cuda", dtype=torch.int32)[:, None]
xx = torch.arange(w, device="cuda", dtype=torch.int32)[None, :]
k = torch.arange(n, device="cuda", dtype=torch.int32)[:, None, None]
noise = torch.randint(0, 24, (n, h, w), device="cuda", dtype=torch.int32, generator=g)
frames = (((xx * yy) // 97 + 31 * k + noise) & 255).to(torch.uint8).contiguous()
del noise
ctx = hgi.Context(0)
side = torch.cuda.Stream()
junk_a = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
junk_b = torch.empty_like(junk_a)
total = 0
for levels, q in ((4, 2), (4, 0), (6, 3)):
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(hgi.QuantizationLevel(q)), levels, ctx=ctx)
    dec = hgi.Decoder(hgi.Crossed, ctx=ctx)
    g0 = enc.encode_device(frames).clone()
    d0 = dec.decode_device(levels, g0).clone()
    torch.cuda.synchronize()
    for i in (0, n // 2, n - 1):
        wg, wr = oc.encode(frames[i].cpu().numpy(), levels, qlevel=q, want_recon=True)
        assert (g0[i].cpu().numpy() == wg).all() and (d0[i].cpu().numpy() == wr).all()
    gi, di = torch.empty_like(g0), torch.empty_like(d0)
    for r in range(reps):
        with torch.cuda.stream(side):
            m = (1 + (r * 37) % 512) << 20
            junk_b[:m].copy_(junk_a[:m], non_blocking=True)
        enc.encode_device(frames, grids_out=gi)
        dec.decode_device(levels, gi, images_out=di)
        assert torch.equal(gi, g0) and torch.equal(di, d0), f"run {r} differs (levels {levels}, q {q})"
        total += 1
torch.cuda.synchronize()
print(f"determinism ok: {total} encode+decode runs of {n} frames, all byte-identical to the first (which equals the oracle)")
