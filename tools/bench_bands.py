#!/usr/bin/env python
"""BASELINE config 4 across the GPUs of one box: synthetic 16384x16384 8-bit plane, level 8, Medium, cut into
row bands with the S+1 = 257-row overlap (rustyhgi_b200/sharding.py) -- strong scaling, no data-path collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_bands.py

Each rank holds its band (+ overlap) in HBM, encodes and decodes it `--steps` times (CUDA events, max over ranks via
one all-reduce) and rank 0 prints one JSON line with the whole-plane Mpixel/s.  Band results are bit-exact parts of
the full-plane result (tests/test_gpu_fullsize.py, tests/test_sharding.py).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import rustyhgi_b200 as hgi
from rustyhgi_b200 import sharding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--one-band-of", type=int, default=0, help="single GPU: time band 0 of an N-band plan (what one rank of N does)")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n, L = a.size, a.levels
    bands = sharding.plan_bands(n, L, a.one_band_of if a.one_band_of else world)
    b = bands[rank] if rank < len(bands) else None
    ctx = hgi.Context(local)
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(hgi.QuantizationLevel.Medium), L, ctx=ctx)
    dec = hgi.Decoder(hgi.Crossed, ctx=ctx)
    ms = 0.0
    if b is not None:
        x = torch.arange(n, device=dev, dtype=torch.int32)
        y = torch.arange(b.y0, b.in_y1, device=dev, dtype=torch.int32)
        band = ((x[None, :] * y[:, None]) & 255).to(torch.uint8).contiguous()      # benches/bench.rs:26-28
        grid = torch.empty_like(band)
        out = torch.empty_like(band)

        def step():
            enc.encode_device(band, grids_out=grid)
            dec.decode_device(L, grid, images_out=out)      # decoding a band needs the same overlap rows of the grid

        for _ in range(a.warmup):
            step()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        assert int((out[:b.rows_out].to(torch.int16) - band[:b.rows_out].to(torch.int16)).abs().max().item()) <= 20
    elif dist:
        dist.barrier()
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"workload": f"synthetic {n}x{n}, level {L}, Medium, {len(bands)} row band(s) + {(1 << L) + 1}-row overlap",
                          "n_gpus": world, "ms_per_step": ms, "metric": "encode+decode Mpixel/s",
                          "value": n * n / (ms * 1e-3) / 1e6, "scaling": "strong",
                          "rows_in_per_rank": [bb.rows_in for bb in bands], "rows_out_per_rank": [bb.rows_out for bb in bands]}),
              flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
