#!/usr/bin/env python
"""End-to-end (host buffers) ceiling and tuning on N ranks of one box; launch under torchrun like bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_e2e_sweep.py [--frames 512]

1. PCIe ceiling with ALL ranks copying at once (what the e2e leg of bench.py competes for): pinned host <-> device, H2D
   alone, D2H alone, both directions at once; per rank and summed over the ranks.
2. The e2e leg itself (hgi_encode_batch_u8 + hgi_decode_batch_u8 on pinned buffers, Lossless + Medium) for each
   (chunk MiB, slots) setting of the host pipeline (hgi_ctx_set_pipeline), as Mpixel/s over all ranks, as GB/s per
   direction per rank, and as a fraction of the both-directions ceiling of step 1.
Rank 0 prints one JSON line per measurement."""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import rustyhgi_b200 as hgi
from bench import H, LEVELS, QLEVELS, W, bind_to_gpu_numa_node

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512, help="frames in the pinned buffers of each rank")
ap.add_argument("--configs", default="16:3,32:3,64:2,64:3,64:4,128:3", help="chunk_mb:slots,...")
ap.add_argument("--no-numa", action="store_true")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
numa = None if (a.no_numa or world == 1) else bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(**kw):
    if rank == 0:
        print(json.dumps(dict(kw, n_ranks=world, numa_binding=numa)), flush=True)


# ---- 1. PCIe ceiling, all ranks at once -------------------------------------------------------------------------------
n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_in.zero_(); h_out.zero_()                      # first touch on this rank's NUMA node
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.zeros(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def copies(h2d, d2h, reps=6):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    return reps * n / dt / 1e9                    # GB/s per direction per rank, at the pace of the slowest rank


copies(True, True, 2)
ceil = {"h2d_alone": copies(True, False), "d2h_alone": copies(False, True), "both": copies(True, True)}
emit(what="pcie_ceiling", GBps_per_direction_per_rank=ceil, GBps_per_direction_all_ranks={k: v * world for k, v in ceil.items()},
     method="512 MiB pinned copies, 6 per direction, all ranks at once, wall time of the slowest rank")
del d_in, d_out, h_in, h_out

# ---- 2. the e2e leg under different pipeline settings ------------------------------------------------------------------
n_e = a.frames
yy = torch.arange(H, device=dev, dtype=torch.int32)[:, None]
xx = torch.arange(W, device=dev, dtype=torch.int32)[None, :]
h_img = torch.empty((n_e, H, W), dtype=torch.uint8, pin_memory=True)
h_grid = torch.empty((n_e, H, W), dtype=torch.uint8, pin_memory=True)
h_back = torch.empty((n_e, H, W), dtype=torch.uint8, pin_memory=True)
for k0 in range(0, n_e, 128):
    k = torch.arange(rank * n_e + k0, rank * n_e + min(k0 + 128, n_e), device=dev, dtype=torch.int32)[:, None, None]
    h_img[k0:k0 + k.shape[0]].copy_(((xx * yy + 31 * k) & 255).to(torch.uint8))
h_grid.zero_(); h_back.zero_()
torch.cuda.synchronize()
L = hgi.lib()
ctx = hgi.Context(local)
params = [hgi.Encoder(hgi.Crossed, hgi.Linear(hgi.QuantizationLevel(q)), LEVELS, ctx=ctx)._p() for q in QLEVELS]
pi, pg, pb = h_img.numpy().ctypes.data, h_grid.numpy().ctypes.data, h_back.numpy().ctypes.data


def e2e_step():
    for p in params:
        ctx.check(L.hgi_encode_batch_u8(ctx._h, pi, n_e, W, H, ctypes.byref(p), pg, None), "encode")
        ctx.check(L.hgi_decode_batch_u8(ctx._h, pg, n_e, W, H, ctypes.byref(p), pb), "decode")


for cfg in a.configs.split(","):
    chunk, slots = (int(v) for v in cfg.split(":"))
    ctx.check(L.hgi_ctx_set_pipeline(ctx._h, chunk, slots), "set_pipeline")
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        e2e_step()
    barrier()
    dt = max_over_ranks((time.perf_counter() - t0) / reps)
    bytes_dir = len(QLEVELS) * 2 * n_e * W * H              # per rank and direction and step
    gbps = bytes_dir / dt / 1e9
    emit(what="e2e", chunk_mb=chunk, slots=slots, frames_per_rank=n_e, ms_per_step=dt * 1e3,
         Mpixel_s_all_ranks=world * len(QLEVELS) * n_e * W * H / dt / 1e6, GBps_per_direction_per_rank=gbps,
         frac_of_both_directions_ceiling=gbps / ceil["both"])
assert int((h_back[:2].to(torch.int16) - h_img[:2].to(torch.int16)).abs().max().item()) <= 20
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
