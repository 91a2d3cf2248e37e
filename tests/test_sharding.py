"""Host-side partitioning across GPUs (rustyhgi_b200/sharding.py), including a world_size-2 gloo
run of the by-image and by-row-band paths.  On the CPU the per-rank compute is done by the oracle
(test infrastructure standing in for the device); what is under test is the partition / overlap /
gather logic that bench.py and a multi-GPU caller use."""
import os
import socket

import numpy as np
import pytest

from conftest import photo_like
from oracle import c as oc
from rustyhgi_b200 import sharding


def test_split_batch_covers_everything():
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [sharding.split_batch(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("h,levels,n", [(16384, 8, 8), (2614, 6, 4), (1080, 4, 8), (100, 8, 4), (256, 4, 3)])
def test_plan_bands(h, levels, n):
    S = 1 << levels
    bands = sharding.plan_bands(h, levels, n)
    assert 1 <= len(bands) <= n
    assert bands[0].y0 == 0 and bands[-1].y1 == h
    for a, b in zip(bands, bands[1:]):
        assert a.y1 == b.y0
    for b in bands:
        assert b.y0 % S == 0 and (b.y1 % S == 0 or b.y1 == h)
        assert b.in_y1 == min(h, b.y1 + S + 1)


@pytest.mark.parametrize("levels,q", [(4, 2), (5, 3), (3, 0)])
def test_bands_reproduce_full_plane(levels, q):
    img = photo_like(190, 300, 21)
    full_g, full_r = oc.encode(img, levels, qlevel=q, want_recon=True)
    g = np.empty_like(img)
    d = np.empty_like(img)
    for b in sharding.plan_bands(img.shape[0], levels, 4):
        g[b.y0:b.y1] = oc.encode(img[b.y0:b.in_y1], levels, qlevel=q)[:b.rows_out]
        d[b.y0:b.y1] = oc.decode(full_g[b.y0:b.in_y1], levels)[:b.rows_out]
    assert (g == full_g).all() and (d == full_r).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        levels = 4
        # by image: every rank encodes its contiguous slice, results are all-gathered
        imgs = np.stack([photo_like(64, 48, s) for s in range(7)])
        first, last = sharding.split_batch(len(imgs), world, rank)
        mine = oc.encode_batch(imgs[first:last], levels, qlevel=2, n_threads=1)
        parts = [None] * world
        dist.all_gather_object(parts, (first, mine))
        grids = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        ok_batch = bool((grids == oc.encode_batch(imgs, levels, qlevel=2, n_threads=1)).all())
        # by row band with the S+1 overlap, no exchange
        plane = photo_like(96, 160, 5)
        bands = sharding.plan_bands(plane.shape[0], levels, world)
        b = bands[rank] if rank < len(bands) else None
        part = oc.encode(plane[b.y0:b.in_y1], levels, qlevel=3)[:b.rows_out] if b else np.zeros((0, 96), np.uint8)
        dist.all_gather_object(parts, (b.y0 if b else 1 << 30, part))
        grid = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        ok_band = bool((grid == oc.encode(plane, levels, qlevel=3)).all())
        # the timing reduction bench.py uses: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put((ok_batch, ok_band, float(t.item())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == (True, True, 2.0)


def test_c_abi_band_plan_equals_the_python_one():
    """hgi_plan_bands (what hgi_pool_* uses; pure host arithmetic, callable without a GPU) == sharding.plan_bands."""
    import ctypes
    import rustyhgi_b200 as hgi
    from rustyhgi_b200 import _lib
    L = hgi.lib()
    for height in (1, 255, 256, 257, 1000, 16384, 16385, 100000):
        for levels in (0, 1, 4, 8, 12, 31):
            for n in (1, 2, 3, 8, 17):
                bands = (_lib.BandStruct * n)()
                cnt = ctypes.c_int(-1)
                assert L.hgi_plan_bands(height, levels, n, ctypes.cast(bands, ctypes.c_void_p), ctypes.byref(cnt)) == 0
                want = hgi.sharding.plan_bands(height, levels, n)
                got = [(bands[i].y0, bands[i].y1, bands[i].in_y1) for i in range(cnt.value)]
                assert got == [(b.y0, b.y1, b.in_y1) for b in want], (height, levels, n)
    assert L.hgi_plan_bands(100, 32, 2, ctypes.cast((_lib.BandStruct * 2)(), ctypes.c_void_p), ctypes.byref(ctypes.c_int())) == -1
