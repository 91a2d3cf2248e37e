"""BASELINE.json configs at full size on the GPU (-m gpu): C4 (16384^2, L8, Medium) against the
oracle directly and through the row-band decomposition; C5-style 1080p batches through
size-independent properties plus sampled frames against the oracle."""
import numpy as np
import pytest

import rustyhgi_b200 as hgi
from oracle import c as oc
from rustyhgi_b200 import sharding

pytestmark = pytest.mark.gpu
Q = hgi.QuantizationLevel


def test_config4_16384_l8_medium_and_bands():
    import torch
    n = 16384
    x = torch.arange(n, device="cuda", dtype=torch.int32)
    img_t = ((x[None, :] * x[:, None]) & 255).to(torch.uint8).contiguous()      # benches/bench.rs:26-28
    img = img_t.cpu().numpy()
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(Q.Medium), 8)
    dec = hgi.Decoder(hgi.Crossed)
    grid_t = enc.encode_device(img_t)
    out_t = dec.decode_device(8, grid_t)
    torch.cuda.synchronize()
    want_g, want_r = oc.encode(img, 8, qlevel=2, want_recon=True)
    assert (grid_t.cpu().numpy() == want_g).all()
    assert (out_t.cpu().numpy() == want_r).all()
    assert int(np.abs(want_r.astype(int) - img).max()) <= 20
    # row bands with the S+1 overlap reproduce the full plane (8 ranks' worth, run back to back)
    for b in sharding.plan_bands(n, 8, 8):
        g = enc.encode_device(img_t[b.y0:b.in_y1].contiguous())
        d = dec.decode_device(8, grid_t[b.y0:b.in_y1].contiguous())
        assert torch.equal(g[:b.rows_out], grid_t[b.y0:b.y1])
        assert torch.equal(d[:b.rows_out], out_t[b.y0:b.y1])


def test_config5_batch_1080p_properties():
    import torch
    n, h, w = 256, 1080, 1920
    yy = torch.arange(h, device="cuda", dtype=torch.int32)[:, None]
    xx = torch.arange(w, device="cuda", dtype=torch.int32)[None, :]
    k = torch.arange(n, device="cuda", dtype=torch.int32)[:, None, None]
    frames = ((xx * yy + 31 * k) & 255).to(torch.uint8).contiguous()
    dec = hgi.Decoder(hgi.Crossed)
    for q, err in ((Q.Lossless, 0), (Q.Medium, 20)):
        enc = hgi.Encoder(hgi.Crossed, hgi.Linear(q), 4)
        hist = torch.empty((n, 256), dtype=torch.int32, device="cuda")
        grids = enc.encode_device(frames, hist_out=hist)
        back = dec.decode_device(4, grids)
        torch.cuda.synchronize()
        diff = (back.to(torch.int16) - frames.to(torch.int16)).abs().max().item()
        assert diff <= err
        assert torch.equal(back[:, ::16, ::16], frames[:, ::16, ::16])           # seeds are raw pixels
        assert int(hist.sum().item()) == n * h * w                               # checksum of histograms
        assert torch.equal(hist.sum(0), torch.bincount(grids.reshape(-1).to(torch.int64), minlength=256).to(torch.int32))
        # encode is deterministic and batch-position independent
        again = enc.encode_device(frames[100:103].contiguous())
        assert torch.equal(again, grids[100:103])
        for i in (0, 101, n - 1):
            f = frames[i].cpu().numpy()
            wg = oc.encode(f, 4, qlevel=int(q))
            assert (grids[i].cpu().numpy() == wg).all()
            assert (back[i].cpu().numpy() == oc.decode(wg, 4)).all()
