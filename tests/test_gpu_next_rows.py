"""SURVEY.md 8(f) "next" rows on the GPU: f-2 RGB->luma, f-3 `hgi test` metrics, f-4 the CLI shell."""
import io
import os
import zlib

import numpy as np
import pytest

import rustyhgi_b200 as hgi
from conftest import GOLDEN, load_plane, sha16
from oracle import c as oc
from rustyhgi_b200 import cli

pytestmark = pytest.mark.gpu


def test_rgb_to_luma_exhaustive_all_16m_triples():
    """Every (r,g,b) in 2^24 against the CPU restatement of image-0.19's f32 formula (no FMA)."""
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
    got = hgi.rgb_to_luma(rgb)
    want = oc.rgb_to_luma(rgb)
    assert int((got != want).sum()) == 0
    gray = np.repeat(np.arange(256, dtype=np.uint8), 3).reshape(1, 256, 3)      # f32 sum lands on v-1 for some v
    assert (hgi.rgb_to_luma(gray) == oc.rgb_to_luma(gray)).all()


def test_rgb_to_luma_ragged_and_unaligned():
    rng = np.random.default_rng(2)
    for h, w in ((1, 1), (3, 5), (7, 16), (33, 31), (250, 243)):
        rgb = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        assert (hgi.rgb_to_luma(rgb) == oc.rgb_to_luma(rgb)).all()


def test_docs_source_image_luma_matches_golden():
    """The palette PNG of the docs pair goes RGB -> luma on the GPU and must equal the golden plane."""
    from PIL import Image
    rgb = np.array(Image.open(os.path.join(GOLDEN, "docs_lena_source.png")).convert("RGB"))
    # the fixture is already luma (gray triples): to_luma of gray input lands on v or v-1
    assert (hgi.rgb_to_luma(rgb) == oc.rgb_to_luma(rgb)).all()


def test_archive_with_gpu_built_frequency_tables():
    """f-1: encode on the GPU with fused per-image histograms, entropy-code on the host from those tables."""
    img = load_plane("fullhd")
    enc = hgi.Encoder(hgi.Crossed, hgi.Linear(hgi.QuantizationLevel.Medium), 4)
    grids, hist = enc.encode_batch(img[None], want_hist=True)
    md = hgi.Metadata(2, 0, 1920, 1080, 4)
    arch = hgi.Archive(md, hgi.Grid(grids[0], 1920))
    out = io.BytesIO()
    arch.serialize_to_writer(out, entropy="huffman", hist=hist)          # one table for the whole grid
    assert hgi.Archive.deserialize_from_reader(io.BytesIO(out.getvalue())) == arch
    out2 = io.BytesIO()
    arch.serialize_to_writer(out2, entropy="huffman", block_rows=120)    # nine tables, built by hgi_histogram_u8
    raw = out2.getvalue()
    assert zlib.decompress(raw[28:], -15)[8:-8] == grids[0].tobytes()
    assert len(raw) < img.size // 4                                       # sanity: it does compress
    after = hgi.Decoder(hgi.Crossed).decode((1920, 1080), 4, hgi.Archive.deserialize_from_reader(io.BytesIO(raw)).grid)
    assert sha16(after) == "96db2c544587bccb"


def test_rle_tables_from_the_gpu_equal_the_cpu_statement_of_the_parse():
    """hgi_rle_hist_kernel counts exactly the tokens hgi_archive_serialize_rle will write (tests/rle_model.py states
    the parse on the CPU); the resulting archive inflates to bincode(Grid) and is within 10 % of zlib level 9."""
    import torch
    from rle_model import rle_table
    ctx = hgi.Context(0)
    L = hgi.lib()
    rng = np.random.default_rng(3)
    planes = [oc.encode(load_plane("fullhd"), 4, qlevel=2), oc.encode(load_plane("lena_tif"), 4, qlevel=0),
              rng.integers(0, 256, (37, 123)).astype(np.uint8), np.zeros((50, 700), np.uint8),
              np.repeat(rng.integers(0, 4, (20, 40)).astype(np.uint8), 40, axis=1)]
    for g in planes:
        buf = g.reshape(-1)
        for block in (None, 512 * 7, 512 * 64):
            nb = 1 if block is None else -(-buf.size // block)
            want = rle_table(buf, block)
            got = np.zeros((nb, 288), np.uint32)
            ctx.check(L.hgi_rle_histogram_u8(ctx._h, buf.ctypes.data, buf.size, block or buf.size, nb, got.ctypes.data), "rle")
            assert (got == want).all(), (g.shape, block)
            d = torch.from_numpy(buf).cuda()
            dh = torch.empty((nb, 288), dtype=torch.int32, device="cuda")
            ctx.check(L.hgi_rle_histogram_dev(ctx._h, d.data_ptr(), buf.size, block or buf.size, nb, dh.data_ptr(), None), "rle dev")
            ctx.synchronize()
            assert (dh.cpu().numpy().astype(np.uint32) == want).all()
    g = planes[0]
    arch = hgi.Archive(hgi.Metadata(2, 0, 1920, 1080, 4), hgi.Grid(g, 1920))
    out = io.BytesIO()
    arch.serialize_to_writer(out, entropy="rle", ctx=ctx)                 # tables built on the GPU inside
    raw = out.getvalue()
    payload = g.size.to_bytes(8, "little") + g.tobytes() + (1920).to_bytes(8, "little")
    assert zlib.decompress(raw[28:], -15) == payload
    z9 = zlib.compressobj(9, zlib.DEFLATED, -15)
    assert len(raw) <= 1.10 * len(z9.compress(payload) + z9.flush()) and len(raw) <= 250 * 1024
    assert hgi.Archive.deserialize_from_reader(io.BytesIO(raw)) == arch
    ctx.close()


def test_cli_test_subcommand_config1(tmp_path, monkeypatch, capsys):
    """BASELINE config 1: `hgi test res/LENA.TIF` (level 4, Medium) -> `Uncompressed: 64 kb`, `SD: 9.17`."""
    monkeypatch.chdir(tmp_path)
    src = os.path.join(GOLDEN, "lena_tif.png")
    cli.main(["test", src, "-l", "4", "-q", "Medium", "-s", "_hgi"])
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Uncompressed: 64 kb"
    assert out[3] == "SD:           9.17"
    assert out[1].startswith("Compressed:   ") and out[2].startswith("Ratio:        ")
    from PIL import Image
    after = np.array(Image.open(tmp_path / "lena_tif_hgi.png"))
    assert sha16(after) == "e17f5ad9f400234e"
    raw = (tmp_path / "lena_tif_hgi.hgi").read_bytes()
    assert raw[:4] == bytes.fromhex("55a5adba")
    payload = zlib.decompress(raw[28:], -15)
    assert sha16(np.frombuffer(payload[8:-8], np.uint8)) == "3a992020370c96a4"


def test_cli_encode_decode_roundtrip(tmp_path):
    src = os.path.join(GOLDEN, "lena_tif.png")
    arch, png = str(tmp_path / "a.hgi"), str(tmp_path / "a.png")
    cli.main(["encode", "-i", src, "-o", arch, "-q", "low"])            # default level 4 (src/options.rs:55)
    cli.main(["decode", "-i", arch, "-o", png])
    from PIL import Image
    got = np.array(Image.open(png))
    want = oc.decode(oc.encode(load_plane("lena_tif"), 4, qlevel=oc.LOW), 4)
    assert (got == want).all()


def test_cli_swallows_errors_like_reference(tmp_path, capsys):
    cli.main(["decode", "-i", str(tmp_path / "missing.hgi"), "-o", str(tmp_path / "x.png")])
    assert "An error occured" in capsys.readouterr().err               # src/main.rs:130-134


def test_host_reductions_in_many_chunks():
    """hgi_histogram_u8 / hgi_error_metrics_u8 / hgi_rgb_to_luma_u8 cut their input into chunks that alternate between
    two pipeline slots; with 1 MB chunks a 5.3 MB input is six of them (odd count, ragged tail)."""
    ctx = hgi.Context()
    ctx.set_pipeline(chunk_mb=1)
    rng = np.random.default_rng(9)
    n = 5 * (1 << 20) + 333_333
    a = rng.integers(0, 256, n).astype(np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, n), 0, 255).astype(np.uint8)
    assert (hgi.histogram(a, ctx=ctx) == np.bincount(a, minlength=256)).all()
    diff = a.astype(np.int64) - b.astype(np.int64)
    m = hgi.error_metrics(a, b, ctx=ctx)
    assert m["sum_sq"] == int((diff * diff).sum()) and m["max_abs"] == int(np.abs(diff).max())
    rgb = rng.integers(0, 256, (1, n // 3, 3)).astype(np.uint8)      # 1.78 M pixels: two chunks of 1 M pixels
    assert (hgi.rgb_to_luma(rgb, ctx=ctx) == oc.rgb_to_luma(rgb)).all()
    for rep in range(3):                                               # slots are reused across calls
        assert (hgi.histogram(b, ctx=ctx) == np.bincount(b, minlength=256)).all()
    ctx.close()
