"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/hgi.h declares,
and its host-only entry points (quantizer table, archive container, argument validation) behave
like the reference.  No GPU compute is attempted here."""
import ctypes
import io
import os
import re
import zlib

import numpy as np
import pytest

import rustyhgi_b200 as hgi
from conftest import ROOT, get_plane
from oracle import c as oc


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hgi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    L = ctypes.CDLL(hgi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/hgi.h but not exported"
    assert sorted(hgi._lib.PROTOTYPES) == names


def test_rust_sys_crate_lists_every_symbol():
    """bindings/rust/hgi-sys (source only, no Rust toolchain here) must stay a complete transcription of hgi.h."""
    text = open(os.path.join(ROOT, "bindings", "rust", "hgi-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (hgi_[a-z0-9_]+)\s*\(", text))
    assert bound == set(declared_functions())


def test_library_has_sm100a_kernels_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", hgi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


@pytest.mark.parametrize("level,err", [(0, 0), (1, 10), (2, 20), (3, 30)])
def test_quantizator_tables(level, err):
    """Linear::from(level) (src/quantizator.rs:41-63): table and error()."""
    lin = hgi.Linear.from_level(hgi.QuantizationLevel(level))
    want, werr = oc.quant_table(oc.QUANT_LINEAR, level)
    assert lin.error() == err == werr
    assert (lin.table == want).all()
    scale = 2 * err + 1
    assert all(lin.quantize(v) == ((v + err) // scale * scale) & 255 for v in range(256))
    noop = hgi.NoOp.from_level(hgi.QuantizationLevel(level))
    assert noop.error() == 0 and (noop.table == np.arange(256)).all()


def test_quantization_level_parsing():
    assert hgi.QuantizationLevel.parse("medium") is hgi.QuantizationLevel.Medium
    assert hgi.QuantizationLevel.parse("LOSSLESS") is hgi.QuantizationLevel.Lossless
    with pytest.raises(ValueError):
        hgi.QuantizationLevel.parse("loseless")       # README's spelling is not a variant


def test_archive_header_and_payload_layout():
    """8x8 Lossless L3 header bytes + bincode(Grid) payload, src/archive.rs:31-41."""
    grid = oc.encode(get_plane("unit_8x8"), 3, qlevel=0)
    md = hgi.Metadata(hgi.QuantizationLevel.Lossless, hgi.InterpolationType.Crossed, 8, 8, 3)
    buf = io.BytesIO()
    hgi.Archive(md, hgi.Grid(grid, 8)).serialize_to_writer(buf)
    raw = buf.getvalue()
    assert raw[:28].hex() == "55a5adba" "00000000" "00000000" "08000000" "08000000" "0300000000000000"
    payload = zlib.decompress(raw[28:], -15)              # raw DEFLATE, no zlib header
    assert payload == (64).to_bytes(8, "little") + grid.tobytes() + (8).to_bytes(8, "little")


def test_archive_serde_roundtrip_like_reference_test():
    """src/lib.rs:99-125 `serde`: serialize -> deserialize -> assert_eq."""
    grid = oc.encode(get_plane("unit_8x8"), 3, qlevel=0)
    archive = hgi.Archive(hgi.Metadata(0, 0, 8, 8, 3), hgi.Grid(grid, 8))
    buf = io.BytesIO()
    archive.serialize_to_writer(buf)
    assert hgi.Archive.deserialize_from_reader(io.BytesIO(buf.getvalue())) == archive


def test_archive_large_and_foreign_stream():
    rng = np.random.default_rng(0)
    grid = rng.integers(0, 7, 300_000).astype(np.uint8)
    md = hgi.Metadata(2, 0, 600, 500, 4)
    buf = io.BytesIO()
    hgi.Archive(md, hgi.Grid(grid, 600)).serialize_to_writer(buf)
    back = hgi.Archive.deserialize_from_reader(io.BytesIO(buf.getvalue()))
    assert back.metadata == md and back.grid == hgi.Grid(grid, 600)
    # a stream deflated by someone else (any conforming raw-DEFLATE encoder, e.g. flate2) is readable
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    foreign = buf.getvalue()[:28] + co.compress((grid.size).to_bytes(8, "little") + grid.tobytes() +
                                                (600).to_bytes(8, "little")) + co.flush()
    assert hgi.Archive.deserialize_from_reader(io.BytesIO(foreign)).grid == hgi.Grid(grid, 600)


@pytest.mark.parametrize("block_rows", [None, 7, 64])
def test_archive_huffman_entropy_stage(block_rows):
    """Frequency tables in, bit-packed dynamic-Huffman DEFLATE out (hgi_archive_serialize_huffman).  Here the
    tables are counted with numpy; tests/test_gpu_next_rows.py feeds the GPU-built ones."""
    rng = np.random.default_rng(4)
    for h, w, levels, q in ((64, 48, 3, 2), (300, 200, 4, 3), (1, 1, 0, 0), (97, 13, 4, 0)):
        grid = oc.encode(rng.integers(0, 256, (h, w)).astype(np.uint8) // (1 if q == 0 else 6), levels, qlevel=q)
        buf = grid.reshape(-1)
        block = buf.size if not block_rows else block_rows * w
        nb = max(1, -(-buf.size // block))
        hist = np.stack([np.bincount(buf[b * block:(b + 1) * block], minlength=256) for b in range(nb)])
        md = hgi.Metadata(q, 0, w, h, levels)
        out = io.BytesIO()
        hgi.Archive(md, hgi.Grid(grid, w)).serialize_to_writer(out, entropy="huffman", hist=hist, block_rows=block_rows)
        raw = out.getvalue()
        assert raw[:4].hex() == "55a5adba"
        payload = zlib.decompress(raw[28:], -15)                  # a conforming inflate (flate2/miniz, zlib) reads it
        assert payload == buf.size.to_bytes(8, "little") + buf.tobytes() + w.to_bytes(8, "little")
        assert hgi.Archive.deserialize_from_reader(io.BytesIO(raw)) == hgi.Archive(md, hgi.Grid(grid, w))
    # a table that does not describe the block is rejected
    bad = hist.copy()
    bad[0, 0] += 1
    with pytest.raises(hgi.HgiError):
        hgi.Archive(md, hgi.Grid(grid, w)).serialize_to_writer(io.BytesIO(), entropy="huffman", hist=bad,
                                                                block_rows=block_rows)


@pytest.mark.parametrize("block_rows", [None, 32])
def test_archive_rle_entropy_stage(block_rows):
    """Token tables in (here from the CPU statement of the parse, tests/rle_model.py; the GPU-built ones are checked
    against it in tests/test_gpu_next_rows.py), literals + distance-1 matches out (hgi_archive_serialize_rle)."""
    from rle_model import rle_table
    rng = np.random.default_rng(14)
    cases = [(64, 48, 3, 2), (300, 200, 4, 3), (1, 1, 0, 0), (97, 13, 4, 0), (40, 1024, 4, 2), (3, 600, 2, 1)]
    for h, w, levels, q in cases:
        img = (rng.integers(0, 256, (h, w)) // (1 if q == 0 else 40) * 3).astype(np.uint8)   # long flat runs when quantized
        grid = oc.encode(img, levels, qlevel=q)
        buf = grid.reshape(-1)
        block = buf.size if not block_rows else -(-block_rows * w // 512) * 512
        hist = rle_table(buf, None if not block_rows else block)
        md = hgi.Metadata(q, 0, w, h, levels)
        out = io.BytesIO()
        hgi.Archive(md, hgi.Grid(grid, w)).serialize_to_writer(out, entropy="rle", hist=hist, block_rows=block_rows)
        raw = out.getvalue()
        assert raw[:4].hex() == "55a5adba"
        payload = zlib.decompress(raw[28:], -15)                  # a conforming inflate (flate2/miniz, zlib) reads it
        assert payload == buf.size.to_bytes(8, "little") + buf.tobytes() + w.to_bytes(8, "little")
        assert hgi.Archive.deserialize_from_reader(io.BytesIO(raw)) == hgi.Archive(md, hgi.Grid(grid, w))
    # runs of every length around the 258 / 512-byte limits, and a table that does not belong to the data
    data = np.concatenate([np.full(n, 7 + (i & 1), np.uint8) for i, n in enumerate([1, 2, 3, 4, 257, 258, 259, 260, 261, 516, 1100, 5])])
    out = io.BytesIO()
    hgi.Archive(hgi.Metadata(0, 0, data.size, 1, 0), hgi.Grid(data, data.size)).serialize_to_writer(out, entropy="rle", hist=rle_table(data))
    assert zlib.decompress(out.getvalue()[28:], -15)[8:-8] == data.tobytes()
    bad = rle_table(data)
    bad[0, 285] = 0                                               # the 258-byte matches lose their code
    with pytest.raises(hgi.HgiError):
        hgi.Archive(hgi.Metadata(0, 0, data.size, 1, 0), hgi.Grid(data, data.size)).serialize_to_writer(io.BytesIO(), entropy="rle", hist=bad)


def test_archive_rle_is_as_small_as_zlib9_on_a_photograph():
    """fullhd Medium within 10 % of zlib level 9 (<= 250 KB)."""
    from conftest import get_plane
    from rle_model import rle_table
    grid = oc.encode(get_plane("fullhd"), 4, qlevel=2)
    payload = grid.size.to_bytes(8, "little") + grid.tobytes() + (1920).to_bytes(8, "little")
    out = io.BytesIO()
    hgi.Archive(hgi.Metadata(2, 0, 1920, 1080, 4), hgi.Grid(grid, 1920)).serialize_to_writer(out, entropy="rle", hist=rle_table(grid))
    raw = out.getvalue()
    assert zlib.decompress(raw[28:], -15) == payload
    z9 = zlib.compressobj(9, zlib.DEFLATED, -15)
    z9 = len(z9.compress(payload) + z9.flush())
    assert len(raw) <= 250 * 1024 and len(raw) <= 1.10 * z9, (len(raw), z9)


def test_archive_reader_refuses_dishonest_length_prefixes():
    """the u64 bincode length prefix is validated before anybody allocates for it; an empty grid reads back."""
    co = zlib.compressobj(9, zlib.DEFLATED, -15)
    lie = co.compress((1 << 39).to_bytes(8, "little") + bytes(32)) + co.flush()
    head = bytes.fromhex("55a5adba") + (0).to_bytes(4, "little") * 2 + (4).to_bytes(4, "little") * 2 + (1).to_bytes(8, "little")
    with pytest.raises(hgi.HgiError) as e:
        hgi.Archive.deserialize_from_reader(io.BytesIO(head + lie))
    assert e.value.status == -6
    empty = io.BytesIO()
    hgi.Archive(hgi.Metadata(0, 0, 0, 0, 1), hgi.Grid(np.zeros(0, np.uint8), 0)).serialize_to_writer(empty)
    back = hgi.Archive.deserialize_from_reader(io.BytesIO(empty.getvalue()))
    assert back.grid.buffer.size == 0 and back.metadata.width == 0


def test_archive_huffman_skewed_tables_respect_length_limit():
    """Fibonacci-like counts would give code lengths > 15 without the length limit."""
    counts = [1, 1]
    while len(counts) < 40:
        counts.append(counts[-1] + counts[-2])
    data = np.concatenate([np.full(c, i, np.uint8) for i, c in enumerate(counts[:30])])
    hist = np.bincount(data, minlength=256)[None, :]
    out = io.BytesIO()
    hgi.Archive(hgi.Metadata(0, 0, data.size, 1, 0), hgi.Grid(data, data.size)).serialize_to_writer(
        out, entropy="huffman", hist=hist)
    assert zlib.decompress(out.getvalue()[28:], -15)[8:-8] == data.tobytes()


def test_archive_errors():
    with pytest.raises(hgi.HgiError) as e:
        hgi.Archive.deserialize_from_reader(io.BytesIO(b"\x00\x01\x02\x03" + bytes(40)))
    assert e.value.status == -5 and "incorrect magic number" in str(e.value)   # src/archive.rs:47-50
    good = io.BytesIO()
    hgi.Archive(hgi.Metadata(0, 0, 4, 4, 1), hgi.Grid(np.arange(16, dtype=np.uint8), 4)).serialize_to_writer(good)
    with pytest.raises(hgi.HgiError) as e:
        hgi.Archive.deserialize_from_reader(io.BytesIO(good.getvalue()[:-3]))
    assert e.value.status == -6
    with pytest.raises(hgi.HgiError):
        hgi.Archive.deserialize_from_reader(io.BytesIO(good.getvalue()[:10]))


def test_no_cpu_fallback_and_argument_validation():
    L = hgi.lib()
    h = ctypes.c_void_p()
    rc = L.hgi_ctx_create(0, ctypes.byref(h))
    if rc == 0:
        L.hgi_ctx_destroy(h)
    else:
        assert rc == -2 and b"no CPU fallback" in L.hgi_strerror(rc)    # fails loudly without a GPU
    assert L.hgi_ctx_create(0, None) == -1
    p = hgi._lib.Params(4, 0, 1, 2)
    buf = np.zeros(16, np.uint8)
    assert L.hgi_encode_u8(None, buf.ctypes.data, 4, 4, ctypes.byref(p), buf.ctypes.data, None) == -1
    assert L.hgi_quant_table(7, 0, buf.ctypes.data, None) == -1
    assert L.hgi_quant_table(1, 9, buf.ctypes.data, None) == -1


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "rustyhgi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "hgi_oracle" not in open(os.path.join(ROOT, "include", f)).read()
