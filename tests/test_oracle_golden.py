"""Pins the oracle (oracle/hgi_oracle.c) -- CPU only.

1. the reference's one golden artefact (docs pair) bit-exactly;
2. SURVEY.md Appendix B.2 fingerprints / the 12x8 KAT;
3. an independent numpy restatement and a pure-Python scalar loop;
4. the properties the reference's own (vacuous, src/lib.rs:61) tests meant to check.
"""
import numpy as np
import pytest

from conftest import get_plane, load_plane, sha16, photo_like
from oracle import c as oc
from oracle import pyref


def test_docs_golden_pair_bit_exact():
    """README.md:8 lena_source.png -> lena_hgi.png is Low, L=4, Crossed with the artefact-era
    final rounding; every other stage (traversal, OOB=0, LUT, fix-up, wrap, decode) is HEAD's."""
    src, want = load_plane("docs_lena_source"), load_plane("docs_lena_hgi")
    grid, recon = oc.encode(src, 4, qlevel=oc.LOW, legacy_round=True, want_recon=True)
    assert int((recon != want).sum()) == 0
    assert int((oc.decode(grid, 4, legacy_round=True) != want).sum()) == 0
    assert (src[::16, ::16] == want[::16, ::16]).all()          # 625 raw lattice pixels
    assert int(np.abs(src.astype(int) - want).max()) == 10      # Low's bound
    # HEAD rounding differs from the artefact in exactly the documented way (47 047 px)
    _, head = oc.encode(src, 4, qlevel=oc.LOW, want_recon=True)
    assert int((head != want).sum()) == 47047


KAT_12x8_MEDIUM = np.array([
    [0, 0, 0, 0, 0, 254, 246, 251, 0, 251, 246, 254],
    [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, 0, 0, 0, 246, 0, 0, 246, 0, 246, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 0, 246, 246, 0, 0],
    [0, 0, 0, 0, 0, 0, 0, 246, 41, 246, 41, 0],
    [0, 0, 0, 0, 0, 0, 246, 246, 246, 246, 41, 41],
    [0, 0, 0, 0, 0, 0, 41, 0, 41, 41, 41, 41],
    [0, 0, 0, 0, 0, 0, 0, 41, 41, 41, 41, 82]], np.uint8)


def test_kat_12x8_medium():
    img = get_plane("unit_12x8")
    assert (oc.encode(img, 3, qlevel=oc.MEDIUM) == KAT_12x8_MEDIUM).all()


SURVEY_B2 = [  # plane, L, q, grid sha, recon sha, max err, fixups  
    ("lena_tif", 4, 0, "ad84562f0403d27a", "f6a26c7641342ed5", 0, 0),
    ("lena_tif", 4, 1, "d00582ba34c73691", "009eca639238603e", 10, 39),
    ("lena_tif", 4, 2, "3a992020370c96a4", "e17f5ad9f400234e", 20, 235),
    ("lena_tif", 4, 3, "6efd1ae27bde1fa5", "1db1b14b159b18a8", 30, 425),
    ("bench_1080p", 4, 0, "0dbac6b86fd3ad08", "7acb2c3ac8a24dd3", 0, 0),
    ("bench_1080p", 4, 2, "93a4da916d2bd2d3", "6ff3684c90d7668c", 20, 73311),
    ("fullhd", 4, 0, "b7406d033c56ee64", "a0d0d3ec8c0363a0", 0, 0),
    ("fullhd", 4, 1, "b160d629f824f145", "c29ace6a2d806c8d", 10, 9510),
    ("fullhd", 4, 2, "b83a38befd476f30", "96db2c544587bccb", 20, 22687),
    ("fullhd", 4, 3, "46147dc7d7d05eb7", "a5b5cec5a93f0819", 30, 26219),
    ("ikonos", 6, 3, "841cb88af9f6c933", "36856863b3a5da03", 30, 59822),
    ("unit_12x8", 3, 0, "38aa39a03be8578c", "7543b9f9570cdd2c", 0, 0),
    ("unit_12x8", 3, 1, "6e51b00b6e83e1d6", "c167f995fa7c6fb4", 10, 8),
    ("unit_12x8", 3, 2, "947ae38b5c870a01", "b34c935d0d8eb4bc", 20, 4),
    ("unit_12x8", 3, 3, "ba1384f0037878fc", "9bdb839d268a27da", 29, 5),
    ("unit_8x8", 3, 0, "b8ba79738f34dc15", "d2404a4add96c0ea", 0, 0),
]


@pytest.mark.parametrize("plane,levels,q,gsha,rsha,maxerr,fixups", SURVEY_B2)
def test_survey_fingerprints(plane, levels, q, gsha, rsha, maxerr, fixups):
    img = get_plane(plane)
    g, r, f = oc.encode(img, levels, qlevel=q, want_recon=True, want_fixups=True)
    assert sha16(g) == gsha and sha16(r) == rsha and f == fixups
    assert int(np.abs(img.astype(int) - r).max()) == maxerr
    assert (oc.decode(g, levels) == r).all()


def test_golden_json_matches_oracle(golden):
    for name, meta in golden["planes"].items():
        assert sha16(get_plane(name)) == meta["sha"], name
    for cs in golden["cases"]:
        if cs["plane"] in ("fullhd", "ikonos") and cs["interp"] != 0:
            continue  # keep the CPU suite short; the GPU suite covers these
        img = get_plane(cs["plane"])
        g, r = oc.encode(img, cs["levels"], interp=cs["interp"], qlevel=cs["qlevel"], want_recon=True)
        assert sha16(g) == cs["grid_sha"] and sha16(r) == cs["recon_sha"], cs


def test_hgi_test_numbers_config1():
    """BASELINE config 1: `hgi test res/LENA.TIF` level 4 Medium -> Uncompressed 64 kb, SD 9.17
    (src/main.rs:84-111)."""
    img = load_plane("lena_tif")
    g = oc.encode(img, 4, qlevel=oc.MEDIUM)
    after = oc.decode(g, 4)
    sd_int, _, mx = oc.sd(img, after)
    assert img.size // 1024 == 64
    assert f"{np.sqrt(float(sd_int)):.2f}" == "9.17"
    assert mx == 20


@pytest.mark.parametrize("w,h,levels", [(12, 8, 3), (8, 8, 3), (1, 1, 4), (1, 9, 2), (9, 1, 2), (37, 23, 5),
                                        (64, 64, 6), (250, 243, 4), (33, 17, 0), (20, 20, 12)])
@pytest.mark.parametrize("q", [0, 1, 2, 3])
def test_c_oracle_vs_numpy_restatement(w, h, levels, q):
    img = photo_like(w, h, seed=w * 131 + h)
    table, err = oc.quant_table(oc.QUANT_LINEAR, q)
    assert (table == pyref.quant_table(pyref.LEVEL_ERRORS[q])).all()
    for interp in (oc.INTERP_CROSSED, oc.INTERP_LEFTTOP):
        g, r, f = oc.encode(img, levels, interp=interp, qlevel=q, want_recon=True, want_fixups=True)
        g2, r2, f2 = pyref.encode(img, levels, interp=interp, table=table)
        assert (g == g2).all() and (r == r2).all() and f == f2
        assert (oc.decode(g, levels, interp=interp) == r).all()
        assert (pyref.decode(g, levels, interp=interp) == r).all()
        assert int(np.abs(img.astype(int) - r).max()) <= err     # the non-vacuous src/lib.rs:71-76
        if w * h <= 600:
            g3, r3 = pyref.scalar_encode(img.tolist(), levels, table, interp=interp)
            assert (g == g3).all() and (r == r3).all()


def test_noop_equals_lossless_linear():
    img = photo_like(61, 47, 5)
    a = oc.encode(img, 4, qkind=oc.QUANT_NOOP, qlevel=2)
    b = oc.encode(img, 4, qkind=oc.QUANT_LINEAR, qlevel=0)
    assert (a == b).all()


def test_band_halo_property():
    """a band [y0,y1) needs input rows [y0, y1+S+1) and nothing above."""
    img = load_plane("lena_tif")
    L, S = 4, 16
    full_g, full_r = oc.encode(img, L, qlevel=2, want_recon=True)
    y0, y1 = 64, 128
    g, r = oc.encode(img[y0:y1 + S + 1], L, qlevel=2, want_recon=True)
    assert (g[:y1 - y0] == full_g[y0:y1]).all() and (r[:y1 - y0] == full_r[y0:y1]).all()
    g16 = oc.encode(img[y0:y1 + S], L, qlevel=2)
    assert (g16[:y1 - y0] != full_g[y0:y1]).any()               # S rows are not enough
    d = oc.decode(full_g[y0:y1 + S + 1], L)
    assert (d[:y1 - y0] == full_r[y0:y1]).all()


def test_self_similarity():
    """Coarse levels on the full plane == the whole codec on the decimated plane (B.4)."""
    img = photo_like(250, 243, 9)
    g, r = oc.encode(img, 5, qlevel=3, want_recon=True)
    for lf in (1, 2, 3):
        gd, rd = oc.encode(img[::1 << lf, ::1 << lf], 5 - lf, qlevel=3, want_recon=True)
        assert (gd == g[::1 << lf, ::1 << lf]).all() and (rd == r[::1 << lf, ::1 << lf]).all()


def test_histogram_and_levels_zero():
    img = photo_like(40, 30, 3)
    g = oc.encode(img, 0, qlevel=2)
    assert (g == img).all() and (oc.decode(g, 0) == img).all()
    assert (oc.histogram(g) == np.bincount(g.reshape(-1), minlength=256)).all()


def test_batch_driver_matches_single():
    imgs = np.stack([photo_like(48, 32, s) for s in range(5)])
    grids = oc.encode_batch(imgs, 3, qlevel=1, n_threads=3)
    for i in range(5):
        assert (grids[i] == oc.encode(imgs[i], 3, qlevel=1)).all()
    back = oc.decode_batch(grids, 3, n_threads=2)
    assert int(np.abs(back.astype(int) - imgs).max()) <= 10
