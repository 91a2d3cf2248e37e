"""The pass/tile decomposition of the CUDA library, executed on the CPU (tests/tile_model.py),
must reproduce the oracle bit for bit -- including with every unstaged shared-memory byte
poisoned.  This checks the halo geometry (need_limit, staged rows, coarse fill) and the pass
plan without a GPU."""
import numpy as np
import pytest

import tile_model as tm
from conftest import photo_like
from oracle import c as oc


@pytest.mark.parametrize("w,h,levels,q", [(150, 70, 4, 2), (129, 65, 1, 3), (257, 129, 8, 3), (12, 8, 3, 2),
                                          (1, 1, 5, 2), (5, 1, 2, 1), (260, 70, 5, 2), (131, 67, 2, 1)])
def test_tile_model_matches_oracle(w, h, levels, q):
    img = photo_like(w, h, seed=w + 7 * h)
    table, _ = oc.quant_table(oc.QUANT_LINEAR, q)
    for crossed in (True, False):
        interp = oc.INTERP_CROSSED if crossed else oc.INTERP_LEFTTOP
        g, r = oc.encode(img, levels, interp=interp, qlevel=q, want_recon=True)
        g2, r2 = tm.encode(img, levels, table, crossed)
        assert (g == g2).all() and (r == r2).all()
        assert (tm.decode(g, levels, crossed) == r).all()


def test_small_tiles_many_passes():
    """Shrunken tiles (32x16) force many tiles and all three pass depths on a small plane."""
    img = photo_like(200, 90, 11)
    table, _ = oc.quant_table(oc.QUANT_LINEAR, 2)
    for levels in (1, 4, 6, 9):
        g, r = oc.encode(img, levels, qlevel=2, want_recon=True)
        g2, r2 = tm.encode(img, levels, table, True, tw=32, th=16)
        assert (g == g2).all() and (r == r2).all()
        assert (tm.decode(g, levels, True, tw=32, th=16) == r).all()


def test_pass_plan():
    assert tm.plan_passes(4) == [(0, 4)]
    assert tm.plan_passes(6) == [(4, 2), (0, 4)]
    assert tm.plan_passes(8) == [(4, 4), (0, 4)]
    assert tm.plan_passes(9) == [(8, 1), (4, 4), (0, 4)]
    assert tm.effective_levels(31, 1920, 1080) == 11
    assert tm.effective_levels(4, 1, 1) == 0
