"""Independent numpy restatement of the HGI encode/decode semantics -- TEST INFRASTRUCTURE ONLY.

Formulated per *cell* and vectorised per level, i.e. deliberately not the
per-pixel traversal that oracle/hgi_oracle.c copies from the reference, so that agreement of the
two is evidence for both.  References: src/encoder.rs:39-71, src/decoder.rs:18-46,
src/utils.rs:11-41, src/interpolator.rs:15-28,41-91, src/quantizator.rs:41-63.
"""
import numpy as np

INTERP_CROSSED, INTERP_LEFTTOP = 0, 3


def quant_table(error):
    """src/quantizator.rs:50-60."""
    scale = 2 * error + 1
    x = np.arange(256)
    return (((x + error) // scale) * scale).astype(np.uint8)


LEVEL_ERRORS = (0, 10, 20, 30)  # src/quantizator.rs:43-48


def _corners(R, step):
    """Corner planes A(x0,y0) B(x0,y1) C(x1,y0) D(x1,y1) per coarse cell, OOB -> 0
    (src/interpolator.rs:70-88)."""
    h, w = R.shape
    lat = R[::step, ::step].astype(np.int64)
    ch, cw = lat.shape
    pad = np.zeros((ch + 1, cw + 1), np.int64)
    pad[:ch, :cw] = lat
    return pad[:ch, :cw], pad[1:, :cw], pad[:ch, 1:], pad[1:, 1:]


def _prediction(R, step, interp, legacy_round):
    A, B, C, D = _corners(R, step)
    if interp == INTERP_LEFTTOP:
        return A
    avg = lambda p, q: (p + q + 1) >> 1                      # src/interpolator.rs:44
    s = avg(A, B) + avg(D, C) + avg(C, A) + avg(D, B)        # :46-49 (left,right,top,bot)
    return (s + 1) >> 2 if legacy_round else s >> 2          # :51


def _new_point_slices(step):
    sub = step // 2
    # (dy, dx) offsets of the three new points of a cell (src/utils.rs:19-38)
    return [(0, sub), (sub, 0), (sub, sub)]


def encode(image, levels, interp=INTERP_CROSSED, table=None, legacy_round=False):
    image = np.asarray(image, np.uint8)
    h, w = image.shape
    if table is None:
        table = np.arange(256, dtype=np.uint8)
    R = image.astype(np.int64).copy()
    G = np.zeros((h, w), np.int64)
    S = 1 << levels
    G[::S, ::S] = R[::S, ::S]                                 # src/encoder.rs:26-37
    fixups = 0
    for level in range(levels):
        step = 1 << (levels - level)
        P = _prediction(R, step, interp, legacy_round)
        for dy, dx in _new_point_slices(step):
            a = R[dy::step, dx::step]
            if a.size == 0:
                continue
            p = P[:a.shape[0], :a.shape[1]]
            d = (a - p) & 255                                 # src/encoder.rs:53
            q = table[d].astype(np.int64)                     # :54
            fix = ((p + q) > 255) != ((p + d) > 255)          # :56-58
            q = np.where(fix, d, q)                           # :59
            fixups += int(fix.sum())
            G[dy::step, dx::step] = q                         # :62
            R[dy::step, dx::step] = (p + q) & 255             # :63-64
    return G.astype(np.uint8), R.astype(np.uint8), fixups


def decode(grid, levels, interp=INTERP_CROSSED, legacy_round=False):
    grid = np.asarray(grid, np.uint8)
    h, w = grid.shape
    G = grid.astype(np.int64)
    R = np.zeros((h, w), np.int64)                            # src/decoder.rs:19
    S = 1 << levels
    R[::S, ::S] = G[::S, ::S]                                 # :22-28
    for level in range(levels):
        step = 1 << (levels - level)
        P = _prediction(R, step, interp, legacy_round)
        for dy, dx in _new_point_slices(step):
            g = G[dy::step, dx::step]
            if g.size == 0:
                continue
            p = P[:g.shape[0], :g.shape[1]]
            R[dy::step, dx::step] = (p + g) & 255             # :39
    return R.astype(np.uint8)


def scalar_encode(image, levels, table, interp=INTERP_CROSSED):
    """Pure-Python per-pixel loop in the reference's exact visiting order (small cases only)."""
    h, w = len(image), len(image[0])
    R = [list(map(int, row)) for row in image]
    G = [[0] * w for _ in range(h)]
    S = 1 << levels
    for y in range(0, h, S):
        for x in range(0, w, S):
            G[y][x] = R[y][x]

    def px(x, y):
        return R[y][x] if x < w and y < h else 0

    def pred(x, y, step):
        x0, y0 = x - (x & (step - 1)), y - (y & (step - 1))
        if interp == INTERP_LEFTTOP:
            return R[y0][x0]
        x1, y1 = x0 + step, y0 + step
        lt, rt, lb, rb = px(x0, y0), px(x0, y1), px(x1, y0), px(x1, y1)
        left, right = (lt + lb + 1) >> 1, (rb + rt + 1) >> 1
        top, bot = (rt + lt + 1) >> 1, (rb + lb + 1) >> 1
        return ((left + right + top + bot) >> 2) & 255

    def visit(x, y, step):
        p = pred(x, y, step)
        a = R[y][x]
        d = (a - p) & 255
        q = int(table[d])
        if ((p + q) > 255) != ((p + d) > 255):
            q = d
        G[y][x] = q
        R[y][x] = (p + q) & 255

    for level in range(levels):
        e = levels - level
        step, sub = 1 << e, 1 << (e - 1)
        y = 0
        while y < h:
            for x in range(sub, w, step):
                visit(x, y, step)
            y += sub
            if y >= h:
                break
            for x in range(0, w, sub):
                visit(x, y, step)
            y += sub
    return np.array(G, np.uint8), np.array(R, np.uint8)
