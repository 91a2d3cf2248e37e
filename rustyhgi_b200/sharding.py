"""Host-side partitioning of the HGI path across the GPUs of one box.

Two modes, neither needs a collective on the data path:
  * by image: a batch is cut into contiguous ranges, one per rank;
  * by row band: a huge plane is cut into bands whose heights are multiples of S = 2^levels; a
    band [y0, y1) is bit-exact when computed from input rows [y0, min(h, y1 + S + 1)) because
    dependencies only point right/down (src/interpolator.rs:70-73) -- the overlap rows are
    recomputed by the rank above instead of being exchanged.
"""
from dataclasses import dataclass


def split_batch(n_images, world_size, rank):
    """Contiguous [first, last) image range of `rank`."""
    base, extra = divmod(n_images, world_size)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Band:
    y0: int        # first output row
    y1: int        # one past the last output row
    in_y1: int     # one past the last input row needed (y1 + S + 1, clamped to the plane)

    @property
    def rows_out(self):
        return self.y1 - self.y0

    @property
    def rows_in(self):
        return self.in_y1 - self.y0


def plan_bands(height, levels, n_bands):
    """Cut `height` rows into <= n_bands bands aligned to S = 2^levels (empty bands are dropped)."""
    S = 1 << levels
    cells = -(-height // S)
    bands = []
    for r in range(n_bands):
        c0, c1 = split_batch(cells, n_bands, r)
        y0, y1 = min(c0 * S, height), min(c1 * S, height)
        if y1 > y0:
            bands.append(Band(y0, y1, min(height, y1 + S + 1)))
    return bands
