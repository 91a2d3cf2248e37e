"""CPU statement of the run-length DEFLATE parse shared by hgi_rle_hist_kernel (GPU table) and
hgi_archive_serialize_rle (host bit-packer): 512-byte segments from the block start; per maximal run of a byte b:
one literal b, matches of min(258, rest) while rest >= 3, then the remaining 0..2 bytes as literals."""
import numpy as np

SEG, SYMS = 512, 288


def len_sym(length):
    if length == 258:
        return 285
    l = length - 3
    e = 0 if l < 8 else l.bit_length() - 3
    return 257 + 4 * e + (l >> e)


def rle_table(data, block_bytes=None):
    """(n_blocks, 288) token frequencies of `data` (1-D uint8), blocks of `block_bytes` (None: one block)."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    n = data.size
    block = n if not block_bytes else int(block_bytes)
    n_blocks = max(1, -(-n // max(block, 1)))
    hist = np.zeros((n_blocks, SYMS), np.uint32)
    for b in range(n_blocks):
        blk = data[b * block:(b + 1) * block]
        for s0 in range(0, blk.size, SEG):
            seg = blk[s0:s0 + SEG]
            # run boundaries inside the segment
            starts = np.flatnonzero(np.concatenate(([True], seg[1:] != seg[:-1])))
            lens = np.diff(np.concatenate((starts, [seg.size])))
            for st, ln in zip(starts, lens):
                v, rem = int(seg[st]), int(ln) - 1
                hist[b, v] += 1
                while rem >= 3:
                    m = min(rem, 258)
                    hist[b, len_sym(m)] += 1
                    rem -= m
                hist[b, v] += rem
    return hist
