"""ctypes binding of oracle/libhgi_oracle.so (see hgi_oracle.c; test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhgi_oracle.so")

INTERP_CROSSED, INTERP_LEFTTOP = 0, 3
QUANT_NOOP, QUANT_LINEAR = 0, 1
LOSSLESS, LOW, MEDIUM, HIGH = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "hgi_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libhgi_oracle.so"])
    return _SO


_lib = None
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.hgi_oracle_encode_ex.argtypes = [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           _u8p, _u8p, ctypes.POINTER(ctypes.c_uint64)]
        L.hgi_oracle_decode_ex.argtypes = [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_int, ctypes.c_int, _u8p]
        L.hgi_oracle_quant_table.argtypes = [ctypes.c_int, ctypes.c_int, _u8p, _u8p]
        L.hgi_oracle_quant_table.restype = None
        L.hgi_oracle_histogram.argtypes = [_u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
        L.hgi_oracle_histogram.restype = None
        L.hgi_oracle_sd.argtypes = [_u8p, _u8p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64),
                                    ctypes.POINTER(ctypes.c_uint32)]
        L.hgi_oracle_sd.restype = ctypes.c_uint64
        L.hgi_oracle_rgb_to_luma.argtypes = [_u8p, ctypes.c_size_t, _u8p]
        L.hgi_oracle_rgb_to_luma.restype = None
        L.hgi_oracle_encode_batch.argtypes = [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              _u8p, ctypes.c_int]
        L.hgi_oracle_decode_batch.argtypes = [_u8p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_int, _u8p, ctypes.c_int]
        L.hgi_oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_u8p)


def _plane(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def quant_table(qkind=QUANT_LINEAR, qlevel=MEDIUM):
    t = np.empty(256, np.uint8)
    e = ctypes.c_uint8(0)
    lib().hgi_oracle_quant_table(qkind, qlevel, _p(t), ctypes.cast(ctypes.byref(e), _u8p))
    return t, e.value


def encode(image, levels, interp=INTERP_CROSSED, qkind=QUANT_LINEAR, qlevel=MEDIUM,
           legacy_round=False, want_recon=False, want_fixups=False):
    image = _plane(image)
    h, w = image.shape
    grid = np.empty_like(image)
    recon = np.empty_like(image)
    fix = ctypes.c_uint64(0)
    rc = lib().hgi_oracle_encode_ex(_p(image), w, h, levels, interp, qkind, qlevel,
                                    int(legacy_round), _p(grid), _p(recon), ctypes.byref(fix))
    if rc:
        raise ValueError(f"hgi_oracle_encode_ex -> {rc}")
    out = [grid]
    if want_recon:
        out.append(recon)
    if want_fixups:
        out.append(fix.value)
    return out[0] if len(out) == 1 else tuple(out)


def decode(grid, levels, interp=INTERP_CROSSED, legacy_round=False):
    grid = _plane(grid)
    h, w = grid.shape
    img = np.empty_like(grid)
    rc = lib().hgi_oracle_decode_ex(_p(grid), w, h, levels, interp, int(legacy_round), _p(img))
    if rc:
        raise ValueError(f"hgi_oracle_decode_ex -> {rc}")
    return img


def histogram(grid):
    g = np.ascontiguousarray(grid, dtype=np.uint8).reshape(-1)
    hist = np.zeros(256, np.uint64)
    lib().hgi_oracle_histogram(_p(g), g.size, hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    return hist


def sd(before, after):
    """`hgi test` numbers (src/main.rs:84-111): (sd_int, sum_sq, max_abs)."""
    b = np.ascontiguousarray(before, dtype=np.uint8).reshape(-1)
    a = np.ascontiguousarray(after, dtype=np.uint8).reshape(-1)
    s = ctypes.c_uint64(0)
    m = ctypes.c_uint32(0)
    v = lib().hgi_oracle_sd(_p(b), _p(a), b.size, ctypes.byref(s), ctypes.byref(m))
    return int(v), int(s.value), int(m.value)


def rgb_to_luma(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    assert rgb.ndim == 3 and rgb.shape[2] == 3
    out = np.empty(rgb.shape[:2], np.uint8)
    lib().hgi_oracle_rgb_to_luma(_p(rgb.reshape(-1)), out.size, _p(out))
    return out


def encode_batch(images, levels, interp=INTERP_CROSSED, qkind=QUANT_LINEAR, qlevel=MEDIUM, n_threads=0):
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, h, w = images.shape
    grids = np.empty_like(images)
    rc = lib().hgi_oracle_encode_batch(_p(images), n, w, h, levels, interp, qkind, qlevel, _p(grids), n_threads)
    if rc:
        raise ValueError(rc)
    return grids


def decode_batch(grids, levels, interp=INTERP_CROSSED, n_threads=0):
    grids = np.ascontiguousarray(grids, dtype=np.uint8)
    n, h, w = grids.shape
    out = np.empty_like(grids)
    rc = lib().hgi_oracle_decode_batch(_p(grids), n, w, h, levels, interp, _p(out), n_threads)
    if rc:
        raise ValueError(rc)
    return out


def max_threads():
    return lib().hgi_oracle_max_threads()
