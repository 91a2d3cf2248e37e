#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference tree (run in the dev container only:
`python tests/golden/make_golden.py`; /root/reference does not exist on the GPU box).

Inputs are *data* files of the reference (res/*, docs/static_files/*), decoded to the raw u8
luma plane that the codec consumes (PIL decode -> f32 luma, the restatement of
image-0.19's `to_luma` that src/main.rs:42,74 calls) and stored losslessly as 8-bit gray PNG.
`golden.json` holds sha256[:16] fingerprints of every plane plus grid / reconstruction
fingerprints, `hgi test` numbers and fix-up counts produced by the C oracle at HEAD semantics.

The only artefact that pins the *oracle itself* is the docs pair (lena_source -> lena_hgi);
tests/test_oracle_golden.py asserts it bit-exactly (Low, L=4, Crossed, legacy final rounding).
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import c as oc  # noqa: E402

REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def luma_plane(path):
    im = Image.open(path)
    if im.mode == "L":
        return np.array(im)
    rgb = np.array(im.convert("RGB"))
    return oc.rgb_to_luma(rgb)


def main():
    planes = {
        "lena_tif": luma_plane(f"{REF}/res/LENA.TIF"),
        "fullhd": luma_plane(f"{REF}/res/fullhd.jpg"),
        "ikonos": luma_plane(f"{REF}/res/ikonos-barcelona-spain.jpg"),
        "docs_lena_source": luma_plane(f"{REF}/docs/static_files/lena_source.png"),
        "docs_lena_hgi": np.array(Image.open(f"{REF}/docs/static_files/lena_hgi.png").convert("L")),
    }
    meta = {"planes": {}, "cases": []}
    for name, p in planes.items():
        Image.fromarray(p, "L").save(os.path.join(HERE, name + ".png"), optimize=True)
        back = np.array(Image.open(os.path.join(HERE, name + ".png")))
        assert back.dtype == np.uint8 and (back == p).all()
        meta["planes"][name] = {"width": int(p.shape[1]), "height": int(p.shape[0]), "sha": sha(p)}

    yy, xx = np.mgrid[0:1080, 0:1920]
    bench = ((xx * yy) & 255).astype(np.uint8)          # benches/bench.rs:24-28
    yy, xx = np.mgrid[0:8, 0:12]
    unit = ((xx * yy) & 255).astype(np.uint8)           # src/lib.rs:36-43 at 12x8
    yy, xx = np.mgrid[0:8, 0:8]
    unit8 = ((xx * yy) & 255).astype(np.uint8)          # src/lib.rs:103
    planes.update(bench_1080p=bench, unit_12x8=unit, unit_8x8=unit8)
    for name in ("bench_1080p", "unit_12x8", "unit_8x8"):
        p = planes[name]
        meta["planes"][name] = {"width": int(p.shape[1]), "height": int(p.shape[0]), "sha": sha(p),
                                "generator": "(x*y)&255"}

    todo = [("lena_tif", 4, q) for q in range(4)] + [("fullhd", 4, q) for q in range(4)] + \
           [("ikonos", 6, 3)] + [("bench_1080p", 4, 0), ("bench_1080p", 4, 2)] + \
           [("unit_12x8", 3, q) for q in range(4)] + [("unit_8x8", 3, 0)]
    for name, levels, q in todo:
        src = planes[name]
        for interp in (oc.INTERP_CROSSED, oc.INTERP_LEFTTOP):
            g, r, fix = oc.encode(src, levels, interp=interp, qlevel=q, want_recon=True, want_fixups=True)
            d = oc.decode(g, levels, interp=interp)
            assert (d == r).all()
            sd, ssq, mx = oc.sd(src, d)
            meta["cases"].append({"plane": name, "levels": levels, "qlevel": q, "interp": interp,
                                  "grid_sha": sha(g), "recon_sha": sha(r), "fixups": int(fix),
                                  "max_err": mx, "sd_int": sd, "sum_sq": ssq,
                                  "distinct_symbols": int(len(np.unique(g)))})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta["planes"], indent=1))


if __name__ == "__main__":
    main()
