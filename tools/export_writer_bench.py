#!/usr/bin/env python
"""export_writer_bench.py -- host cost of packaging one decoded frame into a DNG: this repo's writer
(include/motioncam/Export.hpp, one vectored write straight from the decoded buffer) next to the reference's
(example.cpp:55-139 on thirdparty/tinydng, through oracle/_ref; only when that library is present).
Files go to --dir (default /dev/shm, so the page cache / tmpfs copy is all the I/O there is)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=4080)
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    from motioncam_decoder_b200 import hostapi, testvec as tv
    import oracle_lib as ol
    px = np.random.default_rng(1).integers(0, 1024, (a.height, a.width), dtype=np.uint16)
    fm = {"width": a.width, "height": a.height, "asShotNeutral": [0.5, 1.0, 0.6]}
    cm = tv.DEFAULT_CONTAINER_METADATA
    arms = [("b200_writer", hostapi.library(), "mcb200_")]
    if ol.have_ref():
        arms.append(("reference_writer", hostapi.library(ol.REF_SO, "mcref_"), "mcref_"))
    out = {"frame": f"{a.width}x{a.height}", "bytes": int(px.nbytes), "dir": a.dir}
    for name, lib, prefix in arms:
        path = os.path.join(a.dir, f"_bench_{name}.dng")
        best = 1e9
        for _ in range(a.iters):
            t0 = time.perf_counter()
            hostapi.write_dng(path, px, fm, cm, lib=lib, prefix=prefix)
            best = min(best, time.perf_counter() - t0)
        os.unlink(path)
        out[name] = {"ms_per_frame": round(best * 1e3, 3), "gb_per_s": round(px.nbytes / best / 1e9, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
