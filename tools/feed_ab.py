#!/usr/bin/env python
"""feed_ab.py -- file -> device with the two feeds of Decoder::loadFramesToDevice, torch-free, one GPU:
pread into the pinned ring (default) against MCRAW_FEED=mmap (page-locked read-only mapping, H2D straight from the
page cache).  Synthetic C3-sized clip in tmpfs; outputs spot-checked.

    python tools/feed_ab.py [--frames 32] [--reps 4]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from motioncam_decoder_b200 import capi, hostapi, testvec as tv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    w, h = 4080, 3072
    images = [tv.gen_flatnoise(w, h, 256, seed=s + 1) for s in range(2)]
    streams = [tv.encode_current(im) for im in images]
    path = os.path.join(a.dir, "mcraw_feed_ab.mcraw")
    tv.write_mcraw(path, [{"timestamp": 1000 + i, "data": streams[i % 2], "width": w, "height": h, "compressionType": 7}
                          for i in range(a.frames)], [])
    ctx = capi.Context(0)
    ptrs = [ctx.device_alloc(w * h * 2) for _ in range(a.frames)]
    caps = [w * h] * a.frames
    out = {"frames": a.frames, "frame": f"{w}x{h}", "file_bytes": os.path.getsize(path)}
    buf = np.empty((h, w), np.uint16)
    for mode in ("pread", "mmap"):
        os.environ["MCRAW_FEED"] = mode
        t0 = time.perf_counter()
        dec = hostapi.Decoder(path)
        stamps = dec.get_frames()
        dec.load_frames_to_device(stamps, ptrs, caps)                 # first call: ring allocation / mapping + registration
        first = time.perf_counter() - t0
        best = 1e9
        for _ in range(a.reps):
            t0 = time.perf_counter()
            dec.load_frames_to_device(stamps, ptrs, caps)
            best = min(best, time.perf_counter() - t0)
        ok = True
        for k in (0, a.frames - 1):
            ctx.d2h(buf, ptrs[k])
            ok = ok and bool(np.array_equal(buf, images[k % 2]))
        out[mode] = {"feed": dec.feed_description(), "first_call_s": round(first, 3), "best_call_s": round(best, 4),
                     "gpix_per_s": round(a.frames * w * h / best / 1e9, 2), "file_gb_per_s": round(out["file_bytes"] / best / 1e9, 2),
                     "frames_ok": ok}
        dec.close()
    os.remove(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
