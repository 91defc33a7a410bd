#!/usr/bin/env python
"""c2_steps.py -- torch-free A/B timer for the device-resident C2 step (240 frames 1920x1080, 16 distinct, type 7):
wall clock over many back-to-back mcraw_decode_batch calls on one context (the GPU runs them back to back, the host
stays ahead), outputs checked against the source images afterwards.  Used to compare builds / environment switches
(e.g. MCRAW_CROSS_BATCH, MCRAW_NO_OVERLAP) in seconds; bench.py stays the number of record.

    python tools/c2_steps.py [--steps 400] [--frames 240]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from motioncam_decoder_b200 import capi, testvec as tv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--label", default="")
    ap.add_argument("--levels", type=int, default=0, help="fused black/white level epilogue: 1 = black-subtracted u16, 2 = normalised half")
    a = ap.parse_args()
    w, h = 1920, 1080
    images = [tv.gen_photon(w, h, 4095, seed=s + 1) for s in range(16)]
    streams = [tv.encode_current(img) for img in images]
    ctx = capi.Context(0)
    items = []
    for i in range(a.frames):
        s = streams[i % 16]
        sp = ctx.device_alloc(len(s) + 256)
        dp = ctx.device_alloc(w * h * 2 + 256)
        ctx.h2d(sp, s)
        items.append((sp, len(s), w, h, capi.COMPRESSION_CURRENT, dp, w * h))
    descs, n = capi.Context.make_descs(items)
    levels = [([64, 64, 64, 64], 4095.0, a.levels)] * a.frames if a.levels else None
    if levels:
        larr = (capi.Levels * n)()
        for i, (black, white, mode) in enumerate(levels):
            larr[i].black[:] = [float(v) for v in black]
            larr[i].white, larr[i].mode = float(white), int(mode)

        def decode():
            ctx._check(ctx._c.mcraw_decode_batch_levels(ctx._h, descs, larr, n, None), "mcraw_decode_batch_levels")
    else:
        def decode():
            ctx.decode_batch(descs, n)
    for _ in range(24):                       # every slot of the context has seen this plan
        decode()
    ctx.batch_wait(n)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(a.steps):
            decode()
        written, status = ctx.batch_wait(n)
        best = min(best, (time.perf_counter() - t0) / a.steps)
    ok = all(x == w * h for x in written) and not any(status)
    out = np.empty((h, w), np.uint16)
    for i in sorted(set(range(min(16, a.frames))) | {a.frames - 1}):
        ctx.d2h(out, items[i][5])
        want = capi.apply_levels(images[i % 16], [64, 64, 64, 64], 4095.0, a.levels) if a.levels else images[i % 16]
        ok = ok and bool(np.array_equal(out, want))
    print(json.dumps({"label": a.label, "env": {k: v for k, v in os.environ.items() if k.startswith("MCRAW_")},
                      "ms_per_step": round(best * 1e3, 4), "tpix_per_s": round(a.frames * w * h / best / 1e12, 3), "outputs_ok": ok}))


if __name__ == "__main__":
    main()
