#!/usr/bin/env python
"""Host-side cost of mcraw_decode_batch: wall time of the enqueue calls themselves (GPU work is asynchronous)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from motioncam_decoder_b200 import capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
desc, w, h, ct, frames, streams = bench.make_streams(wl)
ctx = capi.Context(0)
items = []
for i in range(frames):
    s = streams[i % len(streams)]
    sp = ctx.device_alloc(len(s) + 256)
    dp = ctx.device_alloc(w * h * 2 + 256)
    ctx.h2d(sp, s)
    items.append((sp, len(s), w, h, ct, dp, w * h))
descs, n = capi.Context.make_descs(items)
for _ in range(40):
    ctx.decode_batch(descs, n)
ctx.batch_wait(n)
for reps in (4, 12, 12, 100):
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.decode_batch(descs, n)
    t1 = time.perf_counter()
    ctx.batch_wait(n)
    t2 = time.perf_counter()
    print(f"{wl}: {reps} calls: enqueue {1e6 * (t1 - t0) / reps:.1f} us/call, until done {1e6 * (t2 - t0) / reps:.1f} us/call")
