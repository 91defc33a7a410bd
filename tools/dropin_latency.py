#!/usr/bin/env python
"""Latency of the reference-shaped single-frame call (motioncam::raw::Decode: host bytes in, host uint16 out, synchronous)
through this repo's drop-in library, next to the compiled reference on one host core.  BASELINE config 1."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import oracle_lib as ol  # noqa: E402
from motioncam_decoder_b200 import hostapi, testvec as tv  # noqa: E402

w, h = 4080, 3072
img = tv.gen_photon(w, h, 1023, seed=1234)
s = tv.encode_current(img)
out = np.empty(w * h, dtype=np.uint16)


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        n = fn()
    return (time.perf_counter() - t0) / reps, n


ours = hostapi.library()
t, n = timed(lambda: ours.mcb200_decode(out.ctypes.data, w, h, s.ctypes.data, s.size), 20)
assert n == w * h and np.array_equal(out.reshape(h, w), img)
print(f"drop-in raw::Decode (H2D + kernels + D2H inside the call): {t * 1e3:.2f} ms per 4080x3072 frame = {w * h / t / 1e6:.0f} Mpix/s")
if ol.have_ref():
    ref = hostapi.library(ol.REF_SO, "mcref_")
    t, n = timed(lambda: ref.mcref_decode(out.ctypes.data, w, h, s.ctypes.data, s.size), 5)
    print(f"reference raw::Decode (g++ -O3, one host core):           {t * 1e3:.2f} ms per frame = {w * h / t / 1e6:.0f} Mpix/s")
