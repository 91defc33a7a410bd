#!/usr/bin/env python
"""legacy_runup_stats.py -- how long a run-up k_legacy_warp needs: for sampled tile starts of a C4
frame, follow the chains started at all 17 even offsets and report the share of tiles whose chains have NOT become one
after R segments of 124 bytes.  CPU only (numpy + the test-vector encoder).

    python tools/legacy_runup_stats.py [--width 4000 --height 3000 --maxval 1023]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motioncam_decoder_b200 import testvec as tv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=4000)
    ap.add_argument("--height", type=int, default=3000)
    ap.add_argument("--maxval", type=int, default=1023)
    ap.add_argument("--stride", type=int, default=14880, help="tile stride in bytes")
    a = ap.parse_args()
    img = tv.gen_photon(a.width, a.height, a.maxval, seed=1)
    s = np.frombuffer(tv.encode_legacy(img), dtype=np.uint8)

    def step(p):
        b = int(s[p]) >> 4
        return p + 2 + (32 if b > 10 else 2 * b)

    for runup in (2, 3, 4, 6, 8):
        bad = n = 0
        for t in range(1, len(s) // a.stride - 1, 3):
            start, end = a.stride * t, a.stride * t + 124 * runup
            exits = set()
            for e in range(17):
                p = start + 2 * e
                while p < end:
                    p = step(p)
                exits.add(p)
            n += 1
            bad += len(exits) > 1
        print(f"run-up {runup} segments ({124 * runup} bytes): {bad} of {n} sampled tiles not merged ({100.0 * bad / n:.1f} %)")


if __name__ == "__main__":
    main()
