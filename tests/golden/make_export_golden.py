#!/usr/bin/env python
"""Generates tests/golden/export_golden.json: small decoded frames + metadata together with the DNG bytes, and audio
chunks together with the WAV bytes, that the UNMODIFIED reference program writes for them (/root/reference/example.cpp
:27-53, :55-139, compiled into oracle/_ref/libmcraw_ref.so by oracle/Makefile and called through oracle/ref_shim.cpp).

The reference has no fixtures for its example program; these pin the expected files by executing it here, once, and
travel with the repo (tests/test_export_cpu.py::test_golden_*).  Metadata values are chosen to hit every conversion
rule: float and > 16-bit black levels, white levels beyond `short`, matrix entries that are integers, zero, negative,
tiny (denominator beyond 32 bits / capped at 2^127) and huge (numerator beyond 32 bits).

    python tests/golden/make_export_golden.py          # needs oracle/_ref/libmcraw_ref.so
"""
import base64
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from motioncam_decoder_b200 import hostapi  # noqa: E402

IDENT = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


def b64(b):
    return base64.b64encode(bytes(b)).decode()


def build():
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle"
    ref = hostapi.library(ol.REF_SO, "mcref_")
    rng = np.random.default_rng(2024)
    tmp = tempfile.mkdtemp()
    dng, wav = [], []
    matrices = [
        IDENT,
        [0.9, -0.3, 0.01, -0.4, 1.2, 0.2, 0.003, 0.1, 0.7],
        [1e-12, -2.5e-9, 3.0, 1e6, 1e12, -1e10, 0.5, 0.25, 7.0e-4],
        [float(np.float32(3.0) * np.float32(2.0) ** -120), float(np.float32(-5.0) * np.float32(2.0) ** -130), 2.0 ** -126,
         2.0 ** -100, 1.0 / 3.0, -1.0 / 3.0, 16777216.0, 16777217.0, 4294967296.0],
    ]
    cases = [
        dict(w=16, h=8, black=[64, 64, 64, 64], white=1023.0, cfa="rggb", m=(0, 0, 0, 0), neutral=[1.0, 1.0, 1.0]),
        dict(w=34, h=6, black=[64.5, 63.25, 10, 11.9], white=4095.0, cfa="bggr", m=(1, 0, 1, 0), neutral=[0.51, 1.0, 0.63]),
        dict(w=2, h=2, black=[65535, 0, 70000, 1], white=65535.0, cfa="grbg", m=(1, 1, 1, 1), neutral=[0.5, 0.25, 2.0]),
        dict(w=64, h=4, black=[0, 1, 2, 3], white=16383.7, cfa="gbrg", m=(2, 1, 0, 2), neutral=[1e-12, 1e12, 3.0]),
        dict(w=10, h=10, black=[16, 16, 16, 16], white=-3.0, cfa="rggb", m=(3, 2, 3, 1), neutral=[2.0 ** -130, 1.0, 1.0]),
        dict(w=8, h=2, black=[1, 2, 3, 4], white=1e12, cfa="bggr", m=(0, 3, 2, 3), neutral=[-1.5, 0.0, 4294967296.0]),
    ]
    for i, c in enumerate(cases):
        px = rng.integers(0, 65536, (c["h"], c["w"]), dtype=np.uint16)
        cm = {"blackLevel": c["black"], "whiteLevel": c["white"], "sensorArrangment": c["cfa"],
              "colorMatrix1": matrices[c["m"][0]], "colorMatrix2": matrices[c["m"][1]],
              "forwardMatrix1": matrices[c["m"][2]], "forwardMatrix2": matrices[c["m"][3]]}
        fm = {"width": c["w"], "height": c["h"], "asShotNeutral": c["neutral"]}
        path = os.path.join(tmp, "g.dng")
        hostapi.write_dng(path, px, fm, cm, lib=ref, prefix="mcref_")
        dng.append({"name": f"dng{i}", "pixels": b64(px.tobytes()), "frame": fm, "container": cm, "file": b64(open(path, "rb").read())})
    for i, (channels, lens) in enumerate([(2, [8, 2, 32]), (1, [5, 1]), (2, []), (3, [6]), (1, [])]):
        chunks = [rng.integers(-32768, 32768, n, dtype=np.int16) for n in lens]
        path = os.path.join(tmp, "g.wav")
        hostapi.write_audio(path, 48000 if i % 2 == 0 else 44100, channels, chunks, lib=ref, prefix="mcref_")
        wav.append({"name": f"wav{i}", "rate": 48000 if i % 2 == 0 else 44100, "channels": channels,
                    "chunks": [b64(x.tobytes()) for x in chunks], "file": b64(open(path, "rb").read())})
    with open(os.path.join(HERE, "export_golden.json"), "w") as f:
        json.dump({"dng": dng, "wav": wav}, f, indent=0)
    print(f"wrote {len(dng)} DNG and {len(wav)} WAV cases")


if __name__ == "__main__":
    build()
