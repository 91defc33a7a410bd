#!/usr/bin/env python
"""Generates tests/golden/mcraw_golden.npz: small compressed frames of both formats together with the output of
the UNMODIFIED reference decoder (oracle/_ref/libmcraw_ref.so, built by oracle/Makefile from /root/reference).

The reference repository has no test vectors of its own, so these fixtures pin the
expected bytes by executing the reference here, once; the file travels with the repo and is what the oracle and
the CUDA path are checked against where /root/reference does not exist (the GPU box).

    python tests/golden/make_golden.py          # needs oracle/_ref/libmcraw_ref.so
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from motioncam_decoder_b200 import testvec as tv  # noqa: E402


def build():
    assert ol.have_ref(), "build oracle/_ref first: make -C oracle"
    out = {}
    names = []

    def add(name, stream, w, h, legacy):
        n, img = (ol.ref_decode_legacy if legacy else ol.ref_decode)(stream, w, h)
        assert n == w * h, name
        out[name + ".stream"] = np.asarray(stream, dtype=np.uint8)
        out[name + ".expect"] = img.copy()
        out[name + ".meta"] = np.array([w, h, 6 if legacy else 7], dtype=np.int32)
        names.append(name)

    # current format: every header value 0..16 (incl. aliases 7, 9, 11..15) forced on one 64x4 tile each
    for hb in range(17):
        w_needed = {7: 7, 9: 9}.get(hb, hb if hb <= 10 else 16)
        img = tv.gen_forced_widths(64, 4, [w_needed], seed=700 + hb)
        add(f"cur_hdr{hb:02d}", tv.encode_current(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=hb), 64, 4, False)
    img = tv.gen_photon(200, 8, 1023, seed=71)
    add("cur_photon_crop200", tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=1), 200, 8, False)
    img = tv.gen_uniform(128, 4, 0, 65535, seed=72)
    add("cur_wrap16", tv.encode_current(img, ref_wrap=True, seed=2), 128, 4, False)
    rng = np.random.default_rng(73)
    bits = rng.integers(0, 17, 12).astype(np.uint16)
    refs = rng.integers(0, 65536, 12).astype(np.uint16)
    add("cur_random_stream", tv.assemble_current(192, 4, bits, refs, seed=3), 150, 4, False)
    # metadata streams at odd offsets; a frame encoded wider than it is decoded (encodedWidth >= width + 64)
    img = tv.gen_photon(320, 8, 4095, seed=74)
    enc = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=5)
    add("cur_meta_pad_1_0", tv.pad_meta_current(enc, 1, 0), 320, 8, False)
    add("cur_meta_pad_3_2", tv.pad_meta_current(enc, 3, 2), 320, 8, False)
    wide = tv.gen_photon(256, 8, 1023, seed=75)
    add("cur_wide_256_as_100", tv.encode_current(wide, seed=6), 100, 8, False)
    # legacy format: every header nibble 0..15 on one row of 64 px
    for nib in range(16):
        w_needed = nib if nib <= 10 else 16
        img = tv.gen_forced_widths(64, 2, [w_needed], seed=800 + nib)
        add(f"leg_nib{nib:02d}", tv.encode_legacy(img, policy=tv.POLICY_FORCE, policy_arg=nib, seed=nib), 64, 2, True)
    img = tv.gen_photon(100, 5, 4095, seed=81)
    add("leg_photon_crop100", tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=4), 100, 5, True)
    img = tv.gen_flatnoise(256, 6, cell=64, seed=82)
    add("leg_flatnoise_trailer", tv.encode_legacy(img, trailer_records=2), 256, 6, True)
    out["names"] = np.array(names)
    return out


if __name__ == "__main__":
    data = build()
    path = os.path.join(HERE, "mcraw_golden.npz")
    np.savez_compressed(path, **data)
    print(f"wrote {path}: {len(data['names'])} vectors, {os.path.getsize(path)} bytes")
