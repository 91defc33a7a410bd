"""GPU: the optional black / white level epilogue fused into the pixel kernels (mcraw_decode_batch_levels, SURVEY.md 8f-4;
the reference's consumer side starts from the container's blackLevel[4] / whiteLevel, example.cpp:66-67,89-91).

The checker is the oracle's decode followed by capi.apply_levels, a numpy restatement of the arithmetic stated in
include/mcraw_b200.h.  MCRAW_OUT_BLACK_SUB is integer work: bit-exact.  MCRAW_OUT_NORM_F16 is floating point: the
kernel and the restatement both compute (float32(v) - black) * (1 / (white - black)) in float32 (one subtraction, one
multiplication: nothing to contract), clamp, and round once to IEEE half -- tolerance 0 ulp, i.e. the bit patterns match."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

LEVELS = [
    ([64, 64, 64, 64], 1023.0),
    ([64.0, 65.5, 63.25, 70.0], 4095.0),
    ([0, 0, 0, 0], 65535.0),
    ([100, 200, 300, 400], 150.0),          # white below some black levels: range 0 / scale 0 there
    ([256.5, 255.5, 1e9, -5.0], 1000.0),    # halves round to even; out-of-range levels clamp to the u16 range
]


def _frames():
    from motioncam_decoder_b200 import capi, testvec as tv
    out = []
    for k, (w, h, mx) in enumerate([(328, 12, 1023), (1928, 8, 4095), (64, 4, 65535), (100, 6, 16383), (1024, 36, 4095)]):
        img = tv.gen_photon(w, h, mx, seed=90 + k) if mx < 65535 else tv.gen_uniform(w, h, 0, 65535, seed=90 + k)
        h4 = h - h % 4
        out.append((tv.encode_current(img[:h4], policy=tv.POLICY_ALIASES, seed=k), w, h4, capi.COMPRESSION_CURRENT, img[:h4]))
        out.append((tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=k), w, h, capi.COMPRESSION_LEGACY, img))
    return out


@pytest.mark.parametrize("mode", [1, 2])
def test_levels_epilogue(mode):
    from motioncam_decoder_b200 import capi
    ctx = capi.Context(0)
    frames = _frames()
    for black, white in LEVELS:
        batch = capi.DeviceBatch(ctx, [(s, w, h, ct) for (s, w, h, ct, _) in frames])
        batch.fill_outputs(0xA5A5)
        # one frame of the batch stays raw: modes may be mixed inside a batch
        levels = [(black, white, mode if i != 1 else capi.OUT_RAW) for i in range(len(frames))]
        ctx.decode_batch_levels(batch.descs, levels, batch.n)
        written, status = ctx.batch_wait(batch.n)
        for i, (s, w, h, ct, img) in enumerate(frames):
            n, dec = (ol.oracle_decode if ct == capi.COMPRESSION_CURRENT else ol.oracle_decode_legacy)(s, w, h)
            assert n == w * h and np.array_equal(dec, img)
            assert status[i] == 0 and written[i] == w * h
            want = capi.apply_levels(dec, black, white, levels[i][2])
            got = batch.fetch(i)
            assert np.array_equal(got, want), (i, black, white, mode, np.argwhere(got != want)[:4])
        # the same descriptors without levels: raw again (the plan cache keys on the levels too)
        written, status = batch.decode()
        for i, (s, w, h, ct, img) in enumerate(frames):
            assert np.array_equal(batch.fetch(i), img), i
        batch.free()
    ctx.close()


def test_levels_rejects_bad_arguments():
    from motioncam_decoder_b200 import capi, testvec as tv
    ctx = capi.Context(0)
    img = tv.gen_photon(64, 4, 1023, seed=1)
    batch = capi.DeviceBatch(ctx, [(tv.encode_current(img), 64, 4, capi.COMPRESSION_CURRENT)])
    with pytest.raises(capi.McrawError, match="unknown output mode"):
        ctx.decode_batch_levels(batch.descs, [([0, 0, 0, 0], 1023.0, 7)], 1)
    with pytest.raises(capi.McrawError, match="not a number"):
        ctx.decode_batch_levels(batch.descs, [([float("nan"), 0, 0, 0], 1023.0, 1)], 1)
    batch.free()
    ctx.close()
