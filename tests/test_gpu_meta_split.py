"""GPU parity of k_meta_split (the metadata chain of a stream resolved by several CTAs at once, picked for a handful of
big current-format frames): the same frames through a context with the split kernel (default) and one without
(MCRAW_META_SPLIT=0 -> k_meta), both against the oracle.  Reference: RawData.cpp:463-498 (DecodeMetadata), :528-612."""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    from motioncam_decoder_b200 import capi
    split = capi.Context(0)
    os.environ["MCRAW_META_SPLIT"] = "0"
    try:
        plain = capi.Context(0)
    finally:
        del os.environ["MCRAW_META_SPLIT"]
    yield split, plain
    split.close()
    plain.close()


def _big_frames():
    from motioncam_decoder_b200 import testvec as tv
    out = []
    img = tv.gen_photon(1920, 1080, 4095, seed=41)                              # refs stream ~66 KB: 5 windows
    s = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=41)
    out.append(("photon_1080p", s, 1920, 1080, img))
    out.append(("photon_1080p_pad_1_3", tv.pad_meta_current(s, 1, 3), 1920, 1080, img))   # both streams at odd offsets
    out.append(("photon_1080p_pad_8_5", tv.pad_meta_current(s, 8, 5), 1920, 1080, img))
    img = tv.gen_flatnoise(2048, 768, cell=256, seed=42)                        # 0-bit next to 10-bit blocks: 2-byte and 82-byte meta blocks
    out.append(("flatnoise_2048x768", tv.encode_current(img), 2048, 768, img))
    img = tv.gen_uniform(1024, 1024, 0, 65535, seed=43)                         # 16-bit metadata blocks (130 bytes: the longest)
    out.append(("uniform16_1024", tv.encode_current(img, ref_wrap=True, seed=43), 1024, 1024, img))
    img = tv.gen_uniform(4096, 512, 7, 7, seed=44)                              # constant image: every meta block is 2 bytes long
    out.append(("constant_4096x512", tv.encode_current(img), 4096, 512, img))
    rng = np.random.default_rng(45)
    ew, eh = 2048, 512
    nb = ew * eh // 64
    bits = rng.integers(0, 17, nb).astype(np.uint16)
    refs = rng.integers(0, 65536, nb).astype(np.uint16)
    out.append(("random_stream_2048x512", tv.assemble_current(ew, eh, bits, refs, seed=46), 2000, 512, None))
    return out


def _run(ctx, frames):
    from motioncam_decoder_b200 import capi
    batch = capi.DeviceBatch(ctx, [(s, w, h, capi.COMPRESSION_CURRENT) for (_, s, w, h, _) in frames])
    res = []
    for _ in range(2):                                     # second pass: the plan and its epoch-tagged flags are reused
        batch.fill_outputs(0x5A5A)
        written, status = batch.decode()
        res = [(int(written[i]), int(status[i]), batch.fetch(i)) for i in range(len(frames))]
    batch.free()
    return res


def test_split_matches_oracle_and_plain_kernel(ctxs):
    split, plain = ctxs
    frames = _big_frames()
    for group in (frames[:1], frames[1:4], frames[4:]):    # one frame, a few, a few more: different window counts per launch
        got_s = _run(split, group)
        got_p = _run(plain, group)
        for (name, s, w, h, img), (ws, ss, ps), (wp, sp, pp) in zip(group, got_s, got_p):
            n, want = ol.oracle_decode(s, w, h)
            assert ss == 0 and sp == 0, (name, ss, sp)
            assert ws == wp == n, (name, ws, wp, n)
            assert np.array_equal(ps, want), f"{name}: split index kernel differs from the oracle"
            assert np.array_equal(pp, want), name
            if img is not None:
                assert np.array_equal(ps, img), name


def test_split_rejects_like_plain_kernel(ctxs):
    """Truncated buffers, a metadata count that is too small, a bits value > 16: same verdicts from both index kernels,
    and a good frame in the same launch still decodes."""
    from motioncam_decoder_b200 import capi, testvec as tv
    split, plain = ctxs
    img = tv.gen_photon(1920, 1080, 4095, seed=51)
    good = tv.encode_current(img, seed=51)
    ew, eh, boff, roff = (int(v) for v in np.frombuffer(good[:16].tobytes(), dtype="<u4"))
    cases = [("good", good)]
    cases.append(("cut_in_refs", good[: roff + (len(good) - roff) // 2].copy()))
    cases.append(("cut_in_bits", good[: boff + (roff - boff) // 2].copy()))
    cases.append(("cut_last_byte", good[:-1].copy()))
    s = good.copy(); s[boff:boff + 4] = np.frombuffer(np.uint32(7).tobytes(), dtype=np.uint8); cases.append(("small_count", s))
    s = good.copy(); s[boff + 4] = 0xF0 | (s[boff + 4] & 0x0F); s[boff + 5] = 0xFF; cases.append(("bits_ref_gt_16", s))
    frames = [(s, 1920, 1080, capi.COMPRESSION_CURRENT) for _, s in cases]
    out = {}
    for label, ctx in (("split", split), ("plain", plain)):
        batch = capi.DeviceBatch(ctx, frames)
        batch.fill_outputs(0)
        written, status = batch.decode()
        out[label] = [(int(written[i]), int(status[i])) for i in range(len(frames))]
        assert np.array_equal(batch.fetch(0), img), label
        batch.free()
    for (name, s), a, b in zip(cases, out["split"], out["plain"]):
        assert (a[0] == 0) == (b[0] == 0), (name, a, b)
        assert (a[1] == 0) == (b[1] == 0), (name, a, b)
        if name == "good":
            assert a[0] == 1920 * 1080 and a[1] == 0
        else:
            assert a[0] == 0 and a[1] != 0, (name, a)
