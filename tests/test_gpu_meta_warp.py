"""GPU parity of k_meta_warp (the index kernel of a batch: one warp per (frame, metadata stream), 2 KiB windows at fixed
addresses, double-buffered bulk copies).  The default context picks it for batches of more than half an SM count of frames;
MCRAW_META_WARP=2 forces it for every launch so that the small parity vectors reach it too, MCRAW_META_WARP=0 is the
CTA-per-stream kernel it replaced (k_meta).  All three against the oracle.
Reference: RawData.cpp:463-498 (DecodeMetadata), :528-612."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import vectors

pytestmark = pytest.mark.gpu


def _ctx_with(value):
    from motioncam_decoder_b200 import capi
    old = os.environ.get("MCRAW_META_WARP")
    os.environ["MCRAW_META_WARP"] = value
    try:
        return capi.Context(0)
    finally:
        if old is None:
            del os.environ["MCRAW_META_WARP"]
        else:
            os.environ["MCRAW_META_WARP"] = old


@pytest.fixture(scope="module")
def ctxs():
    forced, plain = _ctx_with("2"), _ctx_with("0")
    yield forced, plain
    forced.close()
    plain.close()


def _decode(ctx, frames, passes=2):
    from motioncam_decoder_b200 import capi
    batch = capi.DeviceBatch(ctx, [(s, w, h, capi.COMPRESSION_CURRENT) for (_, s, w, h, _) in frames])
    res = []
    for _ in range(passes):                                 # second pass: the plan of the slot is reused
        batch.fill_outputs(0x5A5A)
        written, status = batch.decode()
        res = [(int(written[i]), int(status[i]), batch.fetch(i)) for i in range(len(frames))]
    batch.free()
    return res


def test_vectors_through_warp_kernel(ctxs):
    """Every current-format parity vector (all header values, odd offsets, wide encodedWidth, random streams), one launch."""
    forced, _ = ctxs
    vecs = vectors.current_vectors(small=True)
    for (name, s, w, h, img), (n_got, st, pix) in zip(vecs, _decode(forced, vecs)):
        n, want = ol.oracle_decode(s, w, h)
        assert st == 0 and n_got == n == w * h, (name, st, n_got, n)
        assert np.array_equal(pix, want), f"{name}: k_meta_warp path differs from the oracle"
        if img is not None:
            assert np.array_equal(pix, img), name


def _long_streams():
    from motioncam_decoder_b200 import testvec as tv
    out = []
    img = tv.gen_photon(1920, 1080, 4095, seed=61)                              # ~17 KB per stream: 9 windows
    s = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=61)
    out.append(("photon_1080p", s, 1920, 1080, img))
    out.append(("photon_1080p_pad_1_3", tv.pad_meta_current(s, 1, 3), 1920, 1080, img))   # odd offsets: plain staging all the way
    out.append(("photon_1080p_pad_8_5", tv.pad_meta_current(s, 8, 5), 1920, 1080, img))
    img = tv.gen_flatnoise(2048, 768, cell=256, seed=62)                        # 2-byte next to 82-byte meta blocks
    out.append(("flatnoise_2048x768", tv.encode_current(img), 2048, 768, img))
    img = tv.gen_uniform(1024, 1024, 0, 65535, seed=63)                         # 130-byte meta blocks: every block straddles into the margin somewhere
    out.append(("uniform16_1024", tv.encode_current(img, ref_wrap=True, seed=63), 1024, 1024, img))
    img = tv.gen_uniform(4096, 1024, 7, 7, seed=64)                             # constant image: 2-byte meta blocks, 1024 per window
    out.append(("constant_4096x1024", tv.encode_current(img), 4096, 1024, img))
    rng = np.random.default_rng(65)
    ew, eh = 2048, 512
    nb = ew * eh // 64
    bits = rng.integers(0, 17, nb).astype(np.uint16)
    refs = rng.integers(0, 65536, nb).astype(np.uint16)
    out.append(("random_stream_2048x512", tv.assemble_current(ew, eh, bits, refs, seed=66), 2000, 512, None))
    return out


def test_long_streams_match_oracle_and_cta_kernel(ctxs):
    forced, plain = ctxs
    frames = _long_streams()
    got_w = _decode(forced, frames)
    got_p = _decode(plain, frames)
    for (name, s, w, h, img), (nw, sw, pw), (np_, sp, pp) in zip(frames, got_w, got_p):
        n, want = ol.oracle_decode(s, w, h)
        assert sw == 0 and sp == 0, (name, sw, sp)
        assert nw == np_ == n, (name, nw, np_, n)
        assert np.array_equal(pw, want), f"{name}: k_meta_warp path differs from the oracle"
        assert np.array_equal(pp, want), name
        if img is not None:
            assert np.array_equal(pw, img), name


def test_default_context_batch_uses_warp_kernel_and_matches():
    """A batch above the few-frames threshold through the DEFAULT context (what bench.py's C2 / C3 legs run)."""
    from motioncam_decoder_b200 import capi, testvec as tv
    ctx = capi.Context(0)
    frames = []
    for k in range(96):
        w, h = (640, 64) if k % 3 else (1000, 32)
        img = tv.gen_photon(w, h, 4095 if k % 2 else 1023, seed=700 + k)
        s = tv.encode_current(img, policy=tv.POLICY_ALIASES if k % 4 else tv.POLICY_MINIMAL, seed=k)
        if k % 5 == 0:
            s = tv.pad_meta_current(s, k % 7, (k // 5) % 4)
        frames.append((f"f{k}", s, w, h, img))
    for (name, s, w, h, img), (n_got, st, pix) in zip(frames, _decode(ctx, frames)):
        assert st == 0 and n_got == w * h, (name, st, n_got)
        assert np.array_equal(pix, img), name
    ctx.close()


def test_warp_kernel_rejects_like_cta_kernel(ctxs):
    """Truncated buffers, a metadata count that is too small, a bits value > 16, a stream that starts in the last bytes of
    the buffer: same verdicts (and status words) from both index kernels, and a good frame in the same launch still decodes."""
    from motioncam_decoder_b200 import capi, testvec as tv
    forced, plain = ctxs
    img = tv.gen_photon(1920, 1080, 4095, seed=71)
    good = tv.encode_current(img, seed=71)
    ew, eh, boff, roff = (int(v) for v in np.frombuffer(good[:16].tobytes(), dtype="<u4"))
    cases = [("good", good)]
    cases.append(("cut_in_refs", good[: roff + (len(good) - roff) // 2].copy()))
    cases.append(("cut_in_bits", good[: boff + (roff - boff) // 2].copy()))
    cases.append(("cut_last_byte", good[:-1].copy()))
    cases.append(("cut_inside_refs_count", good[: roff + 2].copy()))
    cases.append(("cut_before_bits", good[: boff].copy()))
    s = good.copy(); s[boff:boff + 4] = np.frombuffer(np.uint32(7).tobytes(), dtype=np.uint8); cases.append(("small_count", s))
    s = good.copy(); s[boff + 4] = 0xF0 | (s[boff + 4] & 0x0F); s[boff + 5] = 0xFF; cases.append(("bits_ref_gt_16", s))
    s = good.copy(); s[0:4] = np.frombuffer(np.uint32(ew + 1).tobytes(), dtype=np.uint8); cases.append(("encoded_width_odd", s))
    frames = [(s, 1920, 1080, capi.COMPRESSION_CURRENT) for _, s in cases]
    out = {}
    for label, ctx in (("warp", forced), ("cta", plain)):
        batch = capi.DeviceBatch(ctx, frames)
        batch.fill_outputs(0)
        written, status = batch.decode()
        out[label] = [(int(written[i]), int(status[i])) for i in range(len(frames))]
        assert np.array_equal(batch.fetch(0), img), label
        batch.free()
    for (name, s), a, b in zip(cases, out["warp"], out["cta"]):
        n, _ = ol.oracle_decode(s, 1920, 1080)
        assert a == b, (name, a, b)
        assert a[0] == n, (name, a, n)
        if name == "good":
            assert a[0] == 1920 * 1080 and a[1] == 0
        else:
            assert a[0] == 0 and a[1] != 0, (name, a)
