"""GPU parity: the CUDA path (through the C-ABI, ctypes) against the oracle on the shared seeded vectors,
current frame format (compressionType 7).  Bit-exact or fail."""
import numpy as np
import pytest

import oracle_lib as ol
import vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from motioncam_decoder_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _check_batch(ctx, vecs, ctype, oracle_fn):
    from motioncam_decoder_b200 import capi
    frames = [(s, w, h, ctype) for (_, s, w, h, _) in vecs]
    batch = capi.DeviceBatch(ctx, frames)
    batch.fill_outputs(0xA5A5)
    written, status = batch.decode()
    for i, (name, s, w, h, img) in enumerate(vecs):
        n, want = oracle_fn(s, w, h)
        got = batch.fetch(i)
        assert status[i] == 0, (name, status[i])
        assert written[i] == n, (name, written[i], n)
        assert np.array_equal(got, want), f"{name}: CUDA output differs from the oracle"
        if img is not None:
            assert np.array_equal(got, img), name
    batch.free()


def test_current_vectors_batched(ctx):
    from motioncam_decoder_b200 import capi
    _check_batch(ctx, vectors.current_vectors(small=True), capi.COMPRESSION_CURRENT, ol.oracle_decode)


def test_current_vectors_one_by_one_host_call(ctx):
    """mcraw_decode_host = the reference-shaped call (host in / host out)."""
    from motioncam_decoder_b200 import capi
    for name, s, w, h, img in vectors.current_vectors(small=True):
        n, got = ctx.decode_host(s, w, h, capi.COMPRESSION_CURRENT)
        n_or, want = ol.oracle_decode(s, w, h)
        assert n == n_or == w * h, name
        assert np.array_equal(got, want), name


def test_current_full_size(ctx):
    from motioncam_decoder_b200 import capi
    vecs = [v for v in vectors.current_vectors(small=False) if v[0] in ("photon_1080p", "photon_c1")]
    _check_batch(ctx, vecs, capi.COMPRESSION_CURRENT, ol.oracle_decode)


def test_current_rejects_bad_frames(ctx):
    """Same 0-means-failure convention as RawData.cpp:547-554, plus the hardened corners."""
    from motioncam_decoder_b200 import capi, testvec as tv
    img = tv.gen_photon(128, 8, 1023, seed=1)
    good = tv.encode_current(img)

    def u32(v):
        return np.frombuffer(np.uint32(v).tobytes(), dtype=np.uint8)

    cases = {}
    s = good.copy(); s[0:4] = u32(96); cases["ew_not_64"] = (s, 128, capi.FRAME_BAD_HEADER)
    s = good.copy(); s[8:12] = u32(len(s) + 1); cases["bits_off"] = (s, 128, capi.FRAME_BAD_HEADER)
    s = good.copy(); s[12:16] = u32(len(s) + 1); cases["refs_off"] = (s, 128, capi.FRAME_BAD_HEADER)
    cases["ew_lt_width"] = (good.copy(), 192, capi.FRAME_BAD_HEADER)
    cases["truncated"] = (good[: len(good) - 7].copy(), 128, capi.FRAME_TRUNCATED)
    cases["short"] = (good[:8].copy(), 128, capi.FRAME_BAD_HEADER)
    frames = [(s, w, 8, capi.COMPRESSION_CURRENT) for (s, w, _) in cases.values()]
    frames.append((good, 128, 8, capi.COMPRESSION_CURRENT))      # a good frame in the same batch still decodes
    frames.append((good, 128, 8, 5))                             # unknown compression type (Decoder.cpp:233)
    batch = capi.DeviceBatch(ctx, frames)
    written, status = batch.decode()
    for i, (name, (_, _, want_status)) in enumerate(cases.items()):
        assert written[i] == 0, name
        assert status[i] & want_status, (name, status[i])
        n_or, _ = ol.oracle_decode(frames[i][0], frames[i][1], 8)
        assert n_or == 0, name
    assert written[len(cases)] == 128 * 8 and status[len(cases)] == 0
    assert np.array_equal(batch.fetch(len(cases)), img)
    assert written[len(cases) + 1] == 0 and status[len(cases) + 1] == capi.FRAME_BAD_TYPE
    batch.free()


def test_dst_capacity_crops_rows(ctx):
    """encodedHeight rows are emitted (RawData.cpp:571,598-608) but never past dst capacity."""
    from motioncam_decoder_b200 import capi, testvec as tv
    img = tv.gen_photon(256, 16, 1023, seed=3)
    s = tv.encode_current(img)
    sp = ctx.device_alloc(len(s) + 16)
    dp = ctx.device_alloc(256 * 16 * 2)
    ctx.h2d(sp, s)
    ctx.h2d(dp, np.full(256 * 16, 0xA5A5, dtype=np.uint16))
    descs, n = capi.Context.make_descs([(sp, len(s), 256, 16, capi.COMPRESSION_CURRENT, dp, 256 * 10)])
    ctx.decode_batch(descs, n)
    written, status = ctx.batch_wait(n)
    out = np.empty(256 * 16, dtype=np.uint16)
    ctx.d2h(out, dp)
    out = out.reshape(16, 256)
    assert status[0] == 0 and written[0] == 256 * 10
    assert np.array_equal(out[:10], img[:10]) and np.all(out[10:] == 0xA5A5)
    ctx.device_free(sp); ctx.device_free(dp)


def test_full_size_properties(ctx):
    """BASELINE.json full sizes through size-independent properties: C3 geometry (4080x3072 flat+noise: 0-bit and 10-bit
    blocks) decodes to the image it was encoded from, the same frame decodes identically at every batch position, and a
    checksum of checksums over the batch equals frames x the single-frame checksum."""
    from motioncam_decoder_b200 import capi, testvec as tv
    img = tv.gen_flatnoise(4080, 3072, 256, seed=3)
    s = tv.encode_current(img)
    frames = [(s, 4080, 3072, capi.COMPRESSION_CURRENT)] * 5
    batch = capi.DeviceBatch(ctx, frames)
    batch.fill_outputs(0x5A5A)
    written, status = batch.decode()
    assert not any(status) and all(w == 4080 * 3072 for w in written)
    want = tv.fnv1a64(img)
    sums = [tv.fnv1a64(batch.fetch(i)) for i in range(len(frames))]
    assert sums == [want] * len(frames)
    # idempotence: decoding again into the same (now non-trivial) buffers changes nothing
    written, status = batch.decode()
    assert not any(status) and [tv.fnv1a64(batch.fetch(i)) for i in range(len(frames))] == sums
    batch.free()


def test_mixed_sizes_in_one_batch(ctx):
    """Frames of very different sizes share one persistent launch (work queue of k_units)."""
    from motioncam_decoder_b200 import capi, testvec as tv
    specs = [(64, 4, 1), (4080, 64, 2), (128, 8, 3), (1920, 1080, 4), (200, 12, 5), (4096, 8, 6), (1928, 16, 7)]
    frames, imgs = [], []
    for w, h, seed in specs:
        img = tv.gen_photon(w, h, 1023, seed=seed)
        imgs.append(img)
        frames.append((tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=seed), w, h, capi.COMPRESSION_CURRENT))
    batch = capi.DeviceBatch(ctx, frames)
    batch.fill_outputs(0xA5A5)
    for _ in range(3):                      # the same descriptors again: the slot's cached plan is reused
        written, status = batch.decode()
        assert not any(status)
        for i, img in enumerate(imgs):
            assert written[i] == img.size and np.array_equal(batch.fetch(i), img)
    batch.free()


def test_host_batch_pipeline(ctx):
    """mcraw_decode_batch_host: pinned host sources, staged H2D on side streams (several chunks), same results."""
    from motioncam_decoder_b200 import capi, testvec as tv
    imgs = [tv.gen_photon(1920, 1080, 4095, seed=50 + k) for k in range(4)]
    streams = [tv.encode_current(im) for im in imgs]
    n = 120                                  # > 192 MB of input: three staging chunks
    offs, total = [], 0
    for i in range(n):
        offs.append(total)
        total += (len(streams[i % 4]) + 255) & ~255
    ring_ptr, ring = ctx.pinned_array(total + 256)
    dsts, items = [], []
    for i in range(n):
        s = streams[i % 4]
        ring[offs[i]:offs[i] + len(s)] = s
        dp = ctx.device_alloc(1920 * 1080 * 2)
        dsts.append(dp)
        items.append((ring_ptr + offs[i], len(s), 1920, 1080, capi.COMPRESSION_CURRENT, dp, 1920 * 1080))
    descs, m = capi.Context.make_descs(items)
    ctx.decode_batch_host(descs, m)
    written, status = ctx.batch_wait(m)
    assert not any(status) and all(w == 1920 * 1080 for w in written)
    out = np.empty((1080, 1920), dtype=np.uint16)
    for i in (0, 1, 51, 52, 53, 103, 104, 119):   # around the chunk boundaries
        ctx.d2h(out, dsts[i])
        assert np.array_equal(out, imgs[i % 4]), i
    for dp in dsts:
        ctx.device_free(dp)
    ctx.pinned_free(ring_ptr)


def test_encoded_width_wider_than_width(ctx):
    """RawData.cpp:550-554 accepts any encodedWidth that is a multiple of 64 and >= width, and crops the rows to width
    (:598-608).  Device-resident sources: the caller names the header's encodedWidth in the descriptor
    (mcraw_frame_encoded_width); host sources: the library reads the header itself."""
    from motioncam_decoder_b200 import capi, testvec as tv
    wide = tv.gen_photon(320, 12, 1023, seed=77)
    s = tv.encode_current(wide, policy=tv.POLICY_ALIASES, seed=78)          # encodedWidth 320
    for w in (60, 128, 200, 256):                                            # planned default would be 64 / 128 / 256 / 256
        n_or, want = ol.oracle_decode(s, w, 12)
        assert n_or == w * 12 and np.array_equal(want, wide[:, :w])
        hint = capi.frame_encoded_width(s, w, 12)
        assert hint == 320
        batch = capi.DeviceBatch(ctx, [(s, w, 12, capi.COMPRESSION_CURRENT, hint), (s, w, 12, capi.COMPRESSION_CURRENT, 0)])
        batch.fill_outputs(0xA5A5)
        written, status = batch.decode()
        assert written[0] == w * 12 and status[0] == 0, (w, written, status)
        assert np.array_equal(batch.fetch(0), want), w
        # without the hint the work list was planned for width rounded up to 64: reported, never decoded wrongly
        assert written[1] == 0 and status[1] & capi.FRAME_GEOMETRY, (w, written, status)
        batch.free()
        n, got = ctx.decode_host(s, w, 12, capi.COMPRESSION_CURRENT)        # host source: header read by the library
        assert n == w * 12 and np.array_equal(got, want), w
    assert capi.frame_encoded_width(s, 320, 12) == 0 and capi.frame_encoded_width(s, 300, 12) == 0
