"""GPU parity: the CUDA path (through the C-ABI, ctypes) against the oracle on the shared seeded vectors,
legacy frame format (compressionType 6, RawData_Legacy.cpp:445-495).  Bit-exact or fail."""
import numpy as np
import pytest

import oracle_lib as ol
import vectors
from test_gpu_current import _check_batch, ctx  # noqa: F401  (shared fixture / helper)

pytestmark = pytest.mark.gpu


def test_legacy_vectors_batched(ctx):
    from motioncam_decoder_b200 import capi
    _check_batch(ctx, vectors.legacy_vectors(small=True), capi.COMPRESSION_LEGACY, ol.oracle_decode_legacy)


def test_legacy_vectors_one_by_one_host_call(ctx):
    from motioncam_decoder_b200 import capi
    for name, s, w, h, img in vectors.legacy_vectors(small=True):
        n, got = ctx.decode_host(s, w, h, capi.COMPRESSION_LEGACY)
        n_or, want = ol.oracle_decode_legacy(s, w, h)
        assert n == n_or == w * h, name
        assert np.array_equal(got, want), name


def test_legacy_full_size(ctx):
    """C4 geometry: 4000x3000, 750 000 chained blocks, ~9 MB -> 282 tiles of 32 KiB."""
    from motioncam_decoder_b200 import capi
    vecs = [v for v in vectors.legacy_vectors(small=False) if v[0] == "legacy_c4"]
    _check_batch(ctx, vecs, capi.COMPRESSION_LEGACY, ol.oracle_decode_legacy)


def test_legacy_constant_width_regions(ctx):
    """Chains that never merge: every block the same width (uniform noise), so each of the 17 candidate entry
    offsets of a segment leads to a different exit -- the full transfer map is needed, not a guess."""
    from motioncam_decoder_b200 import capi, testvec as tv
    vecs = []
    for k, hb in enumerate([1, 3, 7, 10, 15]):
        img = tv.gen_uniform(2048, 24, 0, (1 << min(hb, 12)) - 1, seed=300 + k)
        vecs.append((f"const_nib{hb}", tv.encode_legacy(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=k), 2048, 24, img))
    _check_batch(ctx, vecs, capi.COMPRESSION_LEGACY, ol.oracle_decode_legacy)


def test_mixed_batch_both_formats(ctx):
    """One batch may mix compression types and sizes (mcraw_frame_desc.compression_type per frame)."""
    from motioncam_decoder_b200 import capi
    cur = vectors.current_vectors(small=True)[:6]
    leg = vectors.legacy_vectors(small=True)[:6]
    frames, want = [], []
    for (a, b) in zip(cur, leg):
        frames.append((a[1], a[2], a[3], capi.COMPRESSION_CURRENT)); want.append(ol.oracle_decode(a[1], a[2], a[3]))
        frames.append((b[1], b[2], b[3], capi.COMPRESSION_LEGACY)); want.append(ol.oracle_decode_legacy(b[1], b[2], b[3]))
    batch = capi.DeviceBatch(ctx, frames)
    batch.fill_outputs(0xA5A5)
    written, status = batch.decode()
    for i, (n, img) in enumerate(want):
        assert status[i] == 0 and written[i] == n
        assert np.array_equal(batch.fetch(i), img)
    batch.free()


def test_legacy_rejects_truncated(ctx):
    """The reference leaves stale samples when the chain runs into the end of the buffer
    (RawData_Legacy.cpp:387,398); here the frame fails with 0, like the oracle."""
    from motioncam_decoder_b200 import capi, testvec as tv
    img = tv.gen_photon(640, 16, 1023, seed=5)
    good = tv.encode_legacy(img)
    cases = [good[: len(good) // 2].copy(), good[: len(good) - 1].copy(), good[:1].copy()]
    frames = [(s, 640, 16, capi.COMPRESSION_LEGACY) for s in cases] + [(good, 640, 16, capi.COMPRESSION_LEGACY)]
    batch = capi.DeviceBatch(ctx, frames)
    written, status = batch.decode()
    for i, s in enumerate(cases):
        assert written[i] == 0 and status[i] & capi.FRAME_TRUNCATED, (i, written[i], status[i])
        assert ol.oracle_decode_legacy(s, 640, 16)[0] == 0
    assert written[-1] == 640 * 16 and status[-1] == 0
    assert np.array_equal(batch.fetch(len(cases)), img)
    batch.free()


def test_legacy_full_size_properties(ctx):
    """C4 geometry, flat+noise variant (2-byte blocks next to constant 22-byte blocks: chains that merge at once next
    to chains that never do), several copies in one batch: every copy equals the source image."""
    from motioncam_decoder_b200 import capi, testvec as tv
    img = tv.gen_flatnoise(4000, 3000, 256, seed=9)
    s = tv.encode_legacy(img)
    frames = [(s, 4000, 3000, capi.COMPRESSION_LEGACY)] * 3
    batch = capi.DeviceBatch(ctx, frames)
    batch.fill_outputs(0x5A5A)
    written, status = batch.decode()
    assert not any(status) and all(w == 4000 * 3000 for w in written)
    want = tv.fnv1a64(img)
    assert [tv.fnv1a64(batch.fetch(i)) for i in range(3)] == [want] * 3
    batch.free()


def test_legacy_whole_frame_one_width(ctx):
    """Worst case of the index: every block of the frame has the same non-zero width, so chains entered at different
    offsets never meet -- no run-up merges them, every tile of k_legacy_warp publishes its map and takes the look-back
    path for its entry.  Small frames (a few tiles) and a batch of larger ones (tens of tiles per frame, several frames
    in flight, so the maps are composed across tiles that are worked at the same time)."""
    from motioncam_decoder_b200 import capi, testvec as tv
    for hb in (4, 10):
        img = tv.gen_uniform(1024, 96, 0, (1 << hb) - 1, seed=77 + hb)
        s = tv.encode_legacy(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=hb)
        n_or, want = ol.oracle_decode_legacy(s, 1024, 96)
        batch = capi.DeviceBatch(ctx, [(s, 1024, 96, capi.COMPRESSION_LEGACY)])
        written, status = batch.decode()
        assert status[0] == 0 and written[0] == n_or == 1024 * 96
        assert np.array_equal(batch.fetch(0), want) and np.array_equal(want, img)
        batch.free()
    frames, wants = [], []
    for k, hb in enumerate((3, 4, 7, 10, 4)):
        w, h = 2048, 512 + 64 * k
        img = tv.gen_uniform(w, h, 0, (1 << hb) - 1, seed=300 + k)
        s = tv.encode_legacy(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=k)
        n_or, want = ol.oracle_decode_legacy(s, w, h)
        assert n_or == w * h and np.array_equal(want, img)
        frames.append((s, w, h, capi.COMPRESSION_LEGACY))
        wants.append(want)
    batch = capi.DeviceBatch(ctx, frames)
    for _ in range(3):                                   # the plan (and its epoch-tagged status words) is reused
        batch.fill_outputs()
        written, status = batch.decode()
        for i, (s, w, h, _ct) in enumerate(frames):
            assert status[i] == 0 and written[i] == w * h, i
            assert np.array_equal(batch.fetch(i), wants[i]), i
    batch.free()


def test_legacy_batches_back_to_back():
    """Legacy batches enqueued back to back without waiting: once a slot holds a batch's plan, its kernel is launched as a
    programmatic dependent of the legacy kernel before (same outputs, or disjoint ones) -- two grids of look-back status
    words, tickets and counters side by side.  Every frame of both batches against its source image, round after round."""
    from motioncam_decoder_b200 import capi, testvec as tv
    ctx = capi.Context(0)
    frames = []
    for k, (w, h) in enumerate([(2048, 96), (1000, 64), (4000, 48), (640, 128), (2048, 96), (3008, 40)]):
        img = tv.gen_photon(w, h, 1023 if k % 2 else 4095, seed=500 + k)
        s = tv.encode_legacy(img, policy=tv.POLICY_ALIASES if k % 3 else tv.POLICY_MINIMAL, seed=k)
        n, want = ol.oracle_decode_legacy(s, w, h)
        assert n == w * h and np.array_equal(want, img)
        frames.append((s, w, h, capi.COMPRESSION_LEGACY))
    img7 = tv.gen_uniform(2048, 64, 0, 127, seed=77)                       # constant width: tiles that publish maps and look back
    frames.append((tv.encode_legacy(img7, policy=tv.POLICY_FORCE, policy_arg=7, seed=3), 2048, 64, capi.COMPRESSION_LEGACY))
    images = [ol.oracle_decode_legacy(s, w, h)[1] for (s, w, h, _) in frames]
    a = capi.DeviceBatch(ctx, frames * 3)
    b = capi.DeviceBatch(ctx, list(reversed(frames)) * 2)
    for rnd in range(3):
        a.fill_outputs(0xA5A5)
        b.fill_outputs(0x5A5A)
        for k in range(14):                                               # 6 slots: from the 7th enqueue on the plans are on the device
            ctx.decode_batch(a.descs, a.n) if (k % 2 == 0 or rnd == 0) else ctx.decode_batch(b.descs, b.n)
        ctx.decode_batch(b.descs, b.n)
        written, status = ctx.batch_wait(b.n)
        assert not any(status) and all(wr == f[1] * f[2] for wr, f in zip(written, b.frames)), rnd
        for i in range(a.n):
            assert np.array_equal(a.fetch(i), images[i % len(frames)]), (rnd, "a", i)
        for i in range(b.n):
            assert np.array_equal(b.fetch(i), images[len(frames) - 1 - i % len(frames)]), (rnd, "b", i)
    a.free()
    b.free()
    ctx.close()
