"""GPU: mcraw_decode_batch_host_out -- host sources in, decoded pixels back in (pinned) host memory, the reference's
loadFrame contract (Decoder.cpp:221-230: outData is a host vector) as a batch.  The device -> host side sends runs of frames as
one copy (back to back), one 2-D copy (one size at a constant pitch on both sides) or frame by frame: every shape, both
formats, chunk boundaries (chunks start small and double), all compared with the oracle."""
import ctypes

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from motioncam_decoder_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _host_out(ctx, frames, dev_offsets, host_offsets, dev_bytes, host_bytes):
    """frames: (stream, w, h, type, expected image); offsets in bytes into one device / one pinned host output area."""
    from motioncam_decoder_b200 import capi
    total = sum((len(f[0]) + 255) & ~255 for f in frames)
    ring_ptr, ring = ctx.pinned_array(total + 256)
    out_ptr, out = ctx.pinned_array(host_bytes)
    out[:] = 0xA5
    dev = ctx.device_alloc(dev_bytes)
    items, off = [], 0
    for (s, w, h, ct, _), do in zip(frames, dev_offsets):
        ring[off:off + len(s)] = s
        items.append((ring_ptr + off, len(s), w, h, ct, dev + do, w * h))
        off += (len(s) + 255) & ~255
    descs, n = capi.Context.make_descs(items)
    host_dst = (ctypes.c_void_p * n)(*[out_ptr + ho for ho in host_offsets])
    for _ in range(2):                                             # the second call reuses the plans of the slots
        ctx.decode_batch_host_out(descs, host_dst, n)
        written, status = ctx.batch_wait(n)
        assert not any(status), status
        for i, (s, w, h, ct, img) in enumerate(frames):
            assert written[i] == w * h
            got = out[host_offsets[i]:host_offsets[i] + 2 * w * h].view(np.uint16).reshape(h, w)
            assert np.array_equal(got, img), f"frame {i} ({w}x{h}, type {ct}) differs in host memory"
        out[:] = 0x5A
    ctx.device_free(dev)
    ctx.pinned_free(ring_ptr)
    ctx.pinned_free(out_ptr)


def _frames(n, sizes, legacy_every=0):
    from motioncam_decoder_b200 import capi, testvec as tv
    out = []
    for k in range(n):
        w, h = sizes[k % len(sizes)]
        img = tv.gen_photon(w, h, 4095, seed=900 + k)
        if legacy_every and k % legacy_every == legacy_every - 1:
            s, ct = tv.encode_legacy(img, seed=k), capi.COMPRESSION_LEGACY
            n_or, want = ol.oracle_decode_legacy(s, w, h)
        else:
            s, ct = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=k), capi.COMPRESSION_CURRENT
            n_or, want = ol.oracle_decode(s, w, h)
        assert n_or == w * h and np.array_equal(want, img)
        out.append((s, w, h, ct, img))
    return out


def test_host_out_back_to_back(ctx):
    """Frames of one size back to back on both sides: one plain copy per chunk."""
    fr = _frames(12, [(640, 64)])
    fb = 2 * 640 * 64
    offs = [k * fb for k in range(12)]
    _host_out(ctx, fr, offs, offs, 12 * fb, 12 * fb)


def test_host_out_constant_pitch(ctx):
    """One size, padded pitches that differ between device and host: one 2-D copy per chunk."""
    fr = _frames(16, [(1000, 32)], legacy_every=4)
    fb = 2 * 1000 * 32
    dp, hp = fb + 512, fb + 4096
    _host_out(ctx, fr, [k * dp for k in range(16)], [k * hp for k in range(16)], 16 * dp, 16 * hp)


def test_host_out_mixed_sizes_and_orders(ctx):
    """Sizes change from frame to frame and the host buffers come in reverse order: runs of one, and 2-D runs that break."""
    sizes = [(640, 64), (640, 64), (640, 64), (328, 16), (1928, 8), (1928, 8)]
    fr = _frames(18, sizes, legacy_every=5)
    slot = 2 * 1928 * 64
    dev_offs = [k * slot for k in range(18)]
    host_offs = [(17 - k) * slot for k in range(18)]
    _host_out(ctx, fr, dev_offs, host_offs, 18 * slot, 18 * slot)


def test_host_out_many_chunks(ctx):
    """More input than the first (small) chunks hold: 1080p frames across several doubling chunks, constant pitch."""
    from motioncam_decoder_b200 import capi, testvec as tv
    imgs = [tv.gen_photon(1920, 1080, 4095, seed=950 + k) for k in range(3)]
    streams = [tv.encode_current(im) for im in imgs]
    n = 40                                                         # ~80 MB of input: chunks of 8, 16, 32, 64 MB
    fr = [(streams[k % 3], 1920, 1080, capi.COMPRESSION_CURRENT, imgs[k % 3]) for k in range(n)]
    pitch = 2 * 1920 * 1080 + 256
    offs = [k * pitch for k in range(n)]
    _host_out(ctx, fr, offs, offs, n * pitch, n * pitch)
