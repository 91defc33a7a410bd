"""GPU: the drop-in C++ API (motioncam::Decoder::loadFrame / loadFrames, motioncam::raw::Decode*) of this repo
against the compiled reference and the oracle, on synthetic .mcraw files."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
import vectors
from motioncam_decoder_b200 import hostapi, testvec as tv

pytestmark = pytest.mark.gpu


def _ref_lib():
    return hostapi.library(ol.REF_SO, "mcref_")


def _clip(tmp_path, n=10):
    frames, images = [], {}
    sizes = [(1928, 16), (640, 12), (4080, 8), (100, 4), (64, 4), (2048, 24)]
    for k in range(n):
        w, h = sizes[k % len(sizes)]
        legacy = k % 3 == 1
        if legacy and h % 4:
            h += 4 - h % 4
        img = tv.gen_photon(w, h, 4095 if k % 2 else 1023, seed=40 + k)
        ts = 5_000_000 + 33_333 * ((k * 7) % n)          # written out of order
        frames.append({"timestamp": ts, "data": tv.encode_legacy(img) if legacy else tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=k),
                       "width": w, "height": h, "compressionType": 6 if legacy else 7, "iso": 100 + k})
        images[ts] = img
    rng = np.random.default_rng(1)
    audio = [(1000 * i, rng.integers(-3000, 3000, 1920, dtype=np.int16)) for i in range(4)]
    path = str(tmp_path / "clip.mcraw")
    tv.write_mcraw(path, frames, audio)
    return path, frames, images, audio


def test_load_frame_matches_reference(tmp_path):
    path, frames, images, audio = _clip(tmp_path)
    ours = hostapi.Decoder(path)
    ref = hostapi.Decoder(path, lib=_ref_lib(), prefix="mcref_") if ol.have_ref() else None
    stamps = ours.get_frames()
    assert stamps == sorted(images)
    for ts in stamps:
        data, meta = ours.load_frame(ts)
        img = images[ts]
        assert data.size == img.size * 2
        assert np.array_equal(data.view(np.uint16).reshape(img.shape), img), ts
        assert meta["width"] == img.shape[1] and meta["height"] == img.shape[0]
        if ref:
            rdata, rmeta = ref.load_frame(ts)
            assert np.array_equal(data, rdata) and meta == rmeta
    got = ours.load_audio()
    assert [(t, d.tobytes()) for t, d in got] == [(t, np.asarray(d).tobytes()) for t, d in audio]


def test_load_frames_batched(tmp_path):
    path, frames, images, _ = _clip(tmp_path, n=12)
    ours = hostapi.Decoder(path)
    stamps = ours.get_frames()
    order = stamps[::-1] + stamps[:3]                      # any order, repeats allowed
    out = ours.load_frames(order)
    for ts, data in zip(order, out):
        img = images[ts]
        assert np.array_equal(data.view(np.uint16).reshape(img.shape), img), ts
    assert ours.load_frames([]) == []
    with pytest.raises(hostapi.DecoderError, match="Frame not found"):
        ours.load_frames([stamps[0], 1])


def test_raw_decode_symbols():
    """motioncam::raw::Decode / DecodeLegacy (host in, host out) == oracle on the shared vectors."""
    for name, s, w, h, img in vectors.current_vectors(small=True)[::3]:
        n, got = hostapi.raw_decode(s, w, h)
        n_or, want = ol.oracle_decode(s, w, h)
        assert n == n_or and np.array_equal(got, want), name
    for name, s, w, h, img in vectors.legacy_vectors(small=True)[::3]:
        n, got = hostapi.raw_decode(s, w, h, legacy=True)
        n_or, want = ol.oracle_decode_legacy(s, w, h)
        assert n == n_or and np.array_equal(got, want), name


def test_decode_failures_raise_like_reference(tmp_path):
    img = tv.gen_photon(128, 8, 1023, seed=3)
    bad = tv.encode_current(img)
    bad[0:4] = np.frombuffer(np.uint32(96).tobytes(), dtype=np.uint8)       # encodedWidth % 64 != 0 -> Decode returns 0
    frames = [
        {"timestamp": 1, "data": bad, "width": 128, "height": 8, "compressionType": 7},
        {"timestamp": 2, "data": tv.encode_current(img), "width": 128, "height": 8, "compressionType": 5},
        {"timestamp": 3, "data": tv.encode_current(img), "width": 128, "height": 8, "compressionType": 7},
    ]
    path = str(tmp_path / "bad.mcraw")
    tv.write_mcraw(path, frames, [])
    ours = hostapi.Decoder(path)
    ref = hostapi.Decoder(path, lib=_ref_lib(), prefix="mcref_") if ol.have_ref() else None
    for ts, text in [(1, "Failed to uncompress frame"), (2, "Invalid compression type")]:
        with pytest.raises(hostapi.DecoderError) as e:
            ours.load_frame(ts)
        assert str(e.value) == text                                          # Decoder.cpp:225-233
        if ref:
            with pytest.raises(hostapi.DecoderError) as r:
                ref.load_frame(ts)
            assert str(r.value) == text
    data, _ = ours.load_frame(3)
    assert np.array_equal(data.view(np.uint16).reshape(8, 128), img)


def test_load_frames_to_device(tmp_path):
    """Decoder::loadFramesToDevice: file -> pinned ring -> staged H2D -> kernels, decoded frames stay on the GPU."""
    from motioncam_decoder_b200 import capi
    path, frames, images, _ = _clip(tmp_path, n=9)
    ours = hostapi.Decoder(path)
    stamps = ours.get_frames()
    ctx = capi.Context(0)                      # only used here to allocate / read back device memory
    ptrs, caps = [], []
    for ts in stamps:
        img = images[ts]
        ptrs.append(ctx.device_alloc(img.size * 2))
        caps.append(img.size)
    for _ in range(2):                         # second call reuses the Decoder's pinned ring
        metas = ours.load_frames_to_device(stamps, ptrs, caps, want_metadata=True)
        for ts, p, meta in zip(stamps, ptrs, metas):
            img = images[ts]
            out = np.empty(img.shape, dtype=np.uint16)
            ctx.d2h(out, p)
            assert np.array_equal(out, img), ts
            assert meta["width"] == img.shape[1] and meta["height"] == img.shape[0]
    with pytest.raises(hostapi.DecoderError, match="Frame not found"):
        ours.load_frames_to_device([12345], ptrs[:1], caps[:1])
    for p in ptrs:
        ctx.device_free(p)
    ctx.close()


def _platform_takes(mode, path):
    """What this box allows, found out independently of the Decoder: O_DIRECT reads of the file, page-locking a mapping."""
    import mmap
    from motioncam_decoder_b200 import capi
    if mode == "direct":
        try:
            fd = os.open(path, os.O_RDONLY | os.O_DIRECT)
        except OSError:
            return False
        try:
            buf = mmap.mmap(-1, 8192)                     # page-aligned
            return os.preadv(fd, [memoryview(buf)[:4096]], 0) > 0
        except OSError:
            return False
        finally:
            os.close(fd)
    if mode == "mmap":
        ctx = capi.Context(0)
        with open(path, "rb") as f:
            m = mmap.mmap(f.fileno(), 0, prot=mmap.PROT_READ)
            arr = np.frombuffer(m, dtype=np.uint8)
            addr = arr.ctypes.data
            ok = ctx._c.mcraw_host_register(ctx._h, addr, arr.size, 1) == 0
            if ok:
                ctx._c.mcraw_host_unregister(ctx._h, addr)
            del arr
            m.close()
        ctx.close()
        return ok
    return None


def test_feed_cufile(tmp_path):
    """MCRAW_FEED=cufile (GPUDirect Storage through libcufile): the attempt runs under a deadline inside the library, so the
    call always comes back -- with the frames decoded through cuFile, or through the ring and the reason in the text.
    Runs in a child process (tests/helpers/feed_child.py) with a hard limit: a hang is a failure."""
    import json
    import subprocess
    import sys
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "feed_child.py")
    r = subprocess.run([sys.executable, helper, "cufile", str(tmp_path)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    print("feed[cufile]:", out["feed"])
    assert out["frames_ok"]
    feed = out["feed"]
    assert feed.startswith("cuFileRead -> device memory") or feed.startswith("pread -> pinned ring (cuFile") or \
        feed.startswith("pread -> pinned ring (O_DIRECT open refused")


@pytest.mark.parametrize("mode", ["ring", "direct", "mmap"])
def test_feeds(tmp_path, monkeypatch, mode):
    """The feeds of Decoder::loadFramesToDevice (MCRAW_FEED): pipelined pread into the pinned ring (default), O_DIRECT
    reads, cuFile (GPUDirect Storage) into device memory, H2D from a page-locked mapping.  Every feed must deliver the
    same frames; a feed the platform refuses falls back to the ring and says why -- and which of the two it is on
    this box is established independently (no assertion that cannot fail)."""
    from motioncam_decoder_b200 import capi
    if mode != "ring":
        monkeypatch.setenv("MCRAW_FEED", mode)
    monkeypatch.setenv("MCRAW_FEED_CHUNK_KB", "16")                    # several chunks even for this small clip
    path, frames, images, _ = _clip(tmp_path, n=9)
    ours = hostapi.Decoder(path)
    assert ours.feed_description() == "pread -> pinned ring"          # decided at the first device load
    stamps = ours.get_frames()
    ctx = capi.Context(0)
    ptrs = [ctx.device_alloc(images[ts].size * 2) for ts in stamps]
    caps = [images[ts].size for ts in stamps]
    for _ in range(2):
        for p, ts in zip(ptrs, stamps):
            ctx.h2d(p, np.zeros(images[ts].shape, np.uint16))
        ours.load_frames_to_device(stamps, ptrs, caps)
        for ts, p in zip(stamps, ptrs):
            out = np.empty(images[ts].shape, dtype=np.uint16)
            ctx.d2h(out, p)
            assert np.array_equal(out, images[ts]), ts
    feed = ours.feed_description()
    print(f"feed[{mode}]:", feed)
    if mode == "ring":
        assert feed == "pread -> pinned ring, read of chunk c+1 overlapping H2D + decode of chunk c"
    elif mode == "direct":
        if _platform_takes("direct", path):
            assert feed.startswith("O_DIRECT pread -> pinned ring")
        else:
            assert feed.startswith("pread -> pinned ring (O_DIRECT")
    elif mode == "mmap":
        if _platform_takes("mmap", path):
            assert feed.startswith("mmap + cudaHostRegister")
        else:
            assert feed.startswith("pread -> pinned ring (cudaHostRegister refused the mapping")
    ours.close()
    for p in ptrs:
        ctx.device_free(p)
    ctx.close()
