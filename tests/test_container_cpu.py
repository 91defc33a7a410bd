"""CPU: the host side of the path -- .mcraw container open / index / audio pass-through of this repo's drop-in
motioncam::Decoder against the compiled reference Decoder (oracle/_ref) on synthetic files written by
testvec.write_mcraw.  No frame is decoded here (that needs the GPU: tests/test_gpu_dropin.py)."""
import os
import struct

import numpy as np
import pytest

import oracle_lib as ol
from motioncam_decoder_b200 import _lib, hostapi, testvec as tv

needs_dropin = pytest.mark.skipif(not os.path.exists(_lib.LIB_DROPIN), reason="drop-in library not built (needs nvcc)")
pytestmark = needs_dropin


def _ref_lib():
    return hostapi.library(ol.REF_SO, "mcref_")


def _clip(tmp_path, name="clip.mcraw", order=(2, 0, 1, 3), with_audio=True):
    """Four tiny frames (both formats) written out of timestamp order + audio chunks with and without metadata."""
    frames = []
    for i, k in enumerate(order):
        w, h = 64 + 32 * k, 8
        img = tv.gen_photon(w, h, 1023, seed=10 + k)
        legacy = bool(k & 1)
        frames.append({"timestamp": 1000 + 10 * k, "data": tv.encode_legacy(img) if legacy else tv.encode_current(img),
                       "width": w, "height": h, "compressionType": 6 if legacy else 7})
    audio = []
    if with_audio:
        rng = np.random.default_rng(5)
        audio = [(123456789, rng.integers(-32768, 32767, 960, dtype=np.int16)),
                 (None, rng.integers(-32768, 32767, 481, dtype=np.int16)),      # odd sample count, no metadata item
                 (223456789, rng.integers(-32768, 32767, 2, dtype=np.int16))]
    path = str(tmp_path / name)
    tv.write_mcraw(path, frames, audio)
    return path, frames, audio


def _open_both(path):
    ours = hostapi.Decoder(path)
    ref = hostapi.Decoder(path, lib=_ref_lib(), prefix="mcref_") if ol.have_ref() else None
    return ours, ref


def test_open_index_metadata(tmp_path):
    path, frames, _ = _clip(tmp_path)
    ours, ref = _open_both(path)
    assert ours.get_frames() == sorted(f["timestamp"] for f in frames)      # Decoder.cpp:266-279
    assert ours.get_container_metadata() == tv.DEFAULT_CONTAINER_METADATA
    assert ours.audio_sample_rate_hz() == 48000 and ours.num_audio_channels() == 2
    assert ours.feed_description() == "pread -> pinned ring"
    if ref:
        assert ours.get_frames() == ref.get_frames()
        assert ours.get_container_metadata() == ref.get_container_metadata()
        assert (ours.audio_sample_rate_hz(), ours.num_audio_channels()) == (ref.audio_sample_rate_hz(), ref.num_audio_channels())


@pytest.mark.parametrize("use_loader", [False, True])
def test_audio_pass_through(tmp_path, use_loader):
    path, _, audio = _clip(tmp_path)
    ours, ref = _open_both(path)
    got = ours.load_audio(use_loader)
    assert len(got) == len(audio)
    for (ts, data), (want_ts, want) in zip(got, audio):
        assert ts == (-1 if want_ts is None else want_ts)                   # Decoder.cpp:58-70
        assert data.size == (want.nbytes + 1) // 2 and np.array_equal(data[:want.size], want)
    if ref:
        want = ref.load_audio(use_loader)
        assert [(t, d.tobytes()) for t, d in got] == [(t, d.tobytes()) for t, d in want]
    if use_loader:
        assert ours.load_audio(True) == []      # the loader's position persists (Decoder.cpp:83-93)
        if ref:
            assert ref.load_audio(True) == []


def test_no_audio_and_empty_index(tmp_path):
    path, _, _ = _clip(tmp_path, with_audio=False)
    ours, ref = _open_both(path)
    assert ours.load_audio() == [] and (ref is None or ref.load_audio() == [])
    p2 = str(tmp_path / "empty.mcraw")
    tv.write_mcraw(p2, [], [])
    ours, ref = _open_both(p2)
    assert ours.get_frames() == [] and (ref is None or ref.get_frames() == [])


def _expect_error(fn, ref_fn, text=None):
    with pytest.raises(hostapi.DecoderError) as e:
        fn()
    if text is not None:
        assert str(e.value) == text
    if ref_fn is not None:
        with pytest.raises(hostapi.DecoderError) as r:
            ref_fn()
        assert str(e.value) == str(r.value), "error text differs from the reference"


def test_error_messages_match_reference(tmp_path):
    path, frames, _ = _clip(tmp_path)
    raw = open(path, "rb").read()
    have_ref = ol.have_ref()

    def opener(p, ref):
        return (lambda: hostapi.Decoder(p, lib=_ref_lib(), prefix="mcref_")) if ref else (lambda: hostapi.Decoder(p))

    def case(name, data, text):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(data)
        _expect_error(opener(p, False), opener(p, True) if have_ref else None, text)

    case("badver.mcraw", raw[:7] + b"\x02" + raw[8:], "Invalid container version")          # Decoder.cpp:123-124
    case("badid.mcraw", b"NOTION " + raw[7:], "Invalid header id")                            # :126-127
    case("badmeta.mcraw", raw[:8] + struct.pack("<I", 2) + raw[12:], "Invalid camera metadata")   # :133-134
    case("short.mcraw", raw[:20], None)                                                       # reads fail
    case("badtrailer.mcraw", raw[:-24] + struct.pack("<I", 3) + raw[-20:], "Invalid file")    # :245-246
    case("badmagic.mcraw", raw[:-16] + struct.pack("<i", 0x1234) + raw[-12:], "Corrupted file")   # :252-253
    missing = str(tmp_path / "does_not_exist.mcraw")
    _expect_error(opener(missing, False), opener(missing, True) if have_ref else None, "Failed to open " + missing)

    ours, ref = _open_both(path)
    _expect_error(lambda: ours.load_frame(42), (lambda: ref.load_frame(42)) if ref else None,
                  "Frame not found (timestamp: 42)")                                          # :185-186


def test_missing_metadata_keys_are_exceptions(tmp_path):
    """The reference reads these keys with json::operator[] on a const object (Decoder.cpp:161-167, :216-218): an
    assertion / undefined behaviour when the key is absent.  Here: IOException."""
    img = tv.gen_photon(64, 8, 1023, seed=1)
    frame = {"timestamp": 5, "data": tv.encode_current(img), "height": 8, "compressionType": 7}      # no "width"
    meta = {k: v for k, v in tv.DEFAULT_CONTAINER_METADATA.items() if k != "extraData"}
    path = str(tmp_path / "nokeys.mcraw")
    tv.write_mcraw(path, [frame], [], container_metadata=meta)
    d = hostapi.Decoder(path)
    with pytest.raises(hostapi.DecoderError, match='Invalid frame metadata \\(no "width"\\)'):
        d.load_frame(5)
    with pytest.raises(hostapi.DecoderError, match='Invalid camera metadata \\(no "extraData"\\)'):
        d.audio_sample_rate_hz()
    with pytest.raises(hostapi.DecoderError, match='Invalid camera metadata \\(no "extraData"\\)'):
        d.num_audio_channels()


def test_file_handle_constructor(tmp_path):
    """Decoder(FILE*) (Decoder.cpp:97-102): same view of the file as Decoder(path); a null handle is "Invalid file";
    the Decoder owns the handle (closed with it)."""
    path, frames, audio = _clip(tmp_path)
    fds_before = len(os.listdir("/proc/self/fd"))
    ours = hostapi.Decoder(path, via_file_handle=True)
    assert ours.get_frames() == sorted(f["timestamp"] for f in frames)
    assert len(ours.load_audio()) == len(audio)
    assert len(os.listdir("/proc/self/fd")) == fds_before + 1
    ours.close()
    assert len(os.listdir("/proc/self/fd")) == fds_before
    with pytest.raises(hostapi.DecoderError) as e:
        hostapi.Decoder(None, via_file_handle=True)
    assert str(e.value) == "Invalid file"
    bad = str(tmp_path / "bad.mcraw")
    with open(bad, "wb") as f:
        f.write(b"NOTION " + open(path, "rb").read()[7:])
    with pytest.raises(hostapi.DecoderError, match="Invalid header id"):
        hostapi.Decoder(bad, via_file_handle=True)
    # a constructor that throws leaves the handle with the caller (the reference never closes it on that path): the
    # wrapper, being the caller, closes it -- a Decoder that had closed it as well would have made this a double close
    assert len(os.listdir("/proc/self/fd")) == fds_before
    # the container header is read from the handle's position (Decoder.cpp:116-141 never seeks), everything else by absolute offset
    shifted = str(tmp_path / "shifted.mcraw")
    tv.write_mcraw(shifted, frames, audio, prefix=b"\x00" * 4096 + b"junk in front")
    for kw in ([dict()] + ([dict(lib=_ref_lib(), prefix="mcref_")] if ol.have_ref() else [])):
        with pytest.raises(hostapi.DecoderError):
            hostapi.Decoder(shifted, via_file_handle=True, **kw)                       # position 0: not a container header
        d = hostapi.Decoder(shifted, via_file_handle=True, handle_offset=4096 + 13, **kw)
        assert d.get_frames() == sorted(f["timestamp"] for f in frames)
        assert len(d.load_audio()) == len(audio)
        d.close()
    assert len(os.listdir("/proc/self/fd")) == fds_before
    if ol.have_ref():
        ref = hostapi.Decoder(path, lib=_ref_lib(), prefix="mcref_", via_file_handle=True)
        assert ref.get_frames() == sorted(f["timestamp"] for f in frames)
        ref.close()
        with pytest.raises(hostapi.DecoderError) as e:
            hostapi.Decoder(None, lib=_ref_lib(), prefix="mcref_", via_file_handle=True)
        assert str(e.value) == "Invalid file"


def test_duplicate_timestamps(tmp_path):
    """getFrames() lists every index entry, duplicates included, in timestamp order (Decoder.cpp:266-279)."""
    frames = []
    for i, ts in enumerate([30, 10, 30, 20, 10]):
        img = tv.gen_photon(64, 4 + 4 * i, 1023, seed=i)
        frames.append({"timestamp": ts, "data": tv.encode_current(img), "width": 64, "height": 4 + 4 * i, "compressionType": 7})
    path = str(tmp_path / "dup.mcraw")
    tv.write_mcraw(path, frames, [])
    ours, ref = _open_both(path)
    assert ours.get_frames() == [10, 10, 20, 30, 30]
    # loadFrame finds the FIRST entry of a timestamp; here "first" is index order (stable sort), in the reference it is
    # whatever std::sort left in front (unspecified), so only the documented rule is checked on this side
    raw = open(path, "rb").read()
    for ts, first in ((30, frames[0]), (10, frames[1]), (20, frames[3])):
        off, size = ours.locate_frame(ts)
        assert size == len(first["data"]) and raw[off:off + size] == bytes(first["data"].tobytes()), ts
    with pytest.raises(hostapi.DecoderError, match="Frame not found"):
        ours.locate_frame(11)
    if ref:
        assert ref.get_frames() == ours.get_frames()
        data, meta = ref.load_frame(20)                      # the reference decodes on the CPU
        assert data.size == 2 * 64 * frames[3]["height"]


def test_exports_reference_symbols():
    """The drop-in library exports the reference's mangled codec symbols (RawData.hpp:25-37)."""
    import ctypes
    c = ctypes.CDLL(_lib.LIB_DROPIN)
    for sym in ("_ZN9motioncam3raw6DecodeEPtiiPKhm", "_ZN9motioncam3raw12DecodeLegacyEPtiiPKhm"):
        assert hasattr(c, sym), sym


def test_cli_synth_and_info(tmp_path):
    """tools/mcraw_tool.py: write a synthetic clip, read its index back through the drop-in Decoder (CPU only)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = str(tmp_path / "cli.mcraw")
    tool = os.path.join(root, "tools", "mcraw_tool.py")
    subprocess.run([sys.executable, tool, "synth", path, "--frames", "3", "--width", "128", "--height", "8", "--legacy"], check=True)
    out = subprocess.run([sys.executable, tool, "info", path], check=True, capture_output=True, text=True).stdout
    info = json.loads(out)
    assert info["frames"] == 3 and info["audio_chunks"] == 4 and info["audio_sample_rate_hz"] == 48000
