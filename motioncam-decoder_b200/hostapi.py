"""Python view of the drop-in C++ API (motioncam::Decoder, motioncam::raw::Decode*) through the flat C wrappers
of csrc/decoder_cwrap.inc.  The same class drives this repo's library (prefix ``mcb200_``) and -- in tests only --
the compiled reference (prefix ``mcref_``, oracle/_ref/libmcraw_ref.so), so both are exercised by identical code.
"""
import ctypes
import json

import numpy as np

from . import _lib


class DecoderError(RuntimeError):
    """An exception thrown by the C++ Decoder (message = e.what())."""


def _bind(c, prefix):
    vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
    f = lambda name: getattr(c, prefix + name)  # noqa: E731
    for name in ("decode", "decode_legacy"):
        f(name).argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, sz]
        f(name).restype = sz
    f("decoder_open").argtypes = [ctypes.c_char_p, ctypes.c_char_p, sz]
    f("decoder_open").restype = vp
    f("decoder_open_fp").argtypes = [ctypes.c_char_p, ctypes.c_char_p, sz]
    f("decoder_open_fp").restype = vp
    f("decoder_open_fp_at").argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, sz]
    f("decoder_open_fp_at").restype = vp
    f("decoder_close").argtypes = [vp]
    f("decoder_close").restype = None
    f("decoder_last_error").argtypes = [vp]
    f("decoder_last_error").restype = ctypes.c_char_p
    f("decoder_num_frames").argtypes = [vp]
    f("decoder_num_frames").restype = i64
    f("decoder_frames").argtypes = [vp, ctypes.POINTER(i64)]
    f("decoder_frames").restype = None
    for name in ("decoder_container_metadata", "decoder_frame_metadata"):
        f(name).argtypes = [vp, ctypes.c_char_p, sz]
        f(name).restype = sz
    f("decoder_load_frame").argtypes = [vp, i64]
    f("decoder_load_frame").restype = i64
    f("decoder_frame_data").argtypes = [vp]
    f("decoder_frame_data").restype = vp
    f("decoder_audio_sample_rate").argtypes = [vp]
    f("decoder_audio_channels").argtypes = [vp]
    f("decoder_load_audio").argtypes = [vp, ctypes.c_int]
    f("decoder_load_audio").restype = i64
    for name in ("decoder_audio_chunk_timestamp", "decoder_audio_chunk_samples"):
        f(name).argtypes = [vp, i64]
        f(name).restype = i64
    f("decoder_audio_chunk_data").argtypes = [vp, i64]
    f("decoder_audio_chunk_data").restype = vp
    f("write_dng").argtypes = [ctypes.c_char_p, vp, sz, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, sz]
    f("write_audio").argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(i64), i64, ctypes.c_char_p, sz]
    if prefix == "mcb200_":
        c.mcb200_export_clip.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, sz]
        c.mcb200_export_clip.restype = i64
        c.mcb200_decoder_load_frames.argtypes = [vp, ctypes.POINTER(i64), i64, ctypes.POINTER(vp), ctypes.POINTER(i64)]
        c.mcb200_decoder_load_frames.restype = i64
        c.mcb200_decoder_load_frames_to_device.argtypes = [vp, ctypes.POINTER(i64), i64, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_uint64)]
        c.mcb200_decoder_load_frames_to_device.restype = i64
        c.mcb200_decoder_frame_metadata_at.argtypes = [i64, ctypes.c_char_p, sz]
        c.mcb200_decoder_feed.argtypes = [vp, ctypes.c_char_p, sz]
        c.mcb200_decoder_locate.argtypes = [vp, i64, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_uint32)]
        c.mcb200_decoder_feed.restype = sz
        c.mcb200_decoder_frame_metadata_at.restype = sz
    return c


_libs = {}


def library(path=None, prefix="mcb200_"):
    key = (path, prefix)
    if key not in _libs:
        c = _lib.load(_lib.LIB_DROPIN) if path is None else ctypes.CDLL(path)
        _libs[key] = _bind(c, prefix)
    return _libs[key]


def raw_decode(stream, width, height, legacy=False, lib=None, prefix="mcb200_", fill=0xA5A5):
    """motioncam::raw::Decode / DecodeLegacy: host bytes in, (elements written, host uint16 image) out."""
    c = lib or library()
    src = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.full(width * height + 64, fill, dtype=np.uint16)
    fn = getattr(c, prefix + ("decode_legacy" if legacy else "decode"))
    n = fn(out.ctypes.data, width, height, src.ctypes.data, src.size)
    assert np.all(out[width * height:] == fill), "decoder wrote past width*height"
    return int(n), out[:width * height].reshape(height, width)


def write_dng(path, pixels, frame_metadata, container_metadata, lib=None, prefix="mcb200_"):
    """motioncam::writeDng (include/motioncam/Export.hpp; the reference's example.cpp:55-139): one decoded frame
    (uint16 array or its bytes) + its metadata -> an uncompressed CFA DNG at `path`."""
    c = lib or library()
    px = np.ascontiguousarray(pixels).view(np.uint8).reshape(-1)
    err = ctypes.create_string_buffer(1024)
    rc = getattr(c, prefix + "write_dng")(str(path).encode(), px.ctypes.data, px.size, json.dumps(frame_metadata).encode(),
                                          json.dumps(container_metadata).encode(), err, len(err))
    if rc != 0:
        raise DecoderError(err.value.decode("utf-8", "replace"))


def write_audio(path, sample_rate_hz, channels, chunks, lib=None, prefix="mcb200_"):
    """motioncam::writeAudio (example.cpp:27-53): int16 chunks -> one 16-bit PCM WAV."""
    c = lib or library()
    chunks = [np.ascontiguousarray(x, dtype=np.int16).reshape(-1) for x in chunks]
    flat = np.concatenate(chunks) if chunks else np.zeros(0, np.int16)
    offs = np.zeros(len(chunks) + 1, dtype=np.int64)
    if chunks:
        offs[1:] = np.cumsum([x.size for x in chunks])
    err = ctypes.create_string_buffer(1024)
    rc = getattr(c, prefix + "write_audio")(str(path).encode(), int(sample_rate_hz), int(channels), flat.ctypes.data,
                                            offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(chunks), err, len(err))
    if rc != 0:
        raise DecoderError(err.value.decode("utf-8", "replace"))


def export_clip(path, out_dir, num_frames=-1, batch=16, writer_threads=4, audio=True, stats=None):
    """motioncam::exportClip: the reference's example program (audio.wav + frame_%06d.dng) on the batched B200 decode.
    `stats` (a dict) receives motioncam::ExportStats."""
    c = library()
    err = ctypes.create_string_buffer(1024)
    st = (ctypes.c_double * 6)()
    n = c.mcb200_export_clip(str(path).encode(), str(out_dir).encode(), int(num_frames), int(batch), int(writer_threads),
                             1 if audio else 0, st, err, len(err))
    if n < 0:
        raise DecoderError(err.value.decode("utf-8", "replace"))
    if stats is not None:
        stats.update(zip(("total_s", "open_audio_s", "decode_s", "first_batch_s", "writer_wait_s", "steady_s"), (round(x, 4) for x in st)))
    return int(n)


class Decoder:
    """motioncam::Decoder (Decoder.hpp:47-73)."""

    def __init__(self, path, lib=None, prefix="mcb200_", via_file_handle=False, handle_offset=0):
        """via_file_handle: construct through Decoder(FILE*) (the wrapper fopens `path` and seeks to handle_offset;
        path=None passes a null handle)."""
        self._c = lib or library()
        self._p = prefix
        err = ctypes.create_string_buffer(1024)
        if via_file_handle:
            self._h = self._f("decoder_open_fp_at")(None if path is None else str(path).encode(), handle_offset, err, len(err))
        else:
            self._h = self._f("decoder_open")(str(path).encode(), err, len(err))
        if not self._h:
            raise DecoderError(err.value.decode("utf-8", "replace"))

    def _f(self, name):
        return getattr(self._c, self._p + name)

    def close(self):
        if getattr(self, "_h", None):
            self._f("decoder_close")(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _text(self, name):
        n = self._f(name)(self._h, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        self._f(name)(self._h, buf, n + 1)
        return buf.value.decode()

    def _raise(self):
        raise DecoderError(self._f("decoder_last_error")(self._h).decode("utf-8", "replace"))

    def get_container_metadata(self):
        return json.loads(self._text("decoder_container_metadata"))

    def get_frames(self):
        n = self._f("decoder_num_frames")(self._h)
        arr = (ctypes.c_int64 * max(1, n))()
        self._f("decoder_frames")(self._h, arr)
        return list(arr[:n])

    def load_frame(self, timestamp):
        """-> (bytes-sized np.uint8 array of 2*width*height, frame metadata dict)."""
        n = self._f("decoder_load_frame")(self._h, int(timestamp))
        if n < 0:
            self._raise()
        ptr = self._f("decoder_frame_data")(self._h)
        data = np.frombuffer((ctypes.c_uint8 * n).from_address(ptr), dtype=np.uint8).copy() if n else np.empty(0, np.uint8)
        return data, json.loads(self._text("decoder_frame_metadata"))

    def load_frames(self, timestamps):
        """Batched addition (this repo's library only): -> list of np.uint8 arrays."""
        n = len(timestamps)
        ts = (ctypes.c_int64 * max(1, n))(*timestamps)
        ptrs = (ctypes.c_void_p * max(1, n))()
        sizes = (ctypes.c_int64 * max(1, n))()
        r = self._c.mcb200_decoder_load_frames(self._h, ts, n, ptrs, sizes)
        if r < 0:
            self._raise()
        return [np.frombuffer((ctypes.c_uint8 * sizes[i]).from_address(ptrs[i]), dtype=np.uint8).copy() for i in range(n)]

    def load_frames_to_device(self, timestamps, dst_ptrs, capacities, want_metadata=False):
        """Batched addition: decode straight into device buffers (dst_ptrs[i]: device pointer, capacities[i]: uint16
        elements).  Returns the list of frame metadata dicts when asked for it."""
        n = len(timestamps)
        ts = (ctypes.c_int64 * max(1, n))(*timestamps)
        ptrs = (ctypes.c_void_p * max(1, n))(*dst_ptrs)
        caps = (ctypes.c_uint64 * max(1, n))(*capacities)
        if self._c.mcb200_decoder_load_frames_to_device(self._h, ts, n, ptrs, caps) < 0:
            self._raise()
        if not want_metadata:
            return None
        out = []
        for i in range(n):
            k = self._c.mcb200_decoder_frame_metadata_at(i, None, 0)
            buf = ctypes.create_string_buffer(k + 1)
            self._c.mcb200_decoder_frame_metadata_at(i, buf, k + 1)
            out.append(json.loads(buf.value.decode()))
        return out

    def locate_frame(self, timestamp):
        """Decoder::locateFrame -> (file offset of the compressed bytes, their size)."""
        off, size = ctypes.c_int64(), ctypes.c_uint32()
        if self._c.mcb200_decoder_locate(self._h, int(timestamp), ctypes.byref(off), ctypes.byref(size)) != 0:
            self._raise()
        return off.value, size.value

    def feed_description(self):
        """How load_frames_to_device moves the compressed bytes (Decoder::feedDescription)."""
        buf = ctypes.create_string_buffer(512)
        self._c.mcb200_decoder_feed(self._h, buf, len(buf))
        return buf.value.decode("utf-8", "replace")

    def audio_sample_rate_hz(self):
        v = self._f("decoder_audio_sample_rate")(self._h)
        if v < 0:
            self._raise()
        return v

    def num_audio_channels(self):
        v = self._f("decoder_audio_channels")(self._h)
        if v < 0:
            self._raise()
        return v

    def load_audio(self, use_loader=False):
        """-> list of (timestamp_ns or -1, np.int16 array).  use_loader: the AudioChunkLoader::next loop."""
        n = self._f("decoder_load_audio")(self._h, int(use_loader))
        if n < 0:
            self._raise()
        out = []
        for i in range(n):
            ns = self._f("decoder_audio_chunk_samples")(self._h, i)
            ptr = self._f("decoder_audio_chunk_data")(self._h, i)
            data = np.frombuffer((ctypes.c_int16 * ns).from_address(ptr), dtype=np.int16).copy() if ns else np.empty(0, np.int16)
            out.append((self._f("decoder_audio_chunk_timestamp")(self._h, i), data))
        return out
