"""CPU test-vector tools: synthetic Bayer images, the inverse encoder for both frame formats and a
``.mcraw`` container writer (inverse of the reference reader, /root/reference/lib/Decoder.cpp:116-315,
PODs /root/reference/lib/include/motioncam/Container.hpp:23-71).

Everything here runs on the host and is used to *produce inputs*; nothing here decodes.
"""
import ctypes
import json
import struct

import numpy as np

from . import _lib

_c = None


def _tools():
    global _c
    if _c is None:
        c = _lib.load(_lib.LIB_TOOLS)
        u16p, u8p = ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint8)
        c.mcraw_gen_photon.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        c.mcraw_gen_flatnoise.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        c.mcraw_gen_uniform.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        c.mcraw_gen_forced_widths.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                              ctypes.c_int, ctypes.c_uint64]
        for f in (c.mcraw_encode_current_bound, c.mcraw_encode_legacy_bound):
            f.argtypes = [ctypes.c_int, ctypes.c_int]
            f.restype = ctypes.c_size_t
        c.mcraw_encode_current.argtypes = [u16p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        c.mcraw_encode_current.restype = ctypes.c_size_t
        c.mcraw_assemble_current.argtypes = [ctypes.c_int, ctypes.c_int, u16p, u16p, u8p, ctypes.c_size_t,
                                             ctypes.c_uint64]
        c.mcraw_assemble_current.restype = ctypes.c_size_t
        c.mcraw_encode_legacy.argtypes = [u16p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        c.mcraw_encode_legacy.restype = ctypes.c_size_t
        c.mcraw_assemble_legacy.argtypes = [ctypes.c_int, ctypes.c_int, u8p, u16p, u8p, ctypes.c_size_t,
                                            ctypes.c_uint64]
        c.mcraw_assemble_legacy.restype = ctypes.c_size_t
        c.mcraw_fnv1a64.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        c.mcraw_fnv1a64.restype = ctypes.c_uint64
        _c = c
    return _c


def _p16(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16))


def _p8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


# ------------------------------------------------------------------------------------------------
# images
# ------------------------------------------------------------------------------------------------
def gen_photon(width, height, maxval=1023, seed=1234):
    img = np.empty((height, width), dtype=np.uint16)
    _tools().mcraw_gen_photon(_p16(img), width, height, maxval, seed)
    return img


def gen_flatnoise(width, height, cell=256, seed=1):
    img = np.empty((height, width), dtype=np.uint16)
    _tools().mcraw_gen_flatnoise(_p16(img), width, height, cell, seed)
    return img


def gen_uniform(width, height, lo=0, hi=65535, seed=1):
    img = np.empty((height, width), dtype=np.uint16)
    _tools().mcraw_gen_uniform(_p16(img), width, height, lo, hi, seed)
    return img


def gen_forced_widths(width, height, widths, seed=1):
    img = np.empty((height, width), dtype=np.uint16)
    arr = (ctypes.c_int * len(widths))(*widths)
    _tools().mcraw_gen_forced_widths(_p16(img), width, height, arr, len(widths), seed)
    return img


# ------------------------------------------------------------------------------------------------
# encoders
# ------------------------------------------------------------------------------------------------
POLICY_MINIMAL, POLICY_ALIASES, POLICY_FORCE = 0, 1, 2


def encode_current(img, policy=POLICY_MINIMAL, policy_arg=0, ref_wrap=False, seed=1):
    """compressionType 7 stream for a (height, width) uint16 image -> np.uint8 array."""
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    c = _tools()
    cap = c.mcraw_encode_current_bound(w, h)
    out = np.empty(cap, dtype=np.uint8)
    n = c.mcraw_encode_current(_p16(img), w, h, _p8(out), cap, policy, policy_arg, int(ref_wrap), seed)
    if n == 0:
        raise ValueError("encode_current failed")
    return out[:n].copy()


def assemble_current(enc_width, enc_height, bits, refs, seed=1):
    """Well-formed compressionType 7 stream from explicit bits[]/refs[] and random payload bytes."""
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    refs = np.ascontiguousarray(refs, dtype=np.uint16)
    n_blocks = enc_width * enc_height // 64
    assert bits.size == n_blocks and refs.size == n_blocks
    c = _tools()
    cap = c.mcraw_encode_current_bound(enc_width, enc_height)
    out = np.empty(cap, dtype=np.uint8)
    n = c.mcraw_assemble_current(enc_width, enc_height, _p16(bits), _p16(refs), _p8(out), cap, seed)
    if n == 0:
        raise ValueError("assemble_current failed")
    return out[:n].copy()


def pad_meta_current(stream, pad_bits=0, pad_refs=0, fill=0xEE):
    """Insert pad_bits filler bytes in front of the "bits" metadata stream and pad_refs in front of the "refs" stream of a
    compressionType 7 frame and patch bitsOffset / refsOffset (bytes 8..15) accordingly.  The result decodes to the same
    image in the reference (RawData.cpp:547-560 only needs the offsets to be <= len); odd pads give metadata streams at
    odd byte offsets, which no known encoder writes."""
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    ew, eh, boff, roff = (int(v) for v in np.frombuffer(stream[:16].tobytes(), dtype="<u4"))
    assert 16 <= boff <= roff <= stream.size, "expected payload | bits stream | refs stream"
    out = np.concatenate([stream[:boff], np.full(pad_bits, fill, np.uint8), stream[boff:roff],
                          np.full(pad_refs, fill, np.uint8), stream[roff:]])
    out[8:16] = np.frombuffer(np.array([boff + pad_bits, roff + pad_bits + pad_refs], dtype="<u4").tobytes(), dtype=np.uint8)
    return out


def encoded_width_of(stream):
    """encodedWidth from the header of a compressionType 7 frame (RawData.cpp:500-524)."""
    return int(np.frombuffer(np.ascontiguousarray(stream[:4], dtype=np.uint8).tobytes(), dtype="<u4")[0])


def encode_legacy(img, policy=POLICY_MINIMAL, policy_arg=0, trailer_records=0, seed=1):
    """compressionType 6 stream for a (height, width) uint16 image -> np.uint8 array."""
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    c = _tools()
    cap = c.mcraw_encode_legacy_bound(w, h) + 5 * trailer_records
    out = np.empty(cap, dtype=np.uint8)
    n = c.mcraw_encode_legacy(_p16(img), w, h, _p8(out), cap, policy, policy_arg, trailer_records, seed)
    if n == 0:
        raise ValueError("encode_legacy failed")
    return out[:n].copy()


def assemble_legacy(width, height, nibbles, refs12, seed=1):
    nibbles = np.ascontiguousarray(nibbles, dtype=np.uint8)
    refs12 = np.ascontiguousarray(refs12, dtype=np.uint16)
    c = _tools()
    cap = c.mcraw_encode_legacy_bound(width, height)
    out = np.empty(cap, dtype=np.uint8)
    n = c.mcraw_assemble_legacy(width, height, _p8(nibbles), _p16(refs12), _p8(out), cap, seed)
    if n == 0:
        raise ValueError("assemble_legacy failed")
    return out[:n].copy()


def fnv1a64(arr):
    arr = np.ascontiguousarray(arr)
    return int(_tools().mcraw_fnv1a64(arr.ctypes.data, arr.nbytes))


# ------------------------------------------------------------------------------------------------
# .mcraw container writer
# ------------------------------------------------------------------------------------------------
T_BUFFER_INDEX, T_BUFFER_INDEX_DATA, T_BUFFER, T_METADATA, T_AUDIO_INDEX, T_AUDIO_DATA, T_AUDIO_DATA_METADATA = range(7)
INDEX_MAGIC = 0x8A905612

DEFAULT_CONTAINER_METADATA = {
    "blackLevel": [64, 64, 64, 64],
    "whiteLevel": 1023.0,
    "sensorArrangment": "rggb",
    "colorMatrix1": [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
    "colorMatrix2": [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
    "forwardMatrix1": [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
    "forwardMatrix2": [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0],
    "extraData": {"audioSampleRate": 48000, "audioChannels": 2},
}


def _item(t, size):
    return struct.pack("<II", t, size)


def write_mcraw(path, frames, audio_chunks=(), container_metadata=None, index_order=None, prefix=b""):
    """Write a version-3 ``.mcraw`` file.

    frames        list of dicts {timestamp, data (bytes/np.uint8), width, height, compressionType, [extra json keys]}
                  written in list order (need not be timestamp order -- the reader sorts, Decoder.cpp:266-279)
    audio_chunks  list of (timestamp_ns_or_None, np.int16 array); None -> no AUDIO_DATA_METADATA item follows
                  (older files, Decoder.cpp:58-70)
    index_order   optional permutation for the BufferOffset array
    prefix        bytes written in front of the container header
    Layout: Header, METADATA json, frame items (BUFFER + METADATA each), audio items, AUDIO_INDEX,
    BUFFER_INDEX_DATA, BUFFER_INDEX (last 24 bytes, Decoder.cpp:237-264).  The audio index must be reachable by
    walking items forward from the frame with the largest timestamp (Decoder.cpp:281-315), which holds for this
    layout whatever the frame order.
    """
    meta = dict(DEFAULT_CONTAINER_METADATA if container_metadata is None else container_metadata)
    offsets = []
    audio_offsets = []
    with open(path, "wb") as f:
        f.write(prefix)        # foreign bytes in front of the container: Decoder(FILE*) starts at the handle's position, and
                               # every offset stored in the file is absolute (Decoder.cpp:116-141 vs :188,239,258,284)
        f.write(b"MOTION " + bytes([3]))
        mj = json.dumps(meta).encode()
        f.write(_item(T_METADATA, len(mj)))
        f.write(mj)

        def write_audio():
            for ts, samples in audio_chunks:
                samples = np.ascontiguousarray(samples, dtype=np.int16)
                audio_offsets.append((f.tell(), -1 if ts is None else int(ts)))
                f.write(_item(T_AUDIO_DATA, samples.nbytes))
                f.write(samples.tobytes())
                if ts is not None:
                    f.write(_item(T_AUDIO_DATA_METADATA, 8))
                    f.write(struct.pack("<q", int(ts)))

        for i, fr in enumerate(frames):
            data = fr["data"]
            if not isinstance(data, (bytes, bytearray)):
                data = np.ascontiguousarray(data, dtype=np.uint8).tobytes()
            offsets.append((f.tell(), int(fr["timestamp"])))
            f.write(_item(T_BUFFER, len(data)))
            f.write(data)
            fm = {k: v for k, v in fr.items() if k not in ("data", "timestamp")}
            fm.setdefault("asShotNeutral", [1.0, 1.0, 1.0])
            fm["timestamp"] = str(int(fr["timestamp"]))
            fj = json.dumps(fm).encode()
            f.write(_item(T_METADATA, len(fj)))
            f.write(fj)
        write_audio()
        f.write(_item(T_AUDIO_INDEX, 16 + 16 * len(audio_offsets)))
        f.write(struct.pack("<qq", len(audio_offsets), 0))
        for off, ts in audio_offsets:
            f.write(struct.pack("<qq", off, ts))
        if index_order is not None:
            offsets = [offsets[i] for i in index_order]
        f.write(_item(T_BUFFER_INDEX_DATA, 16 * len(offsets)))
        index_data_offset = f.tell()
        for off, ts in offsets:
            f.write(struct.pack("<qq", off, ts))
        f.write(_item(T_BUFFER_INDEX, 16))
        f.write(struct.pack("<iiq", INDEX_MAGIC - (1 << 32), len(offsets), index_data_offset))
    return path
