"""ctypes access to the parity checkers under oracle/ (tests only).

  oracle_decode / oracle_decode_legacy   the plain-C restatement (oracle/mcraw_oracle.c)
  ref_decode / ref_decode_legacy         the compiled UNMODIFIED reference (oracle/_ref/libmcraw_ref.so), when present
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libmcraw_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmcraw_ref.so")

_u16p = ctypes.POINTER(ctypes.c_uint16)
_u8p = ctypes.POINTER(ctypes.c_uint8)

_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        c = ctypes.CDLL(ORACLE_SO)
        for f in (c.oracle_decode, c.oracle_decode_legacy):
            f.argtypes = [_u16p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_size_t]
            f.restype = ctypes.c_size_t
        _oracle = c
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        c = ctypes.CDLL(REF_SO)
        for f in (c.mcref_decode, c.mcref_decode_legacy):
            f.argtypes = [_u16p, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_size_t]
            f.restype = ctypes.c_size_t
        _ref = c
    return _ref


def _run(fn, stream, width, height, with_cap, fill):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.full(width * height + 64, fill, dtype=np.uint16)
    args = [out.ctypes.data_as(_u16p)]
    if with_cap:
        args.append(width * height)
    args += [width, height, stream.ctypes.data_as(_u8p), stream.size]
    n = fn(*args)
    assert np.all(out[width * height:] == fill), "decoder wrote past width*height"
    return int(n), out[:width * height].reshape(height, width)


def oracle_decode(stream, width, height, fill=0xA5A5):
    return _run(oracle().oracle_decode, stream, width, height, True, fill)


def oracle_decode_legacy(stream, width, height, fill=0xA5A5):
    return _run(oracle().oracle_decode_legacy, stream, width, height, True, fill)


def ref_decode(stream, width, height, fill=0xA5A5):
    return _run(ref().mcref_decode, stream, width, height, False, fill)


def ref_decode_legacy(stream, width, height, fill=0xA5A5):
    return _run(ref().mcref_decode_legacy, stream, width, height, False, fill)
