"""CPU: the table k_units uses to pull single values out of a metadata block (csrc/mcraw_meta_table.h) against an
independent restatement of the block layouts (SURVEY.md appendix A / RawData.cpp:112-408), every header value and sample."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sample(p, hb, i):
    """Sample i = 8j + l of a block stored with header value hb; p = payload bytes (appendix A)."""
    j, l = divmod(i, 8)
    G = lambda m: int(p[8 * m + l])  # noqa: E731
    if hb == 0:
        return 0
    if hb == 1:
        return (G(0) >> j) & 1
    if hb == 2:
        return (G(j >> 2) >> (2 * (j & 3))) & 3
    if hb == 3:
        return [G(0) & 7, (G(0) >> 3) & 7, ((G(0) >> 6) & 3) | (((G(2) >> 6) & 1) << 2), G(1) & 7, (G(1) >> 3) & 7,
                ((G(1) >> 6) & 3) | (((G(2) >> 7) & 1) << 2), G(2) & 7, (G(2) >> 3) & 7][j]
    if hb == 4:
        return (G(j >> 1) >> (4 * (j & 1))) & 15
    if hb == 5:
        if j <= 4:
            return G(j) & 31
        return [((G(0) >> 5) & 7) | (((G(3) >> 5) & 3) << 3), ((G(1) >> 5) & 7) | (((G(4) >> 5) & 3) << 3),
                ((G(2) >> 5) & 7) | (((G(3) >> 7) & 1) << 3) | (((G(4) >> 7) & 1) << 4)][j - 5]
    if hb == 6:
        if j <= 5:
            return G(j) & 63
        a = 0 if j == 6 else 3
        return (G(a) >> 6) | ((G(a + 1) >> 6) << 2) | ((G(a + 2) >> 6) << 4)
    if hb in (7, 8):
        return G(j)
    if hb in (9, 10):
        return (G(j) | (((G(4) >> (2 * j)) & 3) << 8)) if j < 4 else (G(j + 1) | (((G(9) >> (2 * (j - 4))) & 3) << 8))
    return int(p[2 * i]) | (int(p[2 * i + 1]) << 8)


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("meta") / "libmeta_table_check.so")
    subprocess.run(["gcc", "-std=c11", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "motioncam-decoder_b200", "csrc"),
                    "-o", so, os.path.join(ROOT, "tests", "helpers", "meta_table_check.c")], check=True)
    c = ctypes.CDLL(so)
    c.meta_table_sample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    c.meta_table_sample.restype = ctypes.c_uint
    c.meta_table_pair.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    c.meta_table_pair.restype = ctypes.c_uint
    c.meta_len8x4.argtypes = [ctypes.c_uint]
    c.meta_len8x4.restype = ctypes.c_uint
    return c


def test_table_matches_layouts(lib):
    rng = np.random.default_rng(11)
    for hb in range(16):
        for _ in range(8):
            p = rng.integers(0, 256, 144, dtype=np.uint8)
            for i in range(64):
                assert lib.meta_table_sample(p.ctypes.data, hb, i) == _sample(p, hb, i), (hb, i)
            for i in range(0, 64, 2):           # the two-lane form k_units uses (a lane's even / odd block)
                assert lib.meta_table_pair(p.ctypes.data, hb, i) == _sample(p, hb, i) | (_sample(p, hb, i + 1) << 16), (hb, i)


def test_table_matches_the_oracle_decoder():
    """The same layouts end to end: a frame whose blocks all use header hb, decoded by the oracle, equals per-sample extraction."""
    import oracle_lib as ol
    from motioncam_decoder_b200 import testvec as tv
    rng = np.random.default_rng(5)
    for hb in range(17):
        bits = np.full(4, hb, dtype=np.uint16)
        refs = np.zeros(4, dtype=np.uint16)
        s = tv.assemble_current(64, 4, bits, refs, seed=hb)
        n, img = ol.oracle_decode(s, 64, 4)
        assert n == 256
        length = [0, 8, 16, 24, 32, 40, 48, 64, 64, 80, 80, 128, 128, 128, 128, 128, 128][hb]
        for c in range(4):
            p = np.concatenate([s[16 + c * length:16 + (c + 1) * length], np.zeros(160, np.uint8)])
            for i in rng.integers(0, 64, 16):
                y, x = (c >> 1) + 2 * (int(i) >> 5), 2 * (int(i) & 31) + (c & 1)
                assert img[y, x] == _sample(p, hb, int(i)), (hb, c, int(i))


def test_len8x4(lib):
    """Four block lengths per word (k_meta's payload prefix sums) against RawData.cpp:27-45."""
    table = [0, 8, 16, 24, 32, 40, 48, 64, 64, 80, 80, 128, 128, 128, 128, 128, 128]
    rng = np.random.default_rng(3)
    for _ in range(2000):
        vals = [int(x) for x in rng.integers(0, 17, 4)]
        v = vals[0] | (vals[1] << 8) | (vals[2] << 16) | (vals[3] << 24)
        got = lib.meta_len8x4(v)
        assert [(got >> (8 * i)) & 0xFF for i in range(4)] == [table[x] // 8 for x in vals], vals
