"""CPU: pin the oracle (oracle/mcraw_oracle.c) against the compiled, unmodified reference
(oracle/_ref/libmcraw_ref.so) and check that the test-vector encoder is an exact inverse."""
import numpy as np
import pytest

import oracle_lib as ol
import vectors

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libmcraw_ref.so not built")


@pytest.fixture(scope="module")
def cur():
    return vectors.current_vectors(small=True)


@pytest.fixture(scope="module")
def leg():
    return vectors.legacy_vectors(small=True)


@needs_ref
def test_current_oracle_equals_reference(cur):
    for name, stream, w, h, img in cur:
        n_ref, out_ref = ol.ref_decode(stream, w, h)
        n_or, out_or = ol.oracle_decode(stream, w, h)
        assert n_ref == w * h, name
        assert n_or == n_ref, name
        assert np.array_equal(out_or, out_ref), name
        if img is not None:
            assert np.array_equal(out_ref, img), f"{name}: encoder is not the inverse of the reference decoder"


@needs_ref
def test_legacy_oracle_equals_reference(leg):
    for name, stream, w, h, img in leg:
        n_ref, out_ref = ol.ref_decode_legacy(stream, w, h)
        n_or, out_or = ol.oracle_decode_legacy(stream, w, h)
        assert n_ref == w * h and n_or == n_ref, name
        assert np.array_equal(out_or, out_ref), name
        if img is not None:
            assert np.array_equal(out_ref, img), name


def test_oracle_round_trips_without_reference(cur, leg):
    for name, stream, w, h, img in cur:
        if img is not None:
            n, out = ol.oracle_decode(stream, w, h)
            assert n == w * h and np.array_equal(out, img), name
    for name, stream, w, h, img in leg:
        if img is not None:
            n, out = ol.oracle_decode_legacy(stream, w, h)
            assert n == w * h and np.array_equal(out, img), name


@needs_ref
def test_known_answers():
    """KATs established against the reference."""
    # residual 0xFFFF + reference 2 -> 1 (uint16 wrap): one 64x4 tile, 16-bit blocks, refs = 2
    bits = np.full(4, 16, dtype=np.uint16)
    refs = np.full(4, 2, dtype=np.uint16)
    from motioncam_decoder_b200 import testvec as tv
    s = tv.assemble_current(64, 4, bits, refs, seed=1)
    s[16:16 + 4 * 128] = 0xFF
    for fn in (ol.ref_decode, ol.oracle_decode):
        n, out = fn(s, 64, 4)
        assert n == 256 and np.all(out == 1)
    # header rejections -> 0
    img = tv.gen_photon(64, 4, 1023, seed=1)
    good = tv.encode_current(img)
    for mutate in ("ew_not_64", "bits_off", "refs_off", "ew_lt_width"):
        s = good.copy()
        if mutate == "ew_not_64":
            s[0:4] = np.frombuffer(np.uint32(96).tobytes(), dtype=np.uint8)
        elif mutate == "bits_off":
            s[8:12] = np.frombuffer(np.uint32(len(s) + 1).tobytes(), dtype=np.uint8)
        elif mutate == "refs_off":
            s[12:16] = np.frombuffer(np.uint32(len(s) + 1).tobytes(), dtype=np.uint8)
        w = 128 if mutate == "ew_lt_width" else 64
        assert ol.ref_decode(s, w, 4)[0] == 0, mutate
        assert ol.oracle_decode(s, w, 4)[0] == 0, mutate


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_random_geometry_fuzz(seed):
    """Random sizes / value ranges through encoder -> reference and oracle."""
    from motioncam_decoder_b200 import testvec as tv
    rng = np.random.default_rng(seed)
    w = int(rng.integers(1, 700))
    h4 = int(rng.integers(1, 6)) * 4
    mx = int(rng.choice([1, 15, 255, 1023, 4095, 65535]))
    img = tv.gen_uniform(w, h4, 0, mx, seed=seed) if seed % 2 else tv.gen_photon(w, h4, max(mx, 64), seed=seed)
    s = tv.encode_current(img, policy=tv.POLICY_ALIASES, ref_wrap=bool(seed % 3 == 0), seed=seed)
    assert np.array_equal(ol.ref_decode(s, w, h4)[1], img)
    assert np.array_equal(ol.oracle_decode(s, w, h4)[1], img)
    h = int(rng.integers(1, 9))
    img = tv.gen_photon(w, h, max(mx, 64), seed=seed + 100)
    s = tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=seed)
    assert np.array_equal(ol.ref_decode_legacy(s, w, h)[1], img)
    assert np.array_equal(ol.oracle_decode_legacy(s, w, h)[1], img)
