"""GPU: malformed input.  Mutated and random frame buffers must never fault the GPU; a frame either fails
(0 elements written, like RawData.cpp:547-554) or decodes to exactly what the oracle produces."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _mutants(good, rng, count):
    out = []
    for k in range(count):
        s = good.copy()
        kind = k % 6
        if kind == 0:      # flip a few random bytes anywhere
            for _ in range(int(rng.integers(1, 6))):
                s[int(rng.integers(0, len(s)))] = rng.integers(0, 256)
        elif kind == 1:    # truncate
            s = s[: int(rng.integers(1, len(s)))].copy()
        elif kind == 2:    # damage the tail (metadata streams / trailer)
            n = int(rng.integers(1, 64))
            s[-n:] = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 3:    # damage the head (frame header / first blocks)
            n = int(rng.integers(1, 24))
            s[:n] = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 4:    # pure noise of the same size
            s = rng.integers(0, 256, len(s), dtype=np.uint8)
        else:              # append garbage
            s = np.concatenate([s, rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8)])
        out.append(s)
    return out


def _run(frames, ctype, oracle_fn, w, h):
    from motioncam_decoder_b200 import capi
    ctx = capi.Context(0)
    batch = capi.DeviceBatch(ctx, [(s, w, h, ctype) for s in frames])
    batch.fill_outputs(0xA5A5)
    written, status = batch.decode()          # raises on any CUDA error
    ok = bad = 0
    for i, s in enumerate(frames):
        n, want = oracle_fn(s, w, h)
        assert written[i] == n, (i, written[i], n, status[i])
        if n:
            assert status[i] == 0
            assert np.array_equal(batch.fetch(i), want), i
            ok += 1
        else:
            assert status[i] != 0
            bad += 1
    batch.free()
    # the context is still healthy afterwards
    n, img = ctx.decode_host(frames[-1] if False else _GOOD[ctype][0], w, h, ctype)
    assert n == w * h and np.array_equal(img, _GOOD[ctype][1])
    ctx.close()
    return ok, bad


_GOOD = {}


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_current(seed):
    from motioncam_decoder_b200 import capi, testvec as tv
    rng = np.random.default_rng(seed)
    w, h = 320, 16
    img = tv.gen_photon(w, h, 1023, seed=seed)
    good = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=seed)
    _GOOD[capi.COMPRESSION_CURRENT] = (good, img)
    frames = [good] + _mutants(good, rng, 150)
    ok, bad = _run(frames, capi.COMPRESSION_CURRENT, ol.oracle_decode, w, h)
    assert ok >= 1 and bad >= 20


@pytest.mark.parametrize("seed", [3, 4])
def test_fuzz_legacy(seed):
    from motioncam_decoder_b200 import capi, testvec as tv
    rng = np.random.default_rng(seed)
    w, h = 320, 16
    img = tv.gen_photon(w, h, 1023, seed=seed)
    good = tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=seed)
    _GOOD[capi.COMPRESSION_LEGACY] = (good, img)
    frames = [good] + _mutants(good, rng, 150)
    ok, bad = _run(frames, capi.COMPRESSION_LEGACY, ol.oracle_decode_legacy, w, h)
    assert ok >= 1 and bad >= 10
