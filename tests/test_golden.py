"""Golden fixtures (outputs of the unmodified reference, committed): the oracle on the CPU, the CUDA path on
the GPU.  This is the pin that still holds where /root/reference does not exist."""
import numpy as np
import pytest

import golden_lib
import oracle_lib as ol


def test_golden_fixture_is_complete():
    g = golden_lib.load()
    names = [v[0] for v in g]
    assert len(g) == 41
    assert sum(n.startswith("cur_hdr") for n in names) == 17 and sum(n.startswith("leg_nib") for n in names) == 16


def test_oracle_matches_golden():
    for name, s, w, h, ct, want in golden_lib.load():
        n, got = (ol.oracle_decode if ct == 7 else ol.oracle_decode_legacy)(s, w, h)
        assert n == w * h, name
        assert np.array_equal(got, want), name


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built")
def test_reference_still_matches_golden():
    for name, s, w, h, ct, want in golden_lib.load():
        n, got = (ol.ref_decode if ct == 7 else ol.ref_decode_legacy)(s, w, h)
        assert n == w * h and np.array_equal(got, want), name


@pytest.mark.gpu
def test_cuda_matches_golden():
    from motioncam_decoder_b200 import capi
    g = golden_lib.load()
    ctx = capi.Context(0)
    batch = capi.DeviceBatch(ctx, [(s, w, h, ct) for (_, s, w, h, ct, _) in g])
    batch.fill_outputs(0xA5A5)
    written, status = batch.decode()
    for i, (name, s, w, h, ct, want) in enumerate(g):
        assert status[i] == 0 and written[i] == w * h, name
        assert np.array_equal(batch.fetch(i), want), name
    batch.free()
    ctx.close()
