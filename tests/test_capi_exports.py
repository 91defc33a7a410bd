"""CPU: the C-ABI library loads without a GPU and exports every function include/mcraw_b200.h declares; creating
a context without a device fails loudly (there is no CPU decode path)."""
import ctypes
import os
import re

import pytest

from motioncam_decoder_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_CAPI), reason="libmcraw_b200.so not built (needs nvcc)")


def _declared():
    text = open(os.path.join(ROOT, "include", "mcraw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcraw_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 19
    c = ctypes.CDLL(_lib.LIB_CAPI)
    missing = [n for n in names if not hasattr(c, n)]
    assert not missing, missing
    from motioncam_decoder_b200 import capi
    assert sorted(capi.EXPORTED) == names, "capi.EXPORTED is out of sync with the header"


def test_frame_desc_layout_matches_header():
    from motioncam_decoder_b200 import capi
    assert ctypes.sizeof(capi.FrameDesc) == 48
    assert capi.FrameDesc.dst.offset == 32 and capi.FrameDesc.len.offset == 8


def test_no_device_no_decode():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from motioncam_decoder_b200 import capi
    with pytest.raises(capi.McrawError, match="no CPU decode path|no usable CUDA device"):
        capi.Context(0)
