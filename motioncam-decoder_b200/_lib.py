"""Locations and loaders of the in-tree shared libraries (built by csrc/Makefile)."""
import ctypes
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")

# kernels + C-ABI (needs a GPU to run); MCRAW_B200_LIB names an A/B build of the same sources (csrc/Makefile `variants`)
LIB_CAPI = os.path.join(PKG_DIR, os.environ.get("MCRAW_B200_LIB", "libmcraw_b200.so"))
LIB_TOOLS = os.path.join(PKG_DIR, "libmcraw_tools.so")         # CPU encoder / generators
# drop-in C++ API + flat wrappers; MCRAW_DROPIN_LIB: another build of it (tools/asan_host_fuzz.sh: AddressSanitizer + UBSan)
LIB_DROPIN = os.path.join(PKG_DIR, os.environ.get("MCRAW_DROPIN_LIB", "libmotioncam_decoder_b200.so"))


def load(path):
    if not os.path.exists(path):
        raise ImportError(
            f"{os.path.basename(path)} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {CSRC}` first (there is no CPU fallback for the decode path)")
    return ctypes.CDLL(path)   # RTLD_LOCAL: the reference checker exports the same C++ symbols
