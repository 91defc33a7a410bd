"""motioncam-decoder_b200 -- B200-native MCRAW frame decoder (hot path of mirsadm/motioncam-decoder).

Layout
  csrc/        CUDA kernels (sm_100a), the C-ABI (include/mcraw_b200.h), the drop-in C++ API
               (motioncam::raw::Decode/DecodeLegacy, motioncam::Decoder) and CPU test-vector tools
  capi.py      ctypes binding of the C-ABI (what tests and bench.py call)
  testvec.py   ctypes binding of the CPU test-vector encoder / generators, .mcraw writer
  hostapi.py   Python view of the drop-in motioncam::Decoder (through flat C wrappers)

The product path is CUDA only: importing capi without the built extension raises, there is no CPU fallback.
"""
__version__ = "0.1.0"
