// motioncam/RawData.hpp -- the frame codec entry points, same signatures as the reference's
// lib/include/motioncam/RawData.hpp:25-37 (mangled _ZN9motioncam3raw6DecodeEPtiiPKhm and
// _ZN9motioncam3raw12DecodeLegacyEPtiiPKhm), implemented on the B200 through include/mcraw_b200.h.
//
// Contract (reference: lib/RawData.cpp:528-612, lib/RawData_Legacy.cpp:445-495; caller lib/Decoder.cpp:221-230):
//   output  host buffer of width*height uint16, owned by the caller
//   input   the compressed frame buffer (host), len bytes
//   return  number of uint16 ELEMENTS written; 0 = failure.  Never throws.  Re-entrant: every host thread
//           gets its own device context (MCRAW_B200_DEVICE selects the GPU, default 0).
// There is no CPU decode path: without a usable sm_100 device these functions report the reason on stderr and
// return 0, which motioncam::Decoder::loadFrame turns into IOException like any other decode failure.
#pragma once
#include <stddef.h>
#include <cstdint>

namespace motioncam {
namespace raw {

size_t Decode(uint16_t* output, const int width, const int height, const uint8_t* input, const size_t len);
size_t DecodeLegacy(uint16_t* output, const int width, const int height, const uint8_t* input, const size_t len);

}  // namespace raw
}  // namespace motioncam
