// RawData.cpp -- motioncam::raw::Decode / DecodeLegacy with the reference's signatures
// (/root/reference/lib/include/motioncam/RawData.hpp:25-37), executed by the sm_100a kernels behind
// mcraw_decode_host (include/mcraw_b200.h).  Host buffers in, host buffers out, synchronous -- the contract
// of /root/reference/lib/RawData.cpp:528-612 and lib/RawData_Legacy.cpp:445-495 as seen from
// lib/Decoder.cpp:221-230.  No decode arithmetic happens on the host.
#include <motioncam/RawData.hpp>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "dropin_ctx.hpp"

namespace motioncam {
namespace detail {

namespace {
struct ThreadContext {
    mcraw_ctx* ctx = nullptr;
    bool tried = false;
    ~ThreadContext() {
        if (ctx) mcraw_ctx_destroy(ctx);
    }
};
std::once_flag g_reported;
}  // namespace

mcraw_ctx* threadContext() {
    thread_local ThreadContext t;
    if (!t.tried) {
        t.tried = true;
        int device = 0;
        if (const char* e = std::getenv("MCRAW_B200_DEVICE")) device = std::atoi(e);
        if (mcraw_ctx_create(device, &t.ctx) != MCRAW_OK) {
            t.ctx = nullptr;
            std::call_once(g_reported, [] {
                std::fprintf(stderr, "motioncam-decoder_b200: cannot decode, %s\n", mcraw_last_error(nullptr));
            });
        }
    }
    return t.ctx;
}

}  // namespace detail

namespace raw {

size_t Decode(uint16_t* output, const int width, const int height, const uint8_t* input, const size_t len) {
    mcraw_ctx* ctx = detail::threadContext();
    if (!ctx) return 0;
    return mcraw_decode_host(ctx, output, width, height, input, len, MCRAW_COMPRESSION_CURRENT);
}

size_t DecodeLegacy(uint16_t* output, const int width, const int height, const uint8_t* input, const size_t len) {
    mcraw_ctx* ctx = detail::threadContext();
    if (!ctx) return 0;
    return mcraw_decode_host(ctx, output, width, height, input, len, MCRAW_COMPRESSION_LEGACY);
}

}  // namespace raw
}  // namespace motioncam
