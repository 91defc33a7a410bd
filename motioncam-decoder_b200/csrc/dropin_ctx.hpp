// dropin_ctx.hpp -- per-thread device context shared by the drop-in entry points (internal).
#pragma once
#include "mcraw_b200.h"

namespace motioncam {
namespace detail {

// The calling thread's decoder context on device $MCRAW_B200_DEVICE (default 0), created on first use and
// destroyed when the thread exits.  nullptr when no usable device exists (reason printed once on stderr).
mcraw_ctx* threadContext();

}  // namespace detail
}  // namespace motioncam
