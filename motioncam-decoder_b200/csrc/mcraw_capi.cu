// mcraw_capi.cu -- implementation of the C-ABI in include/mcraw_b200.h on top of the kernels in
// mcraw_kernels.cuh.  Host side only does bookkeeping: descriptor validation, scratch sizing, launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mcraw_b200.h"
#include "mcraw_kernels.cuh"
#include "mcraw_legacy.cuh"

using namespace mcraw;

// ---------------------------------------------------------------------------------------------------------
// k_checksum: position-weighted 64-bit checksum of decoded frames (mcraw_checksum_frames), so that a caller -- bench.py
// after every timed loop, a pipeline that wants an integrity tag -- can check EVERY pixel of a batch without moving it:
//   sum over i of (v[i] + 1) * ((i + 1) * K)  mod 2^64,  K = 0x9E3779B97F4A7C15 (odd: any single changed sample changes the sum)
// grid = (blocks per frame, frames)
// ---------------------------------------------------------------------------------------------------------
namespace mcraw {
constexpr unsigned long long CK_K = 0x9E3779B97F4A7C15ull;
constexpr int CK_THREADS = 256;
__global__ void __launch_bounds__(CK_THREADS) k_checksum(const uint16_t* const* __restrict__ frames, const unsigned long long* __restrict__ elems,
                                                         unsigned long long* __restrict__ out) {
    const uint16_t* __restrict__ p = frames[blockIdx.y];
    const unsigned long long n = elems[blockIdx.y];
    unsigned long long acc = 0;
    const unsigned long long tid = (unsigned long long)blockIdx.x * CK_THREADS + threadIdx.x, nthr = (unsigned long long)gridDim.x * CK_THREADS;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const unsigned long long n8 = n / 8;
        for (unsigned long long g = tid; g < n8; g += nthr) {
            const uint4 q = __ldcs(reinterpret_cast<const uint4*>(p) + g);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
            unsigned long long m = (8 * g + 1) * CK_K;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc += (unsigned long long)((w[k] & 0xFFFFu) + 1u) * m; m += CK_K;
                acc += (unsigned long long)((w[k] >> 16) + 1u) * m; m += CK_K;
            }
        }
        for (unsigned long long i = 8 * n8 + tid; i < n; i += nthr) acc += (unsigned long long)(p[i] + 1u) * ((i + 1) * CK_K);
    } else {
        for (unsigned long long i = tid; i < n; i += nthr) acc += (unsigned long long)(p[i] + 1u) * ((i + 1) * CK_K);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    __shared__ unsigned long long part[CK_THREADS / 32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < CK_THREADS / 32; k++) t += part[k];
        atomicAdd(out + blockIdx.y, t);
    }
}
}  // namespace mcraw

namespace {

thread_local std::string g_create_error;

constexpr int kSlots = 6;              // chunks (sub-batches) that may be in flight per context (each owns its scratch)
constexpr uint32_t kMaxGridY = 32768;  // frames per launch
constexpr int kStage = 3;              // device staging buffers of the host-input pipeline
constexpr size_t kStageBytes = 96u << 20;   // per-chunk overhead (cross-stream events) favours large chunks: 48 MB -> 51.6 GB/s, 96 MB -> 52.7 GB/s
constexpr int kCopyStreams = 2;
constexpr int kHostParts = 16;         // at most that many pieces (and threads) of the host-side copies of the single-frame call

struct Slot {
    // one pinned / device buffer pair per chunk: [FrameDev x n | WorkItem x items]
    uint8_t* h_up = nullptr;        // pinned
    uint8_t* d_up = nullptr;
    size_t up_bytes = 0;
    // device-written words: [queue counter (16 B) | FrameState x n]; nothing here needs zeroing by the host
    uint8_t* d_dyn = nullptr;
    Result* h_results = nullptr;    // pinned; written by the kernels directly (zero-copy), read by harvest()
    uint32_t cap_frames = 0;
    uint8_t* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t done = nullptr, e0 = nullptr, e1 = nullptr, e2 = nullptr;
    bool in_flight = false;
    bool timed = false;
    uint32_t n = 0;
    uint32_t result_offset = 0;     // index of this chunk's first frame inside the logical batch
    uint64_t batch_id = 0;
    // the plan this slot's device tables were built for (see enqueue_chunk)
    std::vector<mcraw_frame_desc> plan_descs;
    std::vector<mcraw_levels> plan_levels;   // empty: raw output
    bool plan_valid = false, any7 = false, any6 = false, any_epi = false;
    uint32_t max_ltiles = 0, nitems = 0;
    uint32_t split_nw = 0;          // > 0: this plan's metadata chains are resolved by k_meta_split with that many windows per stream
    size_t items_off = 0, plan_bytes = 0;
    uint32_t flag_uses = 0;         // index-kernel launches of this SLOT (whatever the plan): the epoch the per-frame done words carry
    uint32_t plan_epoch = 0;        // plan uploads of this slot: the value of the ready word behind the plan (plan_wait in the kernels)
    size_t ready_off = 0;           // offset of that word in h_up / d_up
    cudaStream_t stream = nullptr;  // the stream the chunk was enqueued on
    uintptr_t dst_lo = 0, dst_hi = 0;   // address range spanned by the plan's output buffers
    uint32_t lg_nwork = 0;          // (frame, tile) tickets of the legacy frames of the plan
    size_t lg_work_off = 0;
    uint32_t lg_epoch = 0;          // k_legacy_warp launches on this plan since its status words were zeroed
    size_t plan_scratch = 0;        // scratch bytes the plan uses
};

struct Stage {
    uint8_t* d_buf = nullptr;
    cudaEvent_t copied = nullptr;   // H2D of the chunk finished (recorded on a copy stream)
    cudaEvent_t freed = nullptr;    // decode that read the buffer finished (recorded on the decode stream)
    bool used = false;
};

}  // namespace

struct mcraw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_streams[kCopyStreams] = {nullptr, nullptr};
    cudaStream_t d2h_stream = nullptr;      // mcraw_decode_batch_host_out: device -> host copies of decoded chunks
    cudaStream_t plan_stream = nullptr;     // uploads of new plans (descriptors + work lists): beside the decode stream, not on it
    cudaEvent_t d2h_done = nullptr;
    bool d2h_pending = false;
    // CHAIN: back-to-back batches on one stream are linked by programmatic dependent launches all the way -- k_meta of
    // batch i+1 is a programmatic dependent of k_units of batch i, so it resolves the next batch's metadata in the room
    // k_units leaves (chain_ctas of its resident CTAs held back) while batch i still streams pixels; the kernels only meet
    // through the per-frame done words (epochs).  Stream order is kept for everything the caller can observe: the link is
    // only made when batch i+1 presents the descriptors of batch i again or writes a disjoint range of output addresses,
    // and anything else enqueued on the stream in between ends the overlap by itself (a programmatic edge only relaxes
    // kernel -> kernel).  MCRAW_CHAIN=<CTAs> overrides the hold-back (0 = no chaining); 24 measured best on B200.
    uint32_t chain_ctas = 24;
    int prev_slot = -1;             // slot of the previous enqueue_chunk call
    Slot slots[kSlots];
    int cur = -1;
    Stage stages[kStage];
    int stage_cur = 0;
    int copy_rr = 0;
    uint64_t launches = 0;
    std::vector<FrameDev> tmp_frames;
    std::vector<WorkItem> tmp_items;
    std::vector<LgWork> tmp_lgwork;
    uint32_t sm_count = 0;
    bool meta_small_only = getenv("MCRAW_META_SMALL") != nullptr;   // A/B switch: never use the big-window shape of k_meta
    uint32_t split_resident_ctas = 0;   // CTAs of k_meta_split the device holds at once (all of a launch must be resident)
    uint32_t meta_resident_ctas = 0;    // CTAs of k_meta<K1Batch> the device holds at once (one wave)
    bool plan_side = !(getenv("MCRAW_PLAN_SIDE") && atoi(getenv("MCRAW_PLAN_SIDE")) == 0);   // A/B switch: 0 = new plans are uploaded on the decode stream
    int host_parts = std::max(1, std::min(kHostParts, getenv("MCRAW_HOST_PARTS") ? atoi(getenv("MCRAW_HOST_PARTS")) : 4));   // mcraw_decode_host
    size_t hostout_first_bytes = getenv("MCRAW_HOSTOUT_FIRST_MB") ? (size_t)std::max(0, atoi(getenv("MCRAW_HOSTOUT_FIRST_MB"))) << 20
                                                                   : (size_t)32 << 20;   // first chunk of a host-out batch (measured on C2: 0 / 8 / 16 / 32 / 48 MB -> 23.8 / 23.3 / 23.2 / 21.8 / 22.3 ms per 240 frames)
    bool meta_split = !(getenv("MCRAW_META_SPLIT") && atoi(getenv("MCRAW_META_SPLIT")) == 0) &&
                      !(getenv("MCRAW_META_WARP") && atoi(getenv("MCRAW_META_WARP")) == 2);     // A/B switch
    // batches go to k_meta_warp; A/B and test switch MCRAW_META_WARP: 0 = k_meta<K1Batch> instead, 2 = k_meta_warp for EVERY launch
    // (also where a handful of frames would pick the big-window or the split kernel)
    int meta_warp = getenv("MCRAW_META_WARP") ? atoi(getenv("MCRAW_META_WARP")) : 1;
    uint32_t lgw_resident_ctas = 0; // CTAs of k_legacy_warp<false> the device holds at once
    uint32_t lgw_resident_ctas_epi = 0;   // the same for the variant with the epilogue (more registers)
    bool overlap = getenv("MCRAW_NO_OVERLAP") == nullptr;   // k_units as a programmatic dependent of k_meta
    uint32_t timing_every = 0;      // record kernel-timing events on every n-th chunk (0 = never)
    uint64_t chunk_seq = 0;
    uint32_t resident_ctas = 0;     // CTAs of k_units the device holds at once
    uint64_t batch_id = 0;
    uint32_t batch_n = 0;
    std::vector<uint64_t> res_written;
    std::vector<uint32_t> res_status;
    std::vector<int> res_type;
    float last_kernel_ms = 0.f;
    double acc_meta_ms = 0, acc_main_ms = 0;
    uint64_t acc_chunks = 0;
    std::string err;
    // staging for the single-frame host call
    uint8_t* h_in = nullptr; size_t h_in_cap = 0;
    uint8_t* d_in = nullptr; size_t d_in_cap = 0;
    uint16_t* h_out = nullptr; size_t h_out_cap = 0;
    uint16_t* d_out = nullptr; size_t d_out_cap = 0;
    cudaEvent_t part_done[kHostParts] = {};
};

namespace {

#define CU_TRY(ctx, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            (void)cudaGetLastError(); /* reported here: must not resurface at the next launch check */ \
            return MCRAW_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

int fail_arg(mcraw_ctx* ctx, const std::string& msg) {
    ctx->err = msg;
    return MCRAW_ERR_ARG;
}

int bind(mcraw_ctx* ctx) {
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return MCRAW_OK;
}

int slot_reserve(mcraw_ctx* ctx, Slot& s, uint32_t n, size_t up_bytes, size_t scratch) {
    if (n > s.cap_frames) {
        uint32_t cap = std::max<uint32_t>(n, s.cap_frames * 2 + 64);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.d_dyn) cudaFree(s.d_dyn);
        s.h_results = nullptr; s.d_dyn = nullptr; s.cap_frames = 0;
        CU_TRY(ctx, cudaMallocHost(&s.h_results, sizeof(Result) * cap));
        CU_TRY(ctx, cudaMalloc(&s.d_dyn, 16 + sizeof(FrameState) * cap));
        // queue counters (the kernels leave them at zero) and done words (epochs: zero is older than any launch)
        CU_TRY(ctx, cudaMemset(s.d_dyn, 0, 16 + sizeof(FrameState) * cap));
        s.cap_frames = cap;
    }
    if (up_bytes > s.up_bytes) {
        size_t cap = std::max(up_bytes, s.up_bytes + s.up_bytes / 2);
        if (s.h_up) cudaFreeHost(s.h_up);
        if (s.d_up) cudaFree(s.d_up);
        s.h_up = nullptr; s.d_up = nullptr; s.up_bytes = 0;
        CU_TRY(ctx, cudaMallocHost(&s.h_up, cap));
        CU_TRY(ctx, cudaMalloc(&s.d_up, cap));
        s.up_bytes = cap;
    }
    if (scratch > s.scratch_bytes) {
        size_t cap = std::max(scratch, s.scratch_bytes + s.scratch_bytes / 2);
        if (s.d_scratch) cudaFree(s.d_scratch);
        s.d_scratch = nullptr; s.scratch_bytes = 0;
        CU_TRY(ctx, cudaMalloc(&s.d_scratch, cap));
        s.scratch_bytes = cap;
    }
    return MCRAW_OK;
}

// Wait for a slot's chunk, fold its results into the logical batch (if it still belongs to the current
// one) and its kernel times into the context totals.
int harvest(mcraw_ctx* ctx, Slot& s) {
    if (!s.in_flight) return MCRAW_OK;
    CU_TRY(ctx, cudaEventSynchronize(s.done));
    s.in_flight = false;
    if (s.timed) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, s.e0, s.e1) == cudaSuccess && cudaEventElapsedTime(&b, s.e1, s.e2) == cudaSuccess) {
            ctx->acc_meta_ms += a; ctx->acc_main_ms += b; ctx->acc_chunks += 1;
            ctx->last_kernel_ms = a + b;
        }
    }
    if (s.batch_id == ctx->batch_id) {
        for (uint32_t i = 0; i < s.n; i++) {
            const uint32_t k = s.result_offset + i;
            if (k >= ctx->res_written.size()) break;
            ctx->res_written[k] = s.h_results[i].written;
            ctx->res_status[k] = s.h_results[i].status;
        }
    }
    return MCRAW_OK;
}

// Validate descriptors and build the device-side frame records; tilemeta holds an OFFSET until rebased.
int prepare(mcraw_ctx* ctx, const mcraw_frame_desc* descs, const mcraw_levels* levels, uint32_t n, uint32_t first_index, std::vector<FrameDev>& out,
            size_t& scratch, uint32_t& max_tile_rows, uint32_t& max_units, uint32_t& max_ltiles, bool& any7, bool& any6, uint32_t& split_nw) {
    out.resize(n);
    scratch = 0; max_tile_rows = 0; max_units = 0; max_ltiles = 0; any7 = any6 = false; split_nw = 0;
    for (uint32_t i = 0; i < n; i++) {
        const mcraw_frame_desc& d = descs[i];
        auto who = [&] { return "frame " + std::to_string(first_index + i); };   // only built on the error paths
        FrameDev f;
        std::memset(&f, 0, sizeof f);
        if (!d.src || !d.dst) return fail_arg(ctx, who() + ": null src/dst");
        if (d.width <= 0 || d.height <= 0 || d.width > 65536 || d.height > 65536 ||
            (uint64_t)d.width * (uint64_t)d.height > (1ull << 30))        // payload offsets are 32 bits wide on the device
            return fail_arg(ctx, who() + ": unsupported width/height");
        if (((uintptr_t)d.src & 15) || ((uintptr_t)d.dst & 1))
            return fail_arg(ctx, who() + ": src must be 16-byte aligned, dst 2-byte aligned");
        f.src = d.src; f.len = d.len; f.dst = d.dst; f.dst_cap = d.dst_capacity_elems;
        f.width = d.width; f.height = d.height; f.type = d.compression_type;
        f.tiles_x = (uint32_t)(d.width + 63) / 64;
        f.tile_rows = (uint32_t)(d.height + 3) / 4;
        if (d.compression_type == MCRAW_COMPRESSION_CURRENT && d.encoded_width != 0) {
            // the caller has seen the frame header: plan for its encodedWidth (RawData.cpp:550-554 accepts any multiple of 64 >= width)
            if (d.encoded_width < d.width || (d.encoded_width % 64) != 0 || d.encoded_width > (1 << 24))
                return fail_arg(ctx, who() + ": encoded_width must be a multiple of 64, >= width");
            f.tiles_x = (uint32_t)d.encoded_width / 64;
        }
        f.flags = ((d.width % 8) == 0 && ((uintptr_t)d.dst & 15) == 0) ? FLAG_VEC_STORE : 0;
        if (levels && levels[i].mode != MCRAW_OUT_RAW) {
            const mcraw_levels& L = levels[i];
            if (L.mode != MCRAW_OUT_BLACK_SUB && L.mode != MCRAW_OUT_NORM_F16) return fail_arg(ctx, who() + ": unknown output mode");
            f.epi_mode = L.mode;
            auto u16 = [](float v) { return (uint32_t)std::min(65535l, std::max(0l, lrintf(v))); };
            const uint32_t w = u16(L.white);
            for (int c = 0; c < 4; c++) {
                if (!std::isfinite(L.black[c]) || !std::isfinite(L.white)) return fail_arg(ctx, who() + ": black / white level is not a number");
                const uint32_t b = u16(L.black[c]);
                f.epi_black2[c >> 1] |= b << (16 * (c & 1));
                f.epi_range2[c >> 1] |= (w > b ? w - b : 0u) << (16 * (c & 1));
                f.epi_blackf[c] = L.black[c];
                f.epi_scalef[c] = L.white > L.black[c] ? 1.0f / (L.white - L.black[c]) : 0.0f;
            }
        }
        if (d.compression_type == MCRAW_COMPRESSION_CURRENT) {
            any7 = true;
            const uint64_t ntiles = (uint64_t)f.tiles_x * f.tile_rows;
            f.nunits = (uint32_t)((ntiles + 15) / 16);
            f.inv_tiles_x = (ntiles * f.tiles_x < (1ull << 32)) ? (uint32_t)(((1ull << 32) + f.tiles_x - 1) / f.tiles_x) : 0u;
            if (f.tiles_x == 1) f.inv_tiles_x = 0;   // 2^32 does not fit; plain division
            // scratch layout (offsets for now, rebased on the slot's buffer): unitoff | metarec
            // (128-byte aligned: a frame's scratch never shares a cache line with another frame's)
            f.unitoff = reinterpret_cast<uint32_t*>(scratch);
            scratch += (((size_t)f.nunits + 1) * 4 + 127) & ~(size_t)127;
            f.metarec = reinterpret_cast<uint4*>(scratch);
            scratch += ((size_t)f.nunits * 16 + 127) & ~(size_t)127;
            max_tile_rows = std::max(max_tile_rows, f.tile_rows);
            max_units = std::max(max_units, f.nunits);
        } else if (d.compression_type == MCRAW_COMPRESSION_LEGACY) {
            any6 = true;
            if (d.len >= ((uint64_t)1 << 40)) return fail_arg(ctx, who() + ": legacy frame buffer too large");
            const uint64_t ntile = lgw_ntiles(d.len);
            if (ntile > 0x7FFFFFFFull) return fail_arg(ctx, who() + ": legacy frame buffer too large");
            // scratch layout (offsets for now): exit maps (written only by tiles whose 17 exits differ) | two look-back
            // status words per tile (count word, exit word)
            f.lg_tilemap = reinterpret_cast<uint32_t*>(scratch);
            scratch += ((size_t)ntile * LG_STATES * 4 + 15) & ~(size_t)15;
            f.lg_status = reinterpret_cast<unsigned long long*>(scratch);
            scratch += (size_t)ntile * 16;
            scratch = (scratch + 127) & ~(size_t)127;
            max_ltiles = std::max<uint32_t>(max_ltiles, (uint32_t)ntile);
        }   // any other type: no work is queued, mcraw_batch_wait reports MCRAW_FRAME_BAD_TYPE
        out[i] = f;
    }
    // A handful of big current-format frames: one stream's chain is the critical path, so it is cut into windows worked by
    // several CTAs at once (k_meta_split).  A stream of u units is at most 4 + 130 u bytes long and starts anywhere in the
    // buffer; every CTA of the launch has to be resident (they wait for each other), hence the bound.
    if (any7 && ctx->meta_split) {
        uint64_t nw = 0;
        for (uint32_t i = 0; i < n; i++)
            if (out[i].type == MCRAW_COMPRESSION_CURRENT) {
                const uint64_t extent = std::min<uint64_t>(32 + 130ull * out[i].nunits, out[i].len + 16);
                nw = std::max<uint64_t>(nw, (extent + KS::C - 1) / KS::C + 1);
            }
        if (nw >= 3 && nw <= (uint64_t)KS::MAXW && 2ull * n * nw <= ctx->split_resident_ctas) {
            split_nw = (uint32_t)nw;
            for (uint32_t i = 0; i < n; i++)
                if (out[i].type == MCRAW_COMPRESSION_CURRENT) {
                    out[i].sp_scratch = reinterpret_cast<void*>(scratch);
                    scratch += (2 * ks_stream_scratch(split_nw) + 127) & ~(size_t)127;
                }
        }
    }
    return MCRAW_OK;
}

// Host-source entry points: the frame header is in reach, so the descriptor's encoded_width is filled in from it (when
// the caller left it 0 and the header value is one a valid frame can carry; anything else is left to the kernels' checks).
void peek_encoded_width(mcraw_frame_desc& d) {
    if (d.compression_type != MCRAW_COMPRESSION_CURRENT || d.encoded_width != 0 || !d.src || d.len < 16 || d.width <= 0 || d.height <= 0) return;
    uint32_t ew;
    std::memcpy(&ew, d.src, 4);                                   // little-endian u32 (RawData.cpp:500-524); hosts here are little-endian
    if (ew % 64u || ew < (uint32_t)d.width || ew > (1u << 24)) return;
    // every unit (64 blocks) costs at least two 2-byte metadata block headers: a frame of len bytes cannot hold more
    const uint64_t units = ((uint64_t)(ew / 64u) * (uint64_t)((d.height + 3) / 4) + 15) / 16;
    if (units * 4 > d.len) return;
    d.encoded_width = (int32_t)ew;
}

// Start a new logical batch of n frames: earlier batches' unharvested chunks only keep their timing.
int begin_batch(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n) {
    ctx->batch_id++;
    ctx->batch_n = n;
    ctx->res_written.assign(n, 0);
    ctx->res_status.assign(n, 0);
    ctx->res_type.assign(n, 0);        // frames that are never appended report MCRAW_FRAME_BAD_TYPE
    if (descs)
        for (uint32_t i = 0; i < n; i++) ctx->res_type[i] = descs[i].compression_type;
    return MCRAW_OK;
}

// Pixel work list of the current-format frames of a chunk (see k_units), in frame order.  Items shrink towards the end
// of the list so that the persistent warps finish together (guided self-scheduling).
void build_items(const std::vector<FrameDev>& frames, uint32_t resident_ctas, std::vector<WorkItem>& items) {
    items.clear();
    uint64_t remaining = 0;
    for (const FrameDev& f : frames)
        if (f.type == MCRAW_COMPRESSION_CURRENT) remaining += f.nunits;
    const uint64_t warps = (uint64_t)std::max<uint32_t>(resident_ctas, 1) * KU_WARPS;
    for (uint32_t i = 0; i < frames.size(); i++) {
        if (frames[i].type != MCRAW_COMPRESSION_CURRENT) continue;
        const uint32_t nunits = frames[i].nunits;
        for (uint32_t u = 0; u < nunits;) {
            // an item never holds more than about half of an even share of what is left
            const uint32_t take = std::min<uint32_t>(nunits - u, (uint32_t)std::min<uint64_t>(KU_UPW, std::max<uint64_t>(1, remaining / (2 * warps))));
            items.push_back(WorkItem{i, u | ((take - 1) << 27)});
            u += take;
            remaining -= take;
        }
    }
}

// Enqueue one chunk (device-resident sources) of the current logical batch on `st`.  Everything stays on that one
// stream: on this platform a cross-stream event dependency costs tens of microseconds, more than the index kernels it
// could hide (measured: profiles/README.md).
int enqueue_chunk(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, uint32_t result_offset, cudaStream_t st,
                  const mcraw_levels* levels = nullptr) {
    if (n == 0) return MCRAW_OK;
    // Slots are taken in turn: a caller that presents the same batch again and again has every slot built for it after one
    // round, and consecutive batches always sit in different slots (the chain).  (Looking a plan up by content instead --
    // measured -- leaves stale slots unconverted for as long as finished matching ones turn up, and the conversions, which
    // may have to grow a slot's buffers behind a device-wide synchronisation, then land anywhere in a steady-state loop.)
    const int pick = (ctx->cur + 1) % kSlots;
    ctx->cur = pick;
    Slot& s = ctx->slots[ctx->cur];
    int rc = harvest(ctx, s);
    if (rc) return rc;

    // A slot remembers the descriptors it was last built for: a caller that decodes into a ring of buffers presents the
    // same descriptors again and again, and then the device-side tables of the slot are still valid -- nothing to
    // validate, build or upload.
    // (flag_uses bound: the per-frame done words carry the slot's launch count as their epoch, 32 bits wide; a slot that
    // has launched 2^32 - 16 times has its words zeroed and starts over -- as a plan that is uploaded afresh.)
    const bool same_levels = levels ? (s.plan_levels.size() == n && std::memcmp(s.plan_levels.data(), levels, sizeof(mcraw_levels) * n) == 0)
                                    : s.plan_levels.empty();
    const bool hit = s.plan_valid && same_levels && s.flag_uses < 0xFFFFFFF0u && s.lg_epoch < 0xFFFFF0u && s.plan_descs.size() == n &&
                     std::memcmp(s.plan_descs.data(), descs, sizeof(mcraw_frame_desc) * n) == 0;
    if (!hit) {
        s.plan_valid = false;
        std::vector<FrameDev>& frames = ctx->tmp_frames;     // reused across calls: no allocation in steady state
        std::vector<WorkItem>& items = ctx->tmp_items;
        size_t scratch; uint32_t max_tile_rows, max_units;
        rc = prepare(ctx, descs, levels, n, result_offset, frames, scratch, max_tile_rows, max_units, s.max_ltiles, s.any7, s.any6, s.split_nw);
        if (rc) return rc;
        items.clear();
        if (s.any7) build_items(frames, ctx->resident_ctas, items);
        s.items_off = (sizeof(FrameDev) * n + 15) & ~(size_t)15;
        s.nitems = (uint32_t)items.size();
        // legacy tickets: tile index major, frame minor -- neighbouring tickets belong to different frames, so the look-back
        // chain of every frame only has to advance a few tiles per generation of resident CTAs (k_legacy_warp)
        std::vector<LgWork>& lgwork = ctx->tmp_lgwork;
        lgwork.clear();
        if (s.any6) {
            std::vector<std::pair<uint32_t, uint32_t>> lf;      // (frame, tiles)
            for (uint32_t i = 0; i < n; i++)
                if (frames[i].type == MCRAW_COMPRESSION_LEGACY)
                    lf.emplace_back(i, (uint32_t)lgw_ntiles(frames[i].len));
            for (uint32_t t = 0; t < s.max_ltiles; t++)
                for (const auto& fr : lf)
                    if (t < fr.second) lgwork.push_back(LgWork{fr.first, t});
        }
        s.lg_work_off = (s.items_off + sizeof(WorkItem) * items.size() + 15) & ~(size_t)15;
        s.lg_nwork = (uint32_t)lgwork.size();
        s.plan_bytes = (s.lg_work_off + sizeof(LgWork) * lgwork.size() + 15) & ~(size_t)15;
        s.plan_scratch = scratch;
        s.ready_off = s.plan_bytes;                                  // the plan's ready word travels behind it
        rc = slot_reserve(ctx, s, n, s.plan_bytes + 16, scratch);
        if (rc) return rc;
        FrameDev* h_frames = reinterpret_cast<FrameDev*>(s.h_up);
        for (uint32_t i = 0; i < n; i++) {
            FrameDev& f = frames[i];
            if (f.type == MCRAW_COMPRESSION_CURRENT) {
                f.unitoff = reinterpret_cast<uint32_t*>(s.d_scratch + reinterpret_cast<size_t>(f.unitoff));
                f.metarec = reinterpret_cast<uint4*>(s.d_scratch + reinterpret_cast<size_t>(f.metarec));
                if (s.split_nw) f.sp_scratch = s.d_scratch + reinterpret_cast<size_t>(f.sp_scratch);
            } else if (f.type == MCRAW_COMPRESSION_LEGACY) {
                f.lg_tilemap = reinterpret_cast<uint32_t*>(s.d_scratch + reinterpret_cast<size_t>(f.lg_tilemap));
                f.lg_status = reinterpret_cast<unsigned long long*>(s.d_scratch + reinterpret_cast<size_t>(f.lg_status));
            }
            h_frames[i] = f;
        }
        if (!items.empty()) std::memcpy(s.h_up + s.items_off, items.data(), sizeof(WorkItem) * items.size());
        if (!lgwork.empty()) std::memcpy(s.h_up + s.lg_work_off, lgwork.data(), sizeof(LgWork) * lgwork.size());
        s.plan_descs.assign(descs, descs + n);
        if (levels) s.plan_levels.assign(levels, levels + n); else s.plan_levels.clear();
        s.any_epi = false;
        for (uint32_t i = 0; i < n; i++) s.any_epi = s.any_epi || frames[i].epi_mode != 0;
        s.dst_lo = ~(uintptr_t)0; s.dst_hi = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(descs[i].dst);
            s.dst_lo = std::min(s.dst_lo, a);
            s.dst_hi = std::max(s.dst_hi, a + 2 * (uintptr_t)descs[i].dst_capacity_elems);
        }
    }
    uint32_t* d_counter = reinterpret_cast<uint32_t*>(s.d_dyn);
    FrameState* d_states = reinterpret_cast<FrameState*>(s.d_dyn + 16);
    Result* d_results = s.h_results;   // pinned host memory: the kernels write the 16-byte per-frame results straight to the host
    const FrameDev* d_frames = reinterpret_cast<const FrameDev*>(s.d_up);
    const WorkItem* d_items = reinterpret_cast<const WorkItem*>(s.d_up + s.items_off);
    const uint32_t* d_ready = reinterpret_cast<const uint32_t*>(s.d_up + s.ready_off);
    const bool any7 = s.any7, any6 = s.any6;
    s.n = n; s.result_offset = result_offset; s.batch_id = ctx->batch_id;

    // descriptor + work list upload and the index kernels between e0 and e1, the pixel kernels between e1 and e2
    const bool timed = ctx->timing_every && (ctx->chunk_seq++ % ctx->timing_every) == 0;
    // A NEW plan of current-format frames goes up on the plan stream, beside the decode stream: the kernels wait for its
    // ready word themselves (plan_wait), the per-frame done words carry epochs and the queue counters reset themselves, so
    // the decode stream holds nothing but the two kernels -- a batch of new descriptors chains like a repeated one.
    // Plans with legacy frames or k_meta_split scratch (status words / flags zeroed per plan) keep the in-stream upload.
    const bool wrap = s.flag_uses >= 0xFFFFFFF0u;
    const bool side = !hit && any7 && !any6 && !s.split_nw && !wrap && ctx->plan_side;
    bool chain = false;
    if (ctx->overlap && ctx->chain_ctas && (hit || side) && any7 && !any6 && !timed && ctx->prev_slot >= 0 && ctx->prev_slot != ctx->cur) {
        const Slot& p = ctx->slots[ctx->prev_slot];
        // the previous chunk ended with k_units on this stream, and whatever it still writes cannot collide with this chunk
        if (p.plan_valid && p.any7 && !p.any6 && p.stream == st && !p.timed) {
            const bool same = p.plan_descs.size() == n && std::memcmp(p.plan_descs.data(), descs, sizeof(mcraw_frame_desc) * n) == 0;
            chain = same || p.dst_hi <= s.dst_lo || s.dst_hi <= p.dst_lo;
        }
    }
    s.stream = st;
    if (timed) CU_TRY(ctx, cudaEventRecord(s.e0, st));
    if (!hit) {
        s.plan_epoch += 1;
        std::memcpy(s.h_up + s.ready_off, &s.plan_epoch, sizeof s.plan_epoch);
        cudaStream_t up = side ? ctx->plan_stream : st;
        CU_TRY(ctx, cudaMemcpyAsync(s.d_up, s.h_up, s.plan_bytes, cudaMemcpyHostToDevice, up));
        CU_TRY(ctx, cudaMemcpyAsync(s.d_up + s.ready_off, s.h_up + s.ready_off, 4, cudaMemcpyHostToDevice, up));   // behind the plan, in stream order
        if (wrap) {                                        // the epochs start over: every done word of the slot back to zero
            CU_TRY(ctx, cudaMemsetAsync(s.d_dyn, 0, 16 + sizeof(FrameState) * s.cap_frames, st));
            s.flag_uses = 0;
        }
        s.plan_valid = true;
        // k_legacy_warp: the tiles' status words are tagged with the launch epoch of the plan, which starts over here
        if ((any6 || s.split_nw) && s.plan_scratch) CU_TRY(ctx, cudaMemsetAsync(s.d_scratch, 0, s.plan_scratch, st));
        s.lg_epoch = 0;
    }
    if (any7) {
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof cfg);
        // a handful of frames: one stream's chain is the critical path -> big windows, one CTA per SM (k_meta)
        const bool few = 2 * n <= ctx->sm_count && !ctx->meta_small_only && ctx->meta_warp != 2;
        cfg.gridDim = dim3(2 * n); cfg.blockDim = dim3(few ? K1Few::K1_THREADS : K1Batch::K1_THREADS);
        cfg.dynamicSmemBytes = few ? K1Few::K1_SMEM : K1Batch::K1_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = chain ? 1 : 0;
        if (s.split_nw) {
            // the launch epoch of the plan tags the windows' flags (the scratch was zeroed when the plan was uploaded)
            cfg.gridDim = dim3(2 * n * s.split_nw); cfg.blockDim = dim3(KS::THREADS); cfg.dynamicSmemBytes = KS::SMEM;
            CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_meta_split, d_frames, d_states, s.split_nw, s.flag_uses + 1u));
        } else if (few) CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_meta<K1Few>, d_frames, d_states, s.flag_uses + 1u, d_ready, s.plan_epoch));
        else if (ctx->meta_warp == 2 || (ctx->meta_warp == 1 && (chain || 2 * n > 2 * ctx->meta_resident_ctas))) {
            // a batch whose index work hides behind the pixel kernel of the batch before (chained), or one of more than two
            // waves of k_meta CTAs: one warp per (frame, stream) -- a third of k_meta's instructions and a tenth of its SM
            // time, all streams at once; slower per stream, so a lone smaller batch stays with k_meta (k_meta_warp)
            cfg.gridDim = dim3((2 * n + KW::WARPS - 1) / KW::WARPS); cfg.blockDim = dim3(KW::THREADS); cfg.dynamicSmemBytes = KW::SMEM;
            CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_meta_warp, d_frames, d_states, 2u * n, s.flag_uses + 1u, d_ready, s.plan_epoch));
        } else CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_meta<K1Batch>, d_frames, d_states, s.flag_uses + 1u, d_ready, s.plan_epoch));
        ctx->launches += 1;
    }
    if (timed) CU_TRY(ctx, cudaEventRecord(s.e1, st));
    if (any7) {
        const uint32_t want = (s.nitems + KU_WARPS - 1) / KU_WARPS;
        const uint32_t hold = chain ? ctx->chain_ctas : 0u;
        const uint32_t room = ctx->resident_ctas > hold ? ctx->resident_ctas - hold : ctx->resident_ctas;
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(std::max<uint32_t>(room, 1), want));
        // Programmatic dependent launch: k_units becomes resident while k_meta's last wave is still running and synchronises
        // per frame (see k_units).  Timed batches are launched the ordinary way, so that the events bracket one kernel each.
        s.flag_uses += 1;
        const bool pdl = ctx->overlap && !timed;
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(KD_THREADS); cfg.dynamicSmemBytes = KU_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        if (s.any_epi) CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_units<true>, d_frames, (const FrameState*)d_states, d_results, d_items, s.nitems,
                                                      d_counter, pdl ? s.flag_uses : 0u, d_ready, s.plan_epoch));
        else CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_units<false>, d_frames, (const FrameState*)d_states, d_results, d_items, s.nitems,
                                            d_counter, pdl ? s.flag_uses : 0u, d_ready, s.plan_epoch));
        ctx->launches += 1;
    }
    if (any6) {
        // one pass over the stream: transfer maps, decoupled look-back and the pixel work in one persistent kernel
        s.lg_epoch += 1;
        const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>(s.any_epi ? ctx->lgw_resident_ctas_epi : ctx->lgw_resident_ctas, s.lg_nwork));
        const LgWork* d_work = reinterpret_cast<const LgWork*>(s.d_up + s.lg_work_off);
        // Chained like the current-format kernels: a batch of legacy frames that follows a batch of legacy frames on the same
        // stream, with a plan that is already on the device (nothing but the kernel goes onto the stream) and outputs that
        // are the same or disjoint, is a programmatic dependent of the kernel before -- no bubble while that one's last tiles finish.
        bool chain6 = false;
        if (ctx->overlap && ctx->chain_ctas && hit && !any7 && !timed && ctx->prev_slot >= 0 && ctx->prev_slot != ctx->cur) {
            const Slot& p = ctx->slots[ctx->prev_slot];
            if (p.plan_valid && p.any6 && !p.any7 && p.stream == st && !p.timed) {
                const bool same = p.plan_descs.size() == n && std::memcmp(p.plan_descs.data(), descs, sizeof(mcraw_frame_desc) * n) == 0;
                chain6 = same || p.dst_hi <= s.dst_lo || s.dst_hi <= p.dst_lo;
            }
        }
        cudaLaunchConfig_t cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(LGW_THREADS); cfg.dynamicSmemBytes = LGW_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = chain6 ? 1 : 0;
        if (s.any_epi) CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_legacy_warp<true>, d_frames, d_results, d_work, s.lg_nwork, d_counter, s.lg_epoch));
        else CU_TRY(ctx, cudaLaunchKernelEx(&cfg, k_legacy_warp<false>, d_frames, d_results, d_work, s.lg_nwork, d_counter, s.lg_epoch));
        ctx->launches += 1;
    }
    if (timed) CU_TRY(ctx, cudaEventRecord(s.e2, st));
    s.timed = timed && (any7 || any6);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventRecord(s.done, st));
    s.in_flight = true;
    ctx->prev_slot = ctx->cur;
    return MCRAW_OK;
}

int enqueue(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, cudaStream_t st, const mcraw_levels* levels = nullptr) {
    if (!descs && n) return fail_arg(ctx, "descs is null");
    int rc = bind(ctx);
    if (rc) return rc;
    begin_batch(ctx, descs, n);
    for (uint32_t base = 0; base < n; base += kMaxGridY) {
        rc = enqueue_chunk(ctx, descs + base, std::min(kMaxGridY, n - base), base, st, levels ? levels + base : nullptr);
        if (rc) return rc;
    }
    return MCRAW_OK;
}

// Run fn(0) .. fn(parts - 1), one thread each (the caller takes part 0).
template <typename Fn>
void parallel_parts(int parts, Fn fn) {
    std::vector<std::thread> pool;
    for (int k = 1; k < parts; k++) pool.emplace_back(fn, k);
    fn(0);
    for (std::thread& t : pool) t.join();
}

template <typename T>
int grow(mcraw_ctx* ctx, T*& p, size_t& cap, size_t need, bool pinned) {
    if (need <= cap) return MCRAW_OK;
    size_t ncap = std::max(need, cap + cap / 2);
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
    if (pinned) CU_TRY(ctx, cudaMallocHost(&p, ncap));
    else CU_TRY(ctx, cudaMalloc(&p, ncap));
    cap = ncap;
    return MCRAW_OK;
}

}  // namespace

extern "C" {

const char* mcraw_version(void) { return "mcraw_b200 0.1.0 (sm_100a)"; }

int mcraw_ctx_create(int device, mcraw_ctx** out) {
    if (!out) return MCRAW_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU decode path)";
        return MCRAW_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { g_create_error = "device index out of range"; return MCRAW_ERR_ARG; }
    mcraw_ctx* ctx = new mcraw_ctx();
    ctx->device = device;
    auto bail = [&](int rc) { g_create_error = ctx->err; mcraw_ctx_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(MCRAW_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return bail(MCRAW_ERR_CUDA); }
    if (prop.major < 10) {
        ctx->err = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   "; this library only carries sm_100a code";
        return bail(MCRAW_ERR_NO_DEVICE);
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream create failed"; return bail(MCRAW_ERR_CUDA); }
    for (auto& cs : ctx->copy_streams)
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream create failed"; return bail(MCRAW_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&ctx->plan_stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "stream create failed"; return bail(MCRAW_ERR_CUDA); }
    if (const char* e = getenv("MCRAW_CHAIN")) ctx->chain_ctas = (uint32_t)std::max(0, atoi(e));
    if (cudaFuncSetAttribute(k_meta<K1Batch>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1Batch::K1_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_meta<K1Few>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1Few::K1_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_meta_split, cudaFuncAttributeMaxDynamicSharedMemorySize, KS::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_units<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, KU_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_units<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, KU_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_legacy_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LGW_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_legacy_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LGW_SMEM) != cudaSuccess) {
        ctx->err = "cudaFuncSetAttribute(smem) failed"; return bail(MCRAW_ERR_CUDA);
    }
    {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_units<true>, KD_THREADS, KU_SMEM) != cudaSuccess || per_sm < 1) {
            ctx->err = "k_units does not fit on this device"; return bail(MCRAW_ERR_CUDA);
        }
        ctx->resident_ctas = (uint32_t)per_sm * (uint32_t)prop.multiProcessorCount;
        ctx->sm_count = (uint32_t)prop.multiProcessorCount;
        ctx->chain_ctas = std::min(ctx->chain_ctas, ctx->resident_ctas / 2);
        int per_sm_split = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_split, k_meta_split, KS::THREADS, KS::SMEM) != cudaSuccess) per_sm_split = 0;
        ctx->split_resident_ctas = (uint32_t)std::max(0, per_sm_split) * (uint32_t)prop.multiProcessorCount;
        int per_sm_meta = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_meta, k_meta<K1Batch>, K1Batch::K1_THREADS, K1Batch::K1_SMEM) != cudaSuccess) per_sm_meta = 0;
        ctx->meta_resident_ctas = (uint32_t)std::max(1, per_sm_meta) * (uint32_t)prop.multiProcessorCount;
        int per_sm_epi = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_legacy_warp<false>, LGW_THREADS, LGW_SMEM) != cudaSuccess || per_sm < 1 ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_epi, k_legacy_warp<true>, LGW_THREADS, LGW_SMEM) != cudaSuccess || per_sm_epi < 1) {
            ctx->err = "k_legacy_warp does not fit on this device"; return bail(MCRAW_ERR_CUDA);
        }
        if (const char* e = getenv("MCRAW_LGW_CTAS_PER_SM")) { per_sm = std::max(1, std::min(per_sm, atoi(e))); per_sm_epi = std::min(per_sm_epi, per_sm); }
        ctx->lgw_resident_ctas = (uint32_t)per_sm * (uint32_t)prop.multiProcessorCount;
        ctx->lgw_resident_ctas_epi = (uint32_t)per_sm_epi * (uint32_t)prop.multiProcessorCount;
    }
    for (auto& s : ctx->slots) {
        if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&s.e0) != cudaSuccess ||
            cudaEventCreate(&s.e1) != cudaSuccess || cudaEventCreate(&s.e2) != cudaSuccess) {
            ctx->err = "event create failed"; return bail(MCRAW_ERR_CUDA);
        }
    }
    for (auto& g : ctx->stages) {
        if (cudaEventCreateWithFlags(&g.copied, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g.freed, cudaEventDisableTiming) != cudaSuccess) {
            ctx->err = "event create failed"; return bail(MCRAW_ERR_CUDA);
        }
    }
    *out = ctx;
    return MCRAW_OK;
}

void mcraw_ctx_destroy(mcraw_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& s : ctx->slots) {
        if (s.h_up) cudaFreeHost(s.h_up);
        if (s.d_up) cudaFree(s.d_up);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.d_dyn) cudaFree(s.d_dyn);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.done) cudaEventDestroy(s.done);
        if (s.e0) cudaEventDestroy(s.e0);
        if (s.e1) cudaEventDestroy(s.e1);
        if (s.e2) cudaEventDestroy(s.e2);
    }
    for (auto& g : ctx->stages) {
        if (g.d_buf) cudaFree(g.d_buf);
        if (g.copied) cudaEventDestroy(g.copied);
        if (g.freed) cudaEventDestroy(g.freed);
    }
    for (auto& e : ctx->part_done) if (e) cudaEventDestroy(e);
    if (ctx->h_in) cudaFreeHost(ctx->h_in);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    for (auto& cs : ctx->copy_streams) if (cs) cudaStreamDestroy(cs);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->plan_stream) cudaStreamDestroy(ctx->plan_stream);
    if (ctx->d2h_done) cudaEventDestroy(ctx->d2h_done);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mcraw_last_error(const mcraw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int mcraw_ctx_device(const mcraw_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t mcraw_kernel_launches(const mcraw_ctx* ctx) { return ctx ? ctx->launches : 0; }
float mcraw_last_batch_kernel_ms(const mcraw_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

int mcraw_set_kernel_timing(mcraw_ctx* ctx, uint32_t every_n_chunks) {
    if (!ctx) return MCRAW_ERR_ARG;
    ctx->timing_every = every_n_chunks;
    return MCRAW_OK;
}

int mcraw_kernel_time_totals(mcraw_ctx* ctx, double* meta_ms, double* main_ms, uint64_t* chunks) {
    if (!ctx) return MCRAW_ERR_ARG;
    int rc = bind(ctx);
    if (rc) return rc;
    for (auto& s : ctx->slots) { rc = harvest(ctx, s); if (rc) return rc; }
    if (meta_ms) *meta_ms = ctx->acc_meta_ms;
    if (main_ms) *main_ms = ctx->acc_main_ms;
    if (chunks) *chunks = ctx->acc_chunks;
    return MCRAW_OK;
}

int mcraw_decode_batch(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    return enqueue(ctx, descs, n, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int mcraw_decode_batch_levels(mcraw_ctx* ctx, const mcraw_frame_desc* descs, const mcraw_levels* levels, uint32_t n, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    return enqueue(ctx, descs, n, stream ? static_cast<cudaStream_t>(stream) : ctx->stream, levels);
}

int32_t mcraw_frame_encoded_width(const uint8_t* frame, uint64_t len, int32_t width, int32_t height) {
    mcraw_frame_desc d;
    std::memset(&d, 0, sizeof d);
    d.src = frame; d.len = len; d.width = width; d.height = height; d.compression_type = MCRAW_COMPRESSION_CURRENT;
    peek_encoded_width(d);
    return d.encoded_width == ((width + 63) / 64) * 64 ? 0 : d.encoded_width;
}

// Frames [first, first + n) of the current logical batch, sources in host memory (descs / host_dst point at frame `first`).
static int append_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint16_t* const* host_dst, uint32_t first, uint32_t n, void* stream);

int mcraw_decode_batch_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (!descs && n) return fail_arg(ctx, "descs is null");
    begin_batch(ctx, descs, n);
    return append_host(ctx, descs, nullptr, 0, n, stream);
}

int mcraw_decode_batch_host_out(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint16_t* const* host_dst, uint32_t n, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (!descs && n) return fail_arg(ctx, "descs is null");
    if (n && !host_dst) return fail_arg(ctx, "host_dst is null");
    begin_batch(ctx, descs, n);
    return append_host(ctx, descs, host_dst, 0, n, stream);
}

int mcraw_batch_begin(mcraw_ctx* ctx, uint32_t n_total) {
    if (!ctx) return MCRAW_ERR_ARG;
    begin_batch(ctx, nullptr, n_total);
    return MCRAW_OK;
}

int mcraw_batch_append_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t first_index, uint32_t count, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (!descs && count) return fail_arg(ctx, "descs is null");
    if (ctx->batch_id == 0 || (uint64_t)first_index + count > ctx->batch_n) return fail_arg(ctx, "append outside the batch announced by mcraw_batch_begin");
    for (uint32_t i = 0; i < count; i++) ctx->res_type[first_index + i] = descs[i].compression_type;
    return append_host(ctx, descs, nullptr, first_index, count, stream);
}

static int append_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint16_t* const* host_dst, uint32_t first, uint32_t n, void* stream) {
    int rc = bind(ctx);
    if (rc) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (host_dst && !ctx->d2h_stream) {
        CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->d2h_done, cudaEventDisableTiming));
    }
    std::vector<mcraw_frame_desc> chunk;
    uint32_t i = 0;
    int copy_rr = ctx->copy_rr;
    // Host output: the device -> host direction carries about twice the bytes and is the one to keep busy, and nothing
    // goes back before the first chunk has been copied in and decoded -- so the chunks start small and double
    // (MCRAW_HOSTOUT_FIRST_MB, 0 = full-size chunks from the start).
    size_t cap = host_dst && ctx->hostout_first_bytes ? std::min(kStageBytes, ctx->hostout_first_bytes) : kStageBytes;
    while (i < n) {
        // ---- pick the frames of this chunk: as many as fit one staging buffer (at least one)
        size_t bytes = 0;
        uint32_t j = i;
        while (j < n && j - i < kMaxGridY) {
            if (!descs[j].src) return fail_arg(ctx, "frame " + std::to_string(first + j) + ": null src");
            size_t need = (descs[j].len + 255) & ~(size_t)255;
            if (j > i && bytes + need > cap) break;
            bytes += need;
            j++;
        }
        Stage& g = ctx->stages[ctx->stage_cur];
        ctx->stage_cur = (ctx->stage_cur + 1) % kStage;
        cudaStream_t cs = ctx->copy_streams[copy_rr];
        copy_rr = (copy_rr + 1) % kCopyStreams;
        uint8_t* buf = g.d_buf;
        if (bytes > kStageBytes) {   // a single frame larger than a staging buffer: dedicated, synchronous allocation
            if (g.used) CU_TRY(ctx, cudaEventSynchronize(g.freed));
            if (g.d_buf) { cudaFree(g.d_buf); g.d_buf = nullptr; }
            CU_TRY(ctx, cudaMalloc(&g.d_buf, bytes));
            buf = g.d_buf;
        } else if (!buf) {
            CU_TRY(ctx, cudaMalloc(&g.d_buf, kStageBytes));
            buf = g.d_buf;
        }
        // ---- H2D on a side stream once the previous user of this staging buffer has been decoded
        if (g.used) CU_TRY(ctx, cudaStreamWaitEvent(cs, g.freed, 0));
        chunk.assign(descs + i, descs + j);
        for (mcraw_frame_desc& c : chunk) peek_encoded_width(c);
        // frames that lie back to back in host memory at the same 256-byte pitch (a pinned ring filled by a container
        // reader) travel in ONE copy: per-copy overhead is ~3 us, a 2 MB frame is ~40 us of PCIe time
        size_t off = 0;
        for (uint32_t k = 0; k < j - i;) {
            const uint8_t* run_src = descs[i + k].src;
            const size_t run_off = off;
            size_t run_bytes = 0;
            uint32_t m = k;
            for (; m < j - i; m++) {
                if (descs[i + m].src != run_src + run_bytes) break;
                chunk[m].src = buf + off;
                const size_t padded = (descs[i + m].len + 255) & ~(size_t)255;
                // the last frame of a run is copied without its padding (it may end the caller's buffer)
                run_bytes += padded;
                off += padded;
            }
            const size_t last_pad = ((descs[i + m - 1].len + 255) & ~(size_t)255) - descs[i + m - 1].len;
            CU_TRY(ctx, cudaMemcpyAsync(buf + run_off, run_src, run_bytes - last_pad, cudaMemcpyHostToDevice, cs));
            k = m;
        }
        CU_TRY(ctx, cudaEventRecord(g.copied, cs));
        // ---- decode on the main stream after the copy; then the buffer is free again
        CU_TRY(ctx, cudaStreamWaitEvent(st, g.copied, 0));
        rc = enqueue_chunk(ctx, chunk.data(), j - i, first + i, st);
        if (rc) return rc;
        CU_TRY(ctx, cudaEventRecord(g.freed, st));
        g.used = true;
        if (host_dst) {
            // pixels of this chunk go back while the next chunk is copied in and decoded (PCIe is full duplex); frames that
            // lie back to back on both sides travel in one copy
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->d2h_stream, g.freed, 0));
            for (uint32_t k = i; k < j; k++)
                if (!host_dst[k] || !descs[k].dst) return fail_arg(ctx, "frame " + std::to_string(k) + ": null dst");
            for (uint32_t k = i; k < j;) {
                // frames of one size at a constant pitch on both sides (a ring of output buffers) travel as ONE 2-D copy: a copy
                // per 4 MB frame leaves gaps on the link (240 of them per C2 batch); back to back they are one plain copy
                const size_t fbytes = 2 * (size_t)descs[k].width * (size_t)descs[k].height;
                uint32_t m = k + 1;
                ptrdiff_t dpitch = 0, hpitch = 0;
                if (m < j) {
                    dpitch = reinterpret_cast<const uint8_t*>(descs[m].dst) - reinterpret_cast<const uint8_t*>(descs[k].dst);
                    hpitch = reinterpret_cast<const uint8_t*>(host_dst[m]) - reinterpret_cast<const uint8_t*>(host_dst[k]);
                }
                if (m < j && dpitch >= (ptrdiff_t)fbytes && hpitch >= (ptrdiff_t)fbytes && dpitch < ((ptrdiff_t)1 << 30) && hpitch < ((ptrdiff_t)1 << 30)) {
                    for (; m < j; m++) {
                        if (2 * (size_t)descs[m].width * (size_t)descs[m].height != fbytes) break;
                        if (reinterpret_cast<const uint8_t*>(descs[m].dst) - reinterpret_cast<const uint8_t*>(descs[m - 1].dst) != dpitch) break;
                        if (reinterpret_cast<const uint8_t*>(host_dst[m]) - reinterpret_cast<const uint8_t*>(host_dst[m - 1]) != hpitch) break;
                    }
                } else m = k + 1;
                const size_t rows = m - k;
                if (rows > 1 && dpitch == (ptrdiff_t)fbytes && hpitch == (ptrdiff_t)fbytes)
                    CU_TRY(ctx, cudaMemcpyAsync(host_dst[k], descs[k].dst, fbytes * rows, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                else if (rows > 1) {
                    // (a run that spans separate allocations is refused by the runtime: frame by frame then)
                    if (cudaMemcpy2DAsync(host_dst[k], (size_t)hpitch, descs[k].dst, (size_t)dpitch, fbytes, rows, cudaMemcpyDeviceToHost,
                                          ctx->d2h_stream) != cudaSuccess) {
                        (void)cudaGetLastError();
                        for (uint32_t q = k; q < m; q++)
                            CU_TRY(ctx, cudaMemcpyAsync(host_dst[q], descs[q].dst, fbytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    }
                } else
                    CU_TRY(ctx, cudaMemcpyAsync(host_dst[k], descs[k].dst, fbytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                k = m;
            }
            CU_TRY(ctx, cudaEventRecord(ctx->d2h_done, ctx->d2h_stream));
            ctx->d2h_pending = true;
        }
        i = j;
        cap = std::min(kStageBytes, 2 * cap);
    }
    ctx->copy_rr = copy_rr;
    return MCRAW_OK;
}

int mcraw_batch_wait(mcraw_ctx* ctx, uint64_t* written_elems, uint32_t* status, uint32_t n) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (ctx->batch_id == 0) { ctx->err = "no batch was enqueued"; return MCRAW_ERR_STATE; }
    int rc = bind(ctx);
    if (rc) return rc;
    for (auto& s : ctx->slots) { rc = harvest(ctx, s); if (rc) return rc; }
    if (ctx->d2h_pending) {           // mcraw_decode_batch_host_out: the pixels have landed in the caller's host buffers
        CU_TRY(ctx, cudaEventSynchronize(ctx->d2h_done));
        ctx->d2h_pending = false;
    }
    const uint32_t m = std::min(n, ctx->batch_n);
    for (uint32_t i = 0; i < m; i++) {
        uint64_t w = ctx->res_written[i];
        uint32_t st = ctx->res_status[i];
        const int type = ctx->res_type[i];
        if (type != MCRAW_COMPRESSION_CURRENT && type != MCRAW_COMPRESSION_LEGACY) { w = 0; st = MCRAW_FRAME_BAD_TYPE; }
        if (written_elems) written_elems[i] = w;
        if (status) status[i] = st;
    }
    return MCRAW_OK;
}

size_t mcraw_decode_host(mcraw_ctx* ctx, uint16_t* output, int width, int height, const uint8_t* input, size_t len,
                         int compression_type) {
    if (!ctx || !output || !input || width <= 0 || height <= 0 || len == 0) return 0;
    if (bind(ctx)) return 0;
    const size_t out_elems = (size_t)width * (size_t)height, out_bytes = out_elems * 2;
    if (grow(ctx, ctx->h_in, ctx->h_in_cap, len + 16, true) || grow(ctx, ctx->d_in, ctx->d_in_cap, len + 16, false) ||
        grow(ctx, ctx->h_out, ctx->h_out_cap, out_bytes, true) || grow(ctx, ctx->d_out, ctx->d_out_cap, out_bytes, false))
        return 0;
    for (auto& e : ctx->part_done)
        if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return 0;
    // The caller's buffers are pageable: the bytes go through pinned staging, and for a 25 MB frame the two host copies
    // cost more than PCIe and the kernels together.  So large copies are cut into kHostParts pieces handled by as many
    // threads, and on the way back every piece is copied out as soon as ITS device-to-host transfer has landed.
    const int in_parts = len >= (2u << 20) ? ctx->host_parts : 1, out_parts = out_bytes >= (2u << 20) ? ctx->host_parts : 1;
    parallel_parts(in_parts, [&](int k) {
        const size_t a = len * k / in_parts, b = len * (k + 1) / in_parts;
        std::memcpy(reinterpret_cast<uint8_t*>(ctx->h_in) + a, input + a, b - a);
    });
    if (cudaMemcpyAsync(ctx->d_in, ctx->h_in, len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return 0;
    mcraw_frame_desc d;
    std::memset(&d, 0, sizeof d);
    d.src = ctx->d_in; d.len = len; d.width = width; d.height = height; d.compression_type = compression_type;
    d.dst = ctx->d_out; d.dst_capacity_elems = out_elems;
    {
        mcraw_frame_desc h = d;
        h.src = input;
        peek_encoded_width(h);
        d.encoded_width = h.encoded_width;
    }
    if (enqueue(ctx, &d, 1, ctx->stream)) return 0;   // the H2D copy above is still in flight on the stream
    // the transfers back are queued behind the kernels right away (16-byte aligned pieces)
    auto part_begin = [&](int k) { return (out_bytes * k / out_parts) & ~(size_t)15; };
    for (int k = 0; k < out_parts; k++) {
        const size_t a = part_begin(k), b = k + 1 == out_parts ? out_bytes : part_begin(k + 1);
        if (cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->h_out) + a, reinterpret_cast<uint8_t*>(ctx->d_out) + a, b - a,
                            cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaEventRecord(ctx->part_done[k], ctx->stream) != cudaSuccess)
            return 0;
    }
    uint64_t written = 0;
    if (mcraw_batch_wait(ctx, &written, nullptr, 1)) return 0;          // kernels done; the copies back are still running
    if (written == 0 || written > out_elems) { cudaStreamSynchronize(ctx->stream); return 0; }
    const size_t valid = (size_t)written * 2;                            // the reference writes exactly this much (RawData.cpp:611)
    const int device = ctx->device;
    std::atomic<bool> ok{true};
    parallel_parts(out_parts, [&](int k) {
        const size_t a = part_begin(k), b = std::min(valid, k + 1 == out_parts ? out_bytes : part_begin(k + 1));
        if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(ctx->part_done[k]) != cudaSuccess) { ok = false; return; }
        if (b > a) std::memcpy(reinterpret_cast<uint8_t*>(output) + a, reinterpret_cast<uint8_t*>(ctx->h_out) + a, b - a);
    });
    return ok ? (size_t)written : 0;
}

int mcraw_checksum_frames(mcraw_ctx* ctx, const uint16_t* const* frames_dev, const uint64_t* elems, uint32_t n, uint64_t* out,
                          void* stream) {
    if (!ctx || (n && (!frames_dev || !elems || !out))) return MCRAW_ERR_ARG;
    if (n == 0) return MCRAW_OK;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    // [pointers | element counts | sums] in one device allocation (not a hot path: allocated per call)
    uint8_t* d = nullptr;
    const size_t bytes = (size_t)n * 24;
    CU_TRY(ctx, cudaMalloc(&d, bytes));
    cudaError_t e = cudaMemcpyAsync(d, frames_dev, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + (size_t)n * 8, elems, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d + (size_t)n * 16, 0, (size_t)n * 8, st);
    if (e == cudaSuccess) {
        for (uint32_t base = 0; base < n && e == cudaSuccess; base += kMaxGridY) {
            const uint32_t m = std::min(kMaxGridY, n - base);
            k_checksum<<<dim3(64, m), CK_THREADS, 0, st>>>(reinterpret_cast<const uint16_t* const*>(d) + base,
                                                           reinterpret_cast<const unsigned long long*>(d + (size_t)n * 8) + base,
                                                           reinterpret_cast<unsigned long long*>(d + (size_t)n * 16) + base);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + (size_t)n * 16, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) { ctx->err = std::string("mcraw_checksum_frames: ") + cudaGetErrorString(e); (void)cudaGetLastError(); return MCRAW_ERR_CUDA; }
    return MCRAW_OK;
}

int mcraw_device_alloc(mcraw_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return MCRAW_OK;
}
int mcraw_device_free(mcraw_ctx* ctx, void* p) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaFree(p));
    return MCRAW_OK;
}
int mcraw_host_alloc_pinned(mcraw_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return MCRAW_OK;
}
int mcraw_host_free_pinned(mcraw_ctx* ctx, void* p) {
    if (!ctx) return MCRAW_ERR_ARG;
    CU_TRY(ctx, cudaFreeHost(p));
    return MCRAW_OK;
}
int mcraw_host_register(mcraw_ctx* ctx, void* p, size_t bytes, int read_only) {
    if (!ctx || !p || !bytes) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    if (read_only) {
        int supported = 0;
        cudaDeviceGetAttribute(&supported, cudaDevAttrHostRegisterReadOnlySupported, ctx->device);
        if (!supported) {
            ctx->err = "cudaHostRegister: read-only registration is not supported on this device/driver";
            return MCRAW_ERR_CUDA;
        }
    }
    // the range is pinned in whole pages
    const size_t page = 4096;
    const size_t span = (bytes + page - 1) / page * page;
    cudaError_t e = cudaHostRegister(p, span, cudaHostRegisterPortable | (read_only ? cudaHostRegisterReadOnly : 0));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        const cudaError_t first = e;
        e = cudaHostRegister(p, span, read_only ? cudaHostRegisterReadOnly : cudaHostRegisterDefault);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ctx->err = std::string("cudaHostRegister: ") + cudaGetErrorString(first) + " (portable), " + cudaGetErrorString(e) + " (default)";
            return MCRAW_ERR_CUDA;
        }
    }
    return MCRAW_OK;
}
int mcraw_host_unregister(mcraw_ctx* ctx, void* p) {
    if (!ctx || !p) return MCRAW_ERR_ARG;
    CU_TRY(ctx, cudaHostUnregister(p));
    return MCRAW_OK;
}
int mcraw_memcpy_h2d(mcraw_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, stream ? (cudaStream_t)stream : ctx->stream));
    return MCRAW_OK;
}
int mcraw_memcpy_d2h(mcraw_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, stream ? (cudaStream_t)stream : ctx->stream));
    return MCRAW_OK;
}
int mcraw_stream_sync(mcraw_ctx* ctx, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (bind(ctx)) return MCRAW_ERR_CUDA;
    CU_TRY(ctx, cudaStreamSynchronize(stream ? (cudaStream_t)stream : ctx->stream));
    return MCRAW_OK;
}

}  // extern "C"
