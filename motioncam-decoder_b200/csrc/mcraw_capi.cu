// mcraw_capi.cu -- implementation of the C-ABI in include/mcraw_b200.h on top of the kernels in
// mcraw_kernels.cuh.  Host side only does bookkeeping: descriptor validation, scratch sizing, launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mcraw_b200.h"
#include "mcraw_kernels.cuh"

using namespace mcraw;

namespace {

thread_local std::string g_create_error;

constexpr int kSlots = 4;              // batches that may be in flight per context
constexpr uint32_t kMaxGridY = 32768;  // frames per launch

struct Slot {
    FrameDev* h_frames = nullptr;   // pinned
    FrameDev* d_frames = nullptr;
    Result* h_results = nullptr;    // pinned
    Result* d_results = nullptr;
    uint32_t cap_frames = 0;
    uint8_t* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t done = nullptr, k_start = nullptr, k_stop = nullptr;
    bool in_flight = false;
    uint32_t n = 0;
};

}  // namespace

struct mcraw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t* d_tab = nullptr;
    Slot slots[kSlots];
    int cur = -1;
    uint64_t launches = 0;
    float last_kernel_ms = 0.f;
    std::string err;
    // staging for the single-frame host call
    uint8_t* h_in = nullptr; size_t h_in_cap = 0;
    uint8_t* d_in = nullptr; size_t d_in_cap = 0;
    uint16_t* h_out = nullptr; size_t h_out_cap = 0;
    uint16_t* d_out = nullptr; size_t d_out_cap = 0;
};

namespace {

#define CU_TRY(ctx, call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
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

int slot_reserve(mcraw_ctx* ctx, Slot& s, uint32_t n, size_t scratch) {
    if (n > s.cap_frames) {
        uint32_t cap = std::max<uint32_t>(n, s.cap_frames * 2 + 64);
        if (s.h_frames) cudaFreeHost(s.h_frames);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.d_frames) cudaFree(s.d_frames);
        if (s.d_results) cudaFree(s.d_results);
        s.h_frames = nullptr; s.h_results = nullptr; s.d_frames = nullptr; s.d_results = nullptr; s.cap_frames = 0;
        CU_TRY(ctx, cudaMallocHost(&s.h_frames, sizeof(FrameDev) * cap));
        CU_TRY(ctx, cudaMallocHost(&s.h_results, sizeof(Result) * cap));
        CU_TRY(ctx, cudaMalloc(&s.d_frames, sizeof(FrameDev) * cap));
        CU_TRY(ctx, cudaMalloc(&s.d_results, sizeof(Result) * cap));
        s.cap_frames = cap;
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

int pick_threads(uint32_t units) {
    // block size (multiple of 32, 96..256) that wastes the fewest lanes over ceil(units/threads) rounds
    int best = 256;
    double best_waste = 1e9;
    for (int t = 256; t >= 96; t -= 32) {
        uint32_t rounds = (units + t - 1) / t;
        double waste = (double)rounds * t / (double)units;
        if (waste < best_waste - 1e-9) { best_waste = waste; best = t; }
    }
    return best;
}

// Validate descriptors, fill the slot's FrameDev array, return scratch needed.
int prepare(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, std::vector<FrameDev>& out, size_t& scratch,
            uint32_t& max_tile_rows, uint32_t& max_units, bool& any7, bool& any6) {
    out.resize(n);
    scratch = 0; max_tile_rows = 0; max_units = 0; any7 = any6 = false;
    for (uint32_t i = 0; i < n; i++) {
        const mcraw_frame_desc& d = descs[i];
        FrameDev f;
        std::memset(&f, 0, sizeof f);
        if (!d.src || !d.dst) return fail_arg(ctx, "frame " + std::to_string(i) + ": null src/dst");
        if (d.width <= 0 || d.height <= 0 || d.width > 65536 || d.height > 65536)
            return fail_arg(ctx, "frame " + std::to_string(i) + ": unsupported width/height");
        if (((uintptr_t)d.src & 15) || ((uintptr_t)d.dst & 1))
            return fail_arg(ctx, "frame " + std::to_string(i) + ": src must be 16-byte aligned, dst 2-byte aligned");
        f.src = d.src; f.len = d.len; f.dst = d.dst; f.dst_cap = d.dst_capacity_elems;
        f.width = d.width; f.height = d.height; f.type = d.compression_type;
        f.tiles_x = (uint32_t)(d.width + 63) / 64;
        f.tile_rows = (uint32_t)(d.height + 3) / 4;
        f.flags = ((d.width % 8) == 0 && ((uintptr_t)d.dst & 15) == 0) ? FLAG_VEC_STORE : 0;
        if (d.compression_type == MCRAW_COMPRESSION_CURRENT) {
            any7 = true;
            f.tilemeta = reinterpret_cast<uint4*>(scratch);   // offset for now, rebased below
            scratch += (size_t)f.tiles_x * f.tile_rows * sizeof(uint4);
            max_tile_rows = std::max(max_tile_rows, f.tile_rows);
            max_units = std::max(max_units, f.tiles_x * 16u);
        } else if (d.compression_type == MCRAW_COMPRESSION_LEGACY) {
            any6 = true;
        } else {
            f.status = MCRAW_FRAME_BAD_TYPE;
        }
        out[i] = f;
    }
    return MCRAW_OK;
}

int enqueue(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, cudaStream_t st) {
    if (!descs && n) return fail_arg(ctx, "descs is null");
    int rc = bind(ctx);
    if (rc) return rc;
    ctx->cur = (ctx->cur + 1) % kSlots;
    Slot& s = ctx->slots[ctx->cur];
    if (s.in_flight) { CU_TRY(ctx, cudaEventSynchronize(s.done)); s.in_flight = false; }
    s.n = n;
    if (n == 0) { CU_TRY(ctx, cudaEventRecord(s.done, st)); s.in_flight = true; return MCRAW_OK; }

    std::vector<FrameDev> frames;
    size_t scratch; uint32_t max_tile_rows, max_units; bool any7, any6;
    rc = prepare(ctx, descs, n, frames, scratch, max_tile_rows, max_units, any7, any6);
    if (rc) return rc;
    rc = slot_reserve(ctx, s, n, scratch);
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++) {
        if (frames[i].type == MCRAW_COMPRESSION_CURRENT)
            frames[i].tilemeta = reinterpret_cast<uint4*>(s.d_scratch + reinterpret_cast<size_t>(frames[i].tilemeta));
        s.h_frames[i] = frames[i];
    }
    CU_TRY(ctx, cudaMemcpyAsync(s.d_frames, s.h_frames, sizeof(FrameDev) * n, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(s.d_results, 0, sizeof(Result) * n, st));
    CU_TRY(ctx, cudaEventRecord(s.k_start, st));
    for (uint32_t base = 0; base < n; base += kMaxGridY) {
        const uint32_t cnt = std::min(kMaxGridY, n - base);
        if (any7) {
            k_meta<<<2 * cnt, K1_THREADS, 0, st>>>(s.d_frames + base, ctx->d_tab);
            const int threads = pick_threads(max_units);
            k_tiles<<<dim3(max_tile_rows, cnt), threads, 0, st>>>(s.d_frames + base, ctx->d_tab, s.d_results + base);
            ctx->launches += 2;
        }
    }
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaEventRecord(s.k_stop, st));
    CU_TRY(ctx, cudaMemcpyAsync(s.h_results, s.d_results, sizeof(Result) * n, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaEventRecord(s.done, st));
    s.in_flight = true;
    return MCRAW_OK;
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
    std::vector<uint32_t> tab(MCRAW_TAB_ENTRIES * MCRAW_TAB_WORDS);
    mcraw_build_table(tab.data());
    if (cudaMalloc(&ctx->d_tab, tab.size() * 4) != cudaSuccess ||
        cudaMemcpy(ctx->d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        ctx->err = "table upload failed"; return bail(MCRAW_ERR_CUDA);
    }
    for (auto& s : ctx->slots) {
        if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&s.k_start) != cudaSuccess ||
            cudaEventCreate(&s.k_stop) != cudaSuccess) {
            ctx->err = "event create failed"; return bail(MCRAW_ERR_CUDA);
        }
    }
    *out = ctx;
    return MCRAW_OK;
}

void mcraw_ctx_destroy(mcraw_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->slots) {
        if (s.in_flight && s.done) cudaEventSynchronize(s.done);
        if (s.h_frames) cudaFreeHost(s.h_frames);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.d_frames) cudaFree(s.d_frames);
        if (s.d_results) cudaFree(s.d_results);
        if (s.d_scratch) cudaFree(s.d_scratch);
        if (s.done) cudaEventDestroy(s.done);
        if (s.k_start) cudaEventDestroy(s.k_start);
        if (s.k_stop) cudaEventDestroy(s.k_stop);
    }
    if (ctx->h_in) cudaFreeHost(ctx->h_in);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->d_in) cudaFree(ctx->d_in);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_tab) cudaFree(ctx->d_tab);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mcraw_last_error(const mcraw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int mcraw_ctx_device(const mcraw_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t mcraw_kernel_launches(const mcraw_ctx* ctx) { return ctx ? ctx->launches : 0; }
float mcraw_last_batch_kernel_ms(const mcraw_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.f; }

int mcraw_decode_batch(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream) {
    if (!ctx) return MCRAW_ERR_ARG;
    return enqueue(ctx, descs, n, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int mcraw_batch_wait(mcraw_ctx* ctx, uint64_t* written_elems, uint32_t* status, uint32_t n) {
    if (!ctx) return MCRAW_ERR_ARG;
    if (ctx->cur < 0 || !ctx->slots[ctx->cur].in_flight) { ctx->err = "no batch in flight"; return MCRAW_ERR_STATE; }
    int rc = bind(ctx);
    if (rc) return rc;
    Slot& s = ctx->slots[ctx->cur];
    CU_TRY(ctx, cudaEventSynchronize(s.done));
    s.in_flight = false;
    if (s.n) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.k_start, s.k_stop) == cudaSuccess) ctx->last_kernel_ms = ms;
    }
    const uint32_t m = std::min(n, s.n);
    for (uint32_t i = 0; i < m; i++) {
        uint64_t w = s.h_results[i].written;
        uint32_t st = s.h_results[i].status;
        const int type = s.h_frames[i].type;   // host copy of the descriptor
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
    const size_t out_elems = (size_t)width * (size_t)height;
    if (grow(ctx, ctx->h_in, ctx->h_in_cap, len + 16, true) || grow(ctx, ctx->d_in, ctx->d_in_cap, len + 16, false) ||
        grow(ctx, ctx->h_out, ctx->h_out_cap, out_elems * 2, true) || grow(ctx, ctx->d_out, ctx->d_out_cap, out_elems * 2, false))
        return 0;
    std::memcpy(ctx->h_in, input, len);
    if (cudaMemcpyAsync(ctx->d_in, ctx->h_in, len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return 0;
    mcraw_frame_desc d;
    std::memset(&d, 0, sizeof d);
    d.src = ctx->d_in; d.len = len; d.width = width; d.height = height; d.compression_type = compression_type;
    d.dst = ctx->d_out; d.dst_capacity_elems = out_elems;
    if (enqueue(ctx, &d, 1, ctx->stream)) return 0;
    uint64_t written = 0;
    if (mcraw_batch_wait(ctx, &written, nullptr, 1)) return 0;
    if (written == 0 || written > out_elems) return 0;
    if (cudaMemcpyAsync(ctx->h_out, ctx->d_out, written * 2, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return 0;
    std::memcpy(output, ctx->h_out, written * 2);
    return (size_t)written;
}

int mcraw_decode_batch_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream) {
    (void)descs; (void)n; (void)stream;
    if (!ctx) return MCRAW_ERR_ARG;
    ctx->err = "mcraw_decode_batch_host: not built yet";
    return MCRAW_ERR_STATE;
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
