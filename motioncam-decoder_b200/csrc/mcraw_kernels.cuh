// mcraw_kernels.cuh -- sm_100a kernels of the MCRAW frame decoder.
//
// Current format (compressionType 7; reference: /root/reference/lib/RawData.cpp:528-612)
//
//   A frame is a sequence of 64-sample blocks, four per 64x4-pixel tile.  We process it in UNITS of 16
//   consecutive tiles (= 64 blocks = exactly one 64-value block of each metadata stream).
//
//   k_meta   one CTA per (frame, metadata stream).  Resolves the inline-header chain of the stream
//            (RawData.cpp:463-498) by pointer doubling over every candidate position of a staged window and records, per
//            unit, where its metadata block is and what its header says.  For the "bits" stream ONE LANE also decodes one
//            whole 64-value block (same width-specialised SWAR routine as the pixel kernel) to turn the reference's running
//            `offset +=` (RawData.cpp:562,576-579) into prefix sums: the payload offset of every unit.
//   k_units  persistent, launched as a programmatic dependent of k_meta; every WARP takes items (a few consecutive units of
//            one frame) from a queue and decodes one unit at a time.  The unit's payload (contiguous, <= 8 KiB) arrives in
//            shared memory by ONE bulk copy (cp.async.bulk, TMA 1-D, tracked by the warp's mbarrier), its two metadata
//            blocks by 16-byte cp.async with zero fill; every lane pulls
//            the two bits values and the two references of ITS block pair out of the metadata blocks (table-driven,
//            mcraw_meta_table.h), a warp scan gives the pair's payload offset, and the lane decodes the pair: a `switch`
//            on the header bits value selects straight-line code with immediate shifts/masks (lanes that share
//            a bits value run together; real images have 1-3 distinct values per warp).  Even/odd columns are
//            interleaved with PRMT, references added with packed 16-bit adds (mod 2^16 like RawData.cpp:582-592),
//            rows are assembled in shared memory and leave as coalesced 16-byte stores; columns >= width are
//            cropped (RawData.cpp:598-608).
//
// Legacy format (compressionType 6; reference: /root/reference/lib/RawData_Legacy.cpp:445-495): mcraw_legacy.cuh.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcraw_b200.h"
#include "mcraw_meta_table.h"

namespace mcraw {

struct FrameDev {
    const uint8_t* src;            // device, 16-byte aligned
    unsigned long long len;
    uint16_t* dst;                 // device
    unsigned long long dst_cap;    // elements
    int width, height, type;
    unsigned tiles_x;              // planned encodedWidth/64: ceil(width/64), or the descriptor's encoded_width / 64
    unsigned tile_rows;            // expected ceil(encodedHeight/4) upper bound = ceil(height/4)
    unsigned flags;                // FLAG_*
    unsigned inv_tiles_x;          // ceil(2^32 / tiles_x) when tile / tiles_x may use mulhi, else 0
    unsigned nunits;               // ceil(tiles_x*tile_rows / 16)
    uint32_t* unitoff;             // scratch [nunits + 1]   payload offset of each unit (+ end)
    uint4* metarec;                // scratch [nunits]  where the unit's two metadata blocks are: {offset of the bits block,
                                   //                   offset of the refs block, header of the bits block, header of the refs block}
    // optional epilogue fused into the pixel kernels (mcraw_decode_batch_levels): 0 = raw values (the reference's output)
    unsigned epi_mode;             // MCRAW_OUT_*
    unsigned epi_black2[2];        // per row parity: black level of the even column | odd column << 16 (integers)
    unsigned epi_range2[2];        // per row parity: white - black, packed the same way
    float epi_blackf[4];           // black level per CFA position (row parity * 2 + column parity)
    float epi_scalef[4];           // 1 / (white - black)
    void* sp_scratch;              // k_meta_split scratch: two streams x (maps, flags), see ks_stream_scratch
    uint32_t* lg_tilemap;          // legacy scratch [tiles][17]   exit of every entry, for tiles whose exits differ (k_legacy_warp)
    unsigned long long* lg_status; // legacy scratch [tiles][2]    epoch-tagged look-back words of every tile: count, exit
};

// Per-frame words written by the index kernels and read by the pixel kernels.  Every word is written on every path
// (no host-side zeroing): status[s] by the CTA of metadata stream s (not used by the legacy kernel).
struct FrameState {
    unsigned status[2];            // MCRAW_FRAME_* bits
    unsigned tile_rows_dev;        // ceil(encodedHeight/4) from the frame header
    unsigned rows_fit;             // rows emitted: min(4*tile_rows_dev, dst_cap / width)
    unsigned done[2];              // per metadata stream: the launch epoch of the slot (mcraw_capi.cu: Slot::flag_uses) once the index
                                   // kernel has published the stream for this launch; k_units waits until both carry its own epoch.
                                   // Stale values are older epochs, so nothing is zeroed between launches or plans (8-byte aligned pair)
};
static_assert(sizeof(FrameState) == 24, "done[] is read as one 64-bit word: FrameState arrays must keep it 8-byte aligned");

// The plan of a launch (FrameDev[] + work lists) may arrive by a copy on ANOTHER stream while the kernels are already
// resident (mcraw_capi.cu, enqueue_chunk): the copy ends with the plan's epoch in a ready word, which one thread per warp /
// CTA polls before it touches the plan.  Plan data is then read with ld.global.cg (L2): the same addresses held the
// slot's previous plan, of which nothing may come back from L1.
__device__ __forceinline__ void plan_wait(const uint32_t* ready, const uint32_t epoch) {
    uint32_t v;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ready) : "memory");
        if (v == epoch) break;
        __nanosleep(128);
    }
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ready) : "memory");
}
template <class T>
__device__ __forceinline__ T* ldcg_ptr(T* const* p) {
    return reinterpret_cast<T*>(__ldcg(reinterpret_cast<const unsigned long long*>(p)));
}
// what the index kernels need of a frame's descriptor
struct FrameIdx {
    const uint8_t* src;
    unsigned long long len, dst_cap;
    int width, type;
    unsigned tiles_x, tile_rows;
    uint32_t* unitoff;
    uint4* metarec;
};
__device__ __forceinline__ FrameIdx frame_idx_cg(const FrameDev* p) {
    FrameIdx f;
    f.src = ldcg_ptr(&p->src); f.len = __ldcg(&p->len); f.dst_cap = __ldcg(&p->dst_cap);
    f.width = __ldcg(&p->width); f.type = __ldcg(&p->type); f.tiles_x = __ldcg(&p->tiles_x); f.tile_rows = __ldcg(&p->tile_rows);
    f.unitoff = ldcg_ptr(&p->unitoff); f.metarec = ldcg_ptr(&p->metarec);
    return f;
}

struct Result {
    unsigned long long written;    // uint16 elements, 0 = failed
    unsigned status;
    unsigned pad;
};

enum { FLAG_VEC_STORE = 1 };       // width % 8 == 0 and dst 16-byte aligned: rows may leave as 16-byte multiples

__device__ __forceinline__ uint32_t ld_u32le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// payload bytes / 8 of a 64-sample block for header value b in 0..15 (RawData.cpp:27-45): PRMT as a byte LUT
__device__ __forceinline__ uint32_t cur_len8_nib(uint32_t b) {
    const uint32_t lo = __byte_perm(0x03020100u, 0x08060504u, b & 7u);        // b = 0..7  -> 0,1,2,3,4,5,6,8
    const uint32_t hi = __byte_perm(0x100A0A08u, 0x10101010u, b & 7u);        // b = 8..15 -> 8,10,10,16,16,...
    return ((b & 8u) ? hi : lo) & 0xFFu;
}
__device__ __forceinline__ uint32_t cur_len8(uint32_t b) { return b >= 16u ? 16u : cur_len8_nib(b); }

// --------------------------------------------------------------------------------------------------------
// Width-specialised block decode.  G(g) returns the g-th 8-byte group of the block payload as uint2
// (.x = byte lanes 0..3, .y = byte lanes 4..7).  Sample i = 8*j + l sits in byte lane l of plane j;
// L[2j], L[2j+1] receive the low bytes of plane j (lanes 0..3 / 4..7), H[..] the high bytes.
// Restates RawData.cpp:112-408 as 32-bit SWAR; the reference's left shifts are folded into net right shifts
// with pre-shifted masks, e.g. ((G2 >> 6) & 1) << 2 == (G2 >> 4) & 0x04.
// Returns the payload length in bytes (RawData.cpp:27-45).
// --------------------------------------------------------------------------------------------------------
#define MC_REP(m) ((uint32_t)(m) * 0x01010101u)

template <class Fetch>
__device__ __forceinline__ uint32_t decode_block(const uint32_t b, const Fetch& G, uint32_t (&L)[16], uint32_t (&H)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i++) { L[i] = 0; H[i] = 0; }
    switch (b) {
    case 0:
        return 0;
    case 1: {                                                                  // RawData.cpp:112-136
        const uint2 g = G(0);
#pragma unroll
        for (int j = 0; j < 8; j++) { L[2 * j] = (g.x >> j) & MC_REP(1); L[2 * j + 1] = (g.y >> j) & MC_REP(1); }
        return 8;
    }
    case 2: {                                                                  // :138-162
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const uint2 g = G(m);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                L[2 * (4 * m + k)] = (g.x >> (2 * k)) & MC_REP(3);
                L[2 * (4 * m + k) + 1] = (g.y >> (2 * k)) & MC_REP(3);
            }
        }
        return 16;
    }
    case 3: {                                                                  // :164-199
        const uint2 g0 = G(0), g1 = G(1), g2 = G(2);
        L[0] = g0.x & MC_REP(7);                 L[1] = g0.y & MC_REP(7);
        L[2] = (g0.x >> 3) & MC_REP(7);          L[3] = (g0.y >> 3) & MC_REP(7);
        L[4] = ((g0.x >> 6) & MC_REP(3)) | ((g2.x >> 4) & MC_REP(4));
        L[5] = ((g0.y >> 6) & MC_REP(3)) | ((g2.y >> 4) & MC_REP(4));
        L[6] = g1.x & MC_REP(7);                 L[7] = g1.y & MC_REP(7);
        L[8] = (g1.x >> 3) & MC_REP(7);          L[9] = (g1.y >> 3) & MC_REP(7);
        L[10] = ((g1.x >> 6) & MC_REP(3)) | ((g2.x >> 5) & MC_REP(4));
        L[11] = ((g1.y >> 6) & MC_REP(3)) | ((g2.y >> 5) & MC_REP(4));
        L[12] = g2.x & MC_REP(7);                L[13] = g2.y & MC_REP(7);
        L[14] = (g2.x >> 3) & MC_REP(7);         L[15] = (g2.y >> 3) & MC_REP(7);
        return 24;
    }
    case 4: {                                                                  // :201-223
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const uint2 g = G(m);
            L[4 * m] = g.x & MC_REP(15);         L[4 * m + 1] = g.y & MC_REP(15);
            L[4 * m + 2] = (g.x >> 4) & MC_REP(15); L[4 * m + 3] = (g.y >> 4) & MC_REP(15);
        }
        return 32;
    }
    case 5: {                                                                  // :225-262
        const uint2 g0 = G(0), g1 = G(1), g2 = G(2), g3 = G(3), g4 = G(4);
        L[0] = g0.x & MC_REP(31); L[1] = g0.y & MC_REP(31);
        L[2] = g1.x & MC_REP(31); L[3] = g1.y & MC_REP(31);
        L[4] = g2.x & MC_REP(31); L[5] = g2.y & MC_REP(31);
        L[6] = g3.x & MC_REP(31); L[7] = g3.y & MC_REP(31);
        L[8] = g4.x & MC_REP(31); L[9] = g4.y & MC_REP(31);
        L[10] = ((g0.x >> 5) & MC_REP(7)) | ((g3.x >> 2) & MC_REP(0x18));
        L[11] = ((g0.y >> 5) & MC_REP(7)) | ((g3.y >> 2) & MC_REP(0x18));
        L[12] = ((g1.x >> 5) & MC_REP(7)) | ((g4.x >> 2) & MC_REP(0x18));
        L[13] = ((g1.y >> 5) & MC_REP(7)) | ((g4.y >> 2) & MC_REP(0x18));
        L[14] = ((g2.x >> 5) & MC_REP(7)) | ((g3.x >> 4) & MC_REP(0x08)) | ((g4.x >> 3) & MC_REP(0x10));
        L[15] = ((g2.y >> 5) & MC_REP(7)) | ((g3.y >> 4) & MC_REP(0x08)) | ((g4.y >> 3) & MC_REP(0x10));
        return 40;
    }
    case 6: {                                                                  // :264-304
        const uint2 g0 = G(0), g1 = G(1), g2 = G(2), g3 = G(3), g4 = G(4), g5 = G(5);
        L[0] = g0.x & MC_REP(63);  L[1] = g0.y & MC_REP(63);
        L[2] = g1.x & MC_REP(63);  L[3] = g1.y & MC_REP(63);
        L[4] = g2.x & MC_REP(63);  L[5] = g2.y & MC_REP(63);
        L[6] = g3.x & MC_REP(63);  L[7] = g3.y & MC_REP(63);
        L[8] = g4.x & MC_REP(63);  L[9] = g4.y & MC_REP(63);
        L[10] = g5.x & MC_REP(63); L[11] = g5.y & MC_REP(63);
        L[12] = ((g0.x >> 6) & MC_REP(3)) | ((g1.x >> 4) & MC_REP(0x0C)) | ((g2.x >> 2) & MC_REP(0x30));
        L[13] = ((g0.y >> 6) & MC_REP(3)) | ((g1.y >> 4) & MC_REP(0x0C)) | ((g2.y >> 2) & MC_REP(0x30));
        L[14] = ((g3.x >> 6) & MC_REP(3)) | ((g4.x >> 4) & MC_REP(0x0C)) | ((g5.x >> 2) & MC_REP(0x30));
        L[15] = ((g3.y >> 6) & MC_REP(3)) | ((g4.y >> 4) & MC_REP(0x0C)) | ((g5.y >> 2) & MC_REP(0x30));
        return 48;
    }
    case 7:
    case 8: {                                                                  // :306-326,446-449
#pragma unroll
        for (int j = 0; j < 8; j++) { const uint2 g = G(j); L[2 * j] = g.x; L[2 * j + 1] = g.y; }
        return 64;
    }
    case 9:
    case 10: {                                                                 // :328-374,450-453
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const uint2 hg = G(5 * m + 4);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint2 g = G(5 * m + k);
                L[2 * (4 * m + k)] = g.x;        L[2 * (4 * m + k) + 1] = g.y;
                H[2 * (4 * m + k)] = (hg.x >> (2 * k)) & MC_REP(3);
                H[2 * (4 * m + k) + 1] = (hg.y >> (2 * k)) & MC_REP(3);
            }
        }
        return 80;
    }
    default: {                                                                 // 11..16, :376-408 (little-endian u16)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint2 a = G(2 * j), c = G(2 * j + 1);
            L[2 * j] = __byte_perm(a.x, a.y, 0x6420);     H[2 * j] = __byte_perm(a.x, a.y, 0x7531);
            L[2 * j + 1] = __byte_perm(c.x, c.y, 0x6420); H[2 * j + 1] = __byte_perm(c.x, c.y, 0x7531);
        }
        return 128;
    }
    }
}

// --------------------------------------------------------------------------------------------------------
// PTX helpers
// --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// --------------------------------------------------------------------------------------------------------
// k_meta
// --------------------------------------------------------------------------------------------------------
// Two shapes of the kernel: windows of 16 KiB worked by 256 threads (three CTAs per SM: batches, where throughput counts),
// and windows of 48 KiB worked by 1024 threads (one CTA per SM: a handful of frames, where the latency of one stream's
// chain counts -- a third of the windows, each as fast as a small one).  Window offsets are 16 bits wide: < 64 KiB.
template <int CHUNK, int THREADS>
struct K1 {
    static constexpr int K1_THREADS = THREADS;
    static constexpr int K1_CHUNK = CHUNK;        // bytes of the stream staged in shared memory per round
    static constexpr int K1_MB = THREADS;         // meta blocks per round (= threads: one lane decodes one meta block)
    static constexpr int K1_NXT_BYTES = K1_CHUNK + 16;                       // u16 per even offset, byte offset == stream offset
    static constexpr int K1_SMEM = (K1_CHUNK + 32) + 2 * K1_NXT_BYTES + 2 * (K1_MB + 8) + 2 * (K1_MB / 4 + 8);   // stage, nxt1, nxt4, ends, anchors
    static constexpr int K1_STAGE_PER_THREAD = ((K1_CHUNK + 32) / 16 + K1_THREADS - 1) / K1_THREADS;
    static_assert(K1_CHUNK % (16 * K1_THREADS) == 0 && K1_CHUNK < 65536, "window shape");
};
using K1Batch = K1<16384, 256>;
using K1Few = K1<49152, 1024>;

// group fetch from the staged stream at a 2-byte aligned position (meta block payloads follow a 2-byte header)
struct StageFetch {
    const uint32_t* w;   // staged bytes viewed as words
    uint32_t a;          // byte offset of the payload inside the stage buffer (even)
    __device__ __forceinline__ uint2 operator()(int g) const {
        const uint32_t o = a + 8u * (uint32_t)g;
        const uint32_t i = o >> 2, sh = (o & 3u) * 8u;
        const uint32_t w0 = w[i], w1 = w[i + 1], w2 = w[i + 2];
        return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
    }
};


__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// One block of the bits stream, staged at byte p of `stage` (p even): its 64 values are the header bits of the 64 pixel blocks
// of unit `unit` (value = unpacked + header reference, mod 2^16, RawData.cpp:491-492).  Returns the payload length of
// the unit / 8 (the sum of the 64 block lengths, :27-45); bad != 0 when a value of a live tile exceeds 16 (the reference
// would index its length table out of bounds) -- padding values behind the last tile are ignored.
__device__ __forceinline__ uint32_t bits_block_len8(const uint8_t* stage, const uint32_t p, const uint32_t unit, const uint32_t ntiles,
                                                    uint32_t& bad) {
    const uint32_t b = stage[p] >> 4;                                                  // RawData.cpp:106-110
    const uint32_t ref = ((uint32_t)(stage[p] & 0x0F) << 8) | stage[p + 1];
    uint32_t L[16], H[16];
    StageFetch G{reinterpret_cast<const uint32_t*>(stage), p + 2u};
    decode_block(b, G, L, H);
    // word m of L/H holds samples 4m..4m+3 = the four blocks of tile 16*unit + m
    bad = (ref > 16u) ? 1u : 0u;
    const uint32_t refb = MC_REP(ref & 0x1F);
    uint32_t acc[2] = {0, 0};                                                          // byte-wise sums of 8 words each: <= 128 per byte
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const uint32_t tile = unit * 16u + m;
        uint32_t v = (L[m] & MC_REP(0x1F)) + refb;                                     // bytes <= 31 + 16: no carries
        uint32_t badm = H[m] | (L[m] & MC_REP(0xE0));
        badm |= (v + MC_REP(0x6F)) & MC_REP(0x80);                                     // a byte > 16 (reference: OOB table read)
        if (tile >= ntiles) { v = 0; badm = 0; }                                       // padding values are ignored
        bad |= badm;
        acc[m >> 3] += mcraw_len8x4(v & MC_REP(0x1F));                                 // four block lengths at once
    }
    const uint32_t s2 = (acc[0] & 0x00FF00FFu) + ((acc[0] >> 8) & 0x00FF00FFu) +
                        (acc[1] & 0x00FF00FFu) + ((acc[1] >> 8) & 0x00FF00FFu);
    return (s2 & 0xFFFFu) + (s2 >> 16);
}

// grid = 2 * frames, block = Shape::K1_THREADS, dynamic smem = Shape::K1_SMEM
// Publish everything this CTA wrote for (frame, stream): the barrier orders every thread's writes before thread 0, whose
// fence (cumulative) and store of the launch epoch form the release; k_units may be running already (programmatic
// dependent launch) and polls the frame's two words, ending the poll with an acquire load.
__device__ __forceinline__ void meta_done_store(FrameState& S, const uint32_t stream, const uint32_t epoch) {
    __threadfence();
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(&S.done[stream]), "r"(epoch) : "memory");
}
__device__ __forceinline__ void meta_publish(FrameState& S, const uint32_t stream, const uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x == 0) meta_done_store(S, stream, epoch);
}

template <class Shape>
__global__ void __launch_bounds__(Shape::K1_THREADS, Shape::K1_THREADS == 256 ? 3 : 1) k_meta(const FrameDev* frames, FrameState* __restrict__ states,
                                                                                               const uint32_t epoch, const uint32_t* plan_ready, const uint32_t plan_epoch) {
    constexpr int K1_THREADS = Shape::K1_THREADS, K1_CHUNK = Shape::K1_CHUNK, K1_MB = Shape::K1_MB, K1_NXT_BYTES = Shape::K1_NXT_BYTES;
    constexpr int K1_STAGE = Shape::K1_STAGE_PER_THREAD;
    extern __shared__ __align__(16) uint8_t k1_smem[];
    uint8_t* stage = k1_smem;                                                        // K1_CHUNK + 32 bytes of the stream
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t nxt1_s = stage_s + K1_CHUNK + 32;                                 // next position after 1 block
    const uint32_t nxt4_s = nxt1_s + K1_NXT_BYTES;                                   // ... after 4 blocks
    uint16_t* ends = reinterpret_cast<uint16_t*>(k1_smem + (K1_CHUNK + 32) + 2 * K1_NXT_BYTES);     // [K1_MB] position after block t
    uint16_t* anchors = ends + (K1_MB + 8);                                        // [K1_MB / 4 + 1]
    __shared__ uint32_t warp_sums[K1_THREADS / 32];
    __shared__ uint32_t sh_err, sh_bad;
    __shared__ uint32_t sh_hdr[4];

    // k_units may be scheduled as soon as every CTA of this grid has got this far: it waits per frame, not per grid
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const int f = blockIdx.x >> 1;
    const int stream = blockIdx.x & 1;   // 0 = bits, 1 = refs
    const int tid = threadIdx.x;
    if (tid == 0) plan_wait(plan_ready, plan_epoch);
    __syncthreads();
    const FrameIdx F = frame_idx_cg(frames + f);
    FrameState& S = states[f];
    if (F.type != MCRAW_COMPRESSION_CURRENT) return;
    const uint8_t* __restrict__ src = F.src;
    const unsigned long long len = F.len;

    if (tid == 0) {
        uint32_t err = 0;
        uint32_t ew = 0, eh = 0, boff = 0, roff = 0;
        if (len < 16) err = MCRAW_FRAME_BAD_HEADER;
        else {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));                // RawData.cpp:500-524 (little-endian u32 x 4)
            ew = h.x; eh = h.y; boff = h.z; roff = h.w;
            if (boff > len || roff > len) err |= MCRAW_FRAME_BAD_HEADER;            // :547
            if (ew % 64u) err |= MCRAW_FRAME_BAD_HEADER;                            // :550
            if (F.width <= 0 || ew < (uint32_t)F.width) err |= MCRAW_FRAME_BAD_HEADER;  // :553
            if (ew == 0 || eh == 0) err |= MCRAW_FRAME_BAD_HEADER;
            if (!err) {
                if (ew / 64u != F.tiles_x) err |= MCRAW_FRAME_GEOMETRY;
                if ((eh + 3u) / 4u > F.tile_rows) err |= MCRAW_FRAME_GEOMETRY;
            }
        }
        const uint32_t tr = err ? 0u : (eh + 3u) / 4u;
        unsigned long long pos0 = stream ? roff : boff;
        if (!err) {
            const uint32_t nb = (ew / 64u) * tr * 4u;
            if (pos0 + 4 > len) err = MCRAW_FRAME_TRUNCATED;
            else if (ld_u32le(src + pos0) < nb) err = MCRAW_FRAME_BAD_META_COUNT;   // RawData.cpp:470-476
        }
        sh_hdr[0] = ew; sh_hdr[1] = eh; sh_hdr[2] = boff; sh_hdr[3] = roff;
        sh_err = err;
        sh_bad = 0;
        if (stream == 0) {
            unsigned long long fit = F.dst_cap / (unsigned long long)(F.width > 0 ? F.width : 1);
            if (fit > 4ull * tr) fit = 4ull * tr;                                   // reference emits 4 rows per tile row (:598-608)
            S.tile_rows_dev = tr;
            S.rows_fit = (uint32_t)fit;
        }
    }
    __syncthreads();
    if (sh_err) {
        if (tid == 0) S.status[stream] = sh_err;
        meta_publish(S, (uint32_t)stream, epoch);
        return;
    }
    const uint32_t tiles_x = sh_hdr[0] / 64u;
    const uint32_t tile_rows = (sh_hdr[1] + 3u) / 4u;
    const uint32_t ntiles = tiles_x * tile_rows;
    const uint32_t need_mb = (ntiles * 4u + 63u) / 64u;  // = number of units
    unsigned long long pos = (unsigned long long)sh_hdr[2 + stream] + 4;
    const unsigned long long par = pos & 1ull;

    uint32_t* __restrict__ unitoff = F.unitoff;
    uint32_t done = 0;            // meta blocks (= units) finished
    uint32_t carry = 16;          // running payload offset, METADATA_OFFSET (RawData.cpp:25,562)

    while (done < need_mb) {
        // ---- stage [base, base + K1_CHUNK + 32) of the frame buffer in shared memory (zero past len).  Block lengths are
        //      even (2 + 8m), so every position of a chain has the parity of its start: the candidates of the window are the
        //      even offsets from `base`, which is 16-byte aligned for streams that start at an even offset (every encoder
        //      output) and one byte further for the others (the reference reads bytes, RawData.cpp:463-498: any offset goes).
        const unsigned long long base = ((pos - par) & ~15ull) + par;
        {
            uint4 q[K1_STAGE];
#pragma unroll
            for (int k = 0; k < K1_STAGE; k++) {
                const int v = tid + k * K1_THREADS;
                const unsigned long long o = base + (unsigned long long)v * 16;
                q[k] = make_uint4(0, 0, 0, 0);
                if (v < (K1_CHUNK + 32) / 16) {
                    if (!par && o + 16 <= len) q[k] = __ldg(reinterpret_cast<const uint4*>(src + o));
                    else if (o < len) {
                        uint32_t t4[4] = {0, 0, 0, 0};
                        for (int e = 0; e < 16; e++)
                            if (o + e < len) t4[e >> 2] |= (uint32_t)src[o + e] << (8 * (e & 3));
                        q[k] = make_uint4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K1_STAGE; k++) {
                const int v = tid + k * K1_THREADS;
                if (v < (K1_CHUNK + 32) / 16) *reinterpret_cast<uint4*>(stage + v * 16) = q[k];
            }
        }
        __syncthreads();
        // ---- nxt1[p] = p + 2 + payload length of the header at p for every even p whose block fits the window and
        //      the frame (RawData.cpp:419), else a self-looping sentinel.  16 stream bytes = 8 candidates per step.
        const unsigned long long room = len - base;
        const uint32_t lim = (uint32_t)(room < (unsigned long long)K1_CHUNK ? room : (unsigned long long)K1_CHUNK);
        constexpr uint32_t SENT = K1_CHUNK;
#pragma unroll
        for (int k = 0; k < K1_CHUNK / 16 / K1_THREADS; k++) {
            const uint32_t v = tid + k * K1_THREADS;
            const uint4 d = lds128(stage_s + 16u * v);
            const uint32_t w[4] = {d.x, d.y, d.z, d.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t pr[2];
#pragma unroll
                for (int hlf = 0; hlf < 2; hlf++) {
                    const uint32_t p = 16u * v + 4u * i + 2u * hlf;
                    const uint32_t hb = (w[i] >> (16 * hlf + 4)) & 15u;
                    uint32_t q = p + 2u + 8u * cur_len8_nib(hb);
                    if (q > lim) q = SENT;
                    pr[hlf] = q;
                }
                o[i] = pr[0] | (pr[1] << 16);
            }
            sts128(nxt1_s + 16u * v, o[0], o[1], o[2], o[3]);
        }
        if (tid == 0) { sts128(nxt1_s + SENT, SENT * 0x10001u, SENT * 0x10001u, SENT * 0x10001u, SENT * 0x10001u);
                        sts128(nxt4_s + SENT, SENT * 0x10001u, SENT * 0x10001u, SENT * 0x10001u, SENT * 0x10001u); }
        __syncthreads();
        // ---- pointer doubling: nxt2 = nxt1 o nxt1 (kept in the nxt4 array), then nxt4 = nxt2 o nxt2 in place
#pragma unroll
        for (int k = 0; k < K1_CHUNK / 16 / K1_THREADS; k++) {
            const uint32_t v = tid + k * K1_THREADS;
            const uint4 d = lds128(nxt1_s + 16u * v);
            const uint32_t w[4] = {d.x, d.y, d.z, d.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) o[i] = lds_u16(nxt1_s + (w[i] & 0xFFFFu)) | (lds_u16(nxt1_s + (w[i] >> 16)) << 16);
            sts128(nxt4_s + 16u * v, o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        {
            uint32_t o[K1_CHUNK / 16 / K1_THREADS][4];
#pragma unroll
            for (int k = 0; k < K1_CHUNK / 16 / K1_THREADS; k++) {
                const uint32_t v = tid + k * K1_THREADS;
                const uint4 d = lds128(nxt4_s + 16u * v);
                const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int i = 0; i < 4; i++) o[k][i] = lds_u16(nxt4_s + (w[i] & 0xFFFFu)) | (lds_u16(nxt4_s + (w[i] >> 16)) << 16);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K1_CHUNK / 16 / K1_THREADS; k++) sts128(nxt4_s + 16u * (tid + k * K1_THREADS), o[k][0], o[k][1], o[k][2], o[k][3]);
        }
        __syncthreads();
        // ---- serial part of the chain walk (RawData.cpp:485-489), 4 blocks per dependent shared-memory load
        const uint32_t limit = min((uint32_t)K1_MB, need_mb - done);
        if (tid == 0) {
            uint32_t p = (uint32_t)(pos - base);
            const uint32_t hops = (limit >> 2) + 1u;
#pragma unroll 4
            for (uint32_t k = 0; k < hops; k++) {
                anchors[k] = (uint16_t)p;
                p = lds_u16(nxt4_s + p);
            }
        }
        __syncthreads();
        // ---- every thread finishes its own position: t blocks from the start = anchor[t / 4] + (t % 4) single steps
        uint32_t my_start = SENT;
        bool valid = false;
        if ((uint32_t)tid < limit) {
            uint32_t p = anchors[tid >> 2];
            for (int r = 0; r < (tid & 3); r++) p = lds_u16(nxt1_s + p);
            my_start = p;
            const uint32_t e = lds_u16(nxt1_s + p);                       // SENT when the block does not fit
            ends[tid] = (uint16_t)e;
            valid = p != SENT && e != SENT;
        }
        const uint32_t cnt = (uint32_t)__syncthreads_count(valid);     // validity is monotone along the chain
        if (cnt < limit && lim < (uint32_t)K1_CHUNK) {                  // the chain ran into the end of the FRAME
            if (tid == 0) sh_err = MCRAW_FRAME_TRUNCATED;
            __syncthreads();
            break;
        }
        const unsigned long long nextpos = cnt ? base + ends[cnt - 1] : pos;
        // ---- where the block is and what its header says go to the unit's record: k_units extracts the two values each
        //      of its lanes needs straight from the block.  The bits stream is decoded here as well (one lane, one block of
        //      64 values: value = unpacked + header reference, mod 2^16, :491-492) because the payload offset of a unit is
        //      the sum of the lengths of all blocks before it.
        uint32_t unit_len8 = 0;
        const uint32_t unit = done + (uint32_t)tid;
        if ((uint32_t)tid < cnt) {
            const uint32_t p = my_start;
            const uint32_t hdr = (uint32_t)stage[p] | ((uint32_t)stage[p + 1] << 8);
            uint32_t* rec = reinterpret_cast<uint32_t*>(F.metarec + unit);
            rec[stream] = (uint32_t)(base + p);
            rec[2 + stream] = hdr;
            if (stream == 0) {
                uint32_t bad;
                unit_len8 = bits_block_len8(stage, p, unit, ntiles, bad);
                if (bad) sh_bad = 1;
            }
        }
        if (stream == 0) {
            // ---- exclusive prefix sum of the unit payload lengths across the CTA (+ carry from earlier rounds)
            uint32_t incl = unit_len8;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((tid & 31) >= d) incl += o;
            }
            if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
            __syncthreads();
            uint32_t wbase = 0, all = 0;
#pragma unroll
            for (int w = 0; w < K1_THREADS / 32; w++) {
                const uint32_t s = warp_sums[w];
                if (w < (tid >> 5)) wbase += s;
                all += s;
            }
            if ((uint32_t)tid < cnt) unitoff[unit] = carry + 8u * (wbase + incl - unit_len8);
            carry += 8u * all;
            if (sh_bad) { if (tid == 0) sh_err = MCRAW_FRAME_BAD_BITS; __syncthreads(); break; }
        }
        done += cnt;
        pos = nextpos;
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t err = sh_err;
        if (stream == 0) {
            if (!err && (unsigned long long)carry > len) err = MCRAW_FRAME_TRUNCATED;   // RawData.cpp:419
            unitoff[need_mb] = carry;
        }
        S.status[stream] = err;
    }
    meta_publish(S, (uint32_t)stream, epoch);
}

// --------------------------------------------------------------------------------------------------------
// k_meta_warp: the index work of a BATCH, one WARP per (frame, metadata stream).  k_meta spends a CTA of 256 threads and
// 50 KB of shared memory on a stream and computes next pointers for every candidate offset three times over (pointer
// doubling buys latency with work: 15 M warp instructions per C2 batch, a tenth of the SM time of the pixel kernel it is
// supposed to hide behind).  A batch has hundreds of streams, so latency per stream is not what counts: here a warp walks
// its stream through 2 KiB windows at fixed addresses (window j = stream bytes [S0 + j W, S0 + (j + 1) W + 144): a block
// that starts inside a window ends at most 130 bytes behind it), double-buffered by bulk copies (TMA 1-D, one mbarrier per
// buffer).  Per window: next pointers for its 1024 even offsets (once), then rounds of up to 32 hops of the chain taken
// by all lanes together -- lane k keeps the start of block k -- followed by the parallel part exactly as in k_meta: the
// unit's record, and for the bits stream the 64 values of the block -> payload length of the unit -> prefix sums (warp
// scan + carry) -> unitoff.  Same results word for word as k_meta (RawData.cpp:463-498,562,576-579).
// grid = ceil(2 * frames / 4), block = 128, dynamic smem = KW::SMEM
// --------------------------------------------------------------------------------------------------------
struct KW {
    static constexpr int WARPS = 4, THREADS = 32 * WARPS;
    static constexpr int W = 2048;                   // stream bytes a window owns (chain positions p < W)
    static constexpr int MARGIN = 144;               // 130 bytes of the last block + the group fetch's look-ahead, 16-multiple
    static constexpr int STAGE = W + MARGIN;
    static constexpr int WARP_SMEM = 2 * STAGE + W + 32;   // two window buffers, u16 next pointer per even offset (byte offset == stream offset), two mbarriers
    static constexpr int SMEM = WARPS * WARP_SMEM;
    static constexpr uint32_t NONE = 0xFFFFu;        // next pointer of a block that does not fit the frame
    static_assert(STAGE % 16 == 0 && WARP_SMEM % 16 == 0 && W % 512 == 0, "window shape");
    static_assert(SMEM + 4096 < 0xFFFF, "next pointers are 16-bit shared-memory addresses");
};

__global__ void __launch_bounds__(KW::THREADS, 6) k_meta_warp(const FrameDev* frames, FrameState* __restrict__ states,
                                                              const uint32_t npairs, const uint32_t epoch, const uint32_t* plan_ready,
                                                              const uint32_t plan_epoch) {
    extern __shared__ __align__(128) uint8_t kw_smem[];
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t pair = blockIdx.x * KW::WARPS + wid;
    if (pair >= npairs) return;
    const uint32_t f = pair >> 1, stream = pair & 1u;            // 0 = bits, 1 = refs
    if (lane == 0) plan_wait(plan_ready, plan_epoch);
    __syncwarp();
    const FrameIdx F = frame_idx_cg(frames + f);
    FrameState& S = states[f];
    if (F.type != MCRAW_COMPRESSION_CURRENT) return;
    uint8_t* const sm = kw_smem + wid * KW::WARP_SMEM;
    const uint32_t sm_s = smem_u32(sm);
    const uint32_t nxt_s = sm_s + 2u * KW::STAGE;
    const uint32_t bar_s = nxt_s + KW::W;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_s) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_s + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();

    const uint8_t* __restrict__ src = F.src;
    const unsigned long long len = F.len;
    // ---- frame header (RawData.cpp:500-524,547-560): every lane reads the same 16 bytes
    uint32_t err = 0, ew = 0, eh = 0, boff = 0, roff = 0;
    if (len < 16) err = MCRAW_FRAME_BAD_HEADER;
    else {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));
        ew = h.x; eh = h.y; boff = h.z; roff = h.w;
        if (boff > len || roff > len) err |= MCRAW_FRAME_BAD_HEADER;
        if (ew % 64u) err |= MCRAW_FRAME_BAD_HEADER;
        if (F.width <= 0 || ew < (uint32_t)F.width) err |= MCRAW_FRAME_BAD_HEADER;
        if (ew == 0 || eh == 0) err |= MCRAW_FRAME_BAD_HEADER;
        if (!err) {
            if (ew / 64u != F.tiles_x) err |= MCRAW_FRAME_GEOMETRY;
            if ((eh + 3u) / 4u > F.tile_rows) err |= MCRAW_FRAME_GEOMETRY;
        }
    }
    const uint32_t tile_rows = err ? 0u : (eh + 3u) / 4u;
    if (stream == 0 && lane == 0) {
        unsigned long long fit = F.dst_cap / (unsigned long long)(F.width > 0 ? F.width : 1);
        if (fit > 4ull * tile_rows) fit = 4ull * tile_rows;        // the reference emits 4 rows per tile row (:598-608)
        S.tile_rows_dev = tile_rows;
        S.rows_fit = (uint32_t)fit;
    }
    const unsigned long long pos0 = stream ? roff : boff;          // the stream starts with its 32-bit value count
    if (!err && pos0 + 4 > len) err = MCRAW_FRAME_TRUNCATED;
    const uint32_t ntiles = (ew / 64u) * tile_rows;
    const uint32_t need_mb = (ntiles * 4u + 63u) / 64u;            // = number of units
    const unsigned long long par = pos0 & 1ull;                    // block lengths are even: the chain keeps the parity of its start
    const unsigned long long S0 = ((pos0 - par) & ~15ull) + par;   // window 0 starts here (16-byte aligned for even streams)

    // a window that lies inside the buffer and starts 16-byte aligned arrives by one bulk copy, any other by plain loads
    auto bulk_able = [&](const uint32_t j) { return !par && S0 + (unsigned long long)j * KW::W + KW::STAGE <= len; };
    auto issue_bulk = [&](const uint32_t j) {
        if (lane == 0) {
            const uint32_t bar = bar_s + 8u * (j & 1u);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)KW::STAGE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(sm_s + (j & 1u) * KW::STAGE), "l"(src + S0 + (unsigned long long)j * KW::W), "r"((uint32_t)KW::STAGE), "r"(bar) : "memory");
        }
    };
    uint32_t phase0 = 0, phase1 = 0;                               // uses of the two mbarriers
    auto wait_bulk = [&](const uint32_t j) {
        const uint32_t bar = bar_s + 8u * (j & 1u);
        const uint32_t ph = (j & 1u) ? phase1 : phase0;
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(bar), "r"(ph & 1u) : "memory");
        }
        if (j & 1u) phase1++; else phase0++;
    };
    auto stage_plain = [&](const uint32_t j) {
        const unsigned long long base = S0 + (unsigned long long)j * KW::W;
        for (uint32_t c = lane; c < (uint32_t)KW::STAGE / 16u; c += 32u) {
            const unsigned long long o = base + 16ull * c;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (!par && o + 16 <= len) q = __ldg(reinterpret_cast<const uint4*>(src + o));
            else if (o < len) {
                uint32_t t4[4] = {0, 0, 0, 0};
                for (int e = 0; e < 16; e++)
                    if (o + e < len) t4[e >> 2] |= (uint32_t)src[o + e] << (8 * (e & 3));
                q = make_uint4(t4[0], t4[1], t4[2], t4[3]);
            }
            sts128(sm_s + (j & 1u) * KW::STAGE + 16u * c, q.x, q.y, q.z, q.w);
        }
        __syncwarp();
    };

    uint32_t done = 0;                    // meta blocks (= units) finished
    uint32_t carry = 16;                  // running payload offset, METADATA_OFFSET (RawData.cpp:25,562)
    int pending = -1;                     // window whose bulk copy is in flight
    if (!err) {
        if (bulk_able(0)) { issue_bulk(0); pending = 0; }
        uint32_t p = (uint32_t)(pos0 - S0) + 4u;                   // chain position relative to the window (even, < W)
        uint32_t* __restrict__ unitoff = F.unitoff;
        uint4* __restrict__ metarec = F.metarec;
        for (uint32_t j = 0;; j++) {
            if (pending == (int)j) wait_bulk(j); else stage_plain(j);
            pending = -1;
            __syncwarp();
            if (bulk_able(j + 1u)) { issue_bulk(j + 1u); pending = (int)(j + 1u); }   // buffer (j + 1) & 1 was read for the last time a window ago
            const uint8_t* stage = sm + (j & 1u) * KW::STAGE;
            const uint32_t stage_s = sm_s + (j & 1u) * KW::STAGE;
            const unsigned long long base = S0 + (unsigned long long)j * KW::W;
            if (j == 0) {                                          // RawData.cpp:470-476
                const uint32_t c = (uint32_t)(pos0 - S0);
                const uint32_t count = lds_u16(stage_s + c) | (lds_u16(stage_s + c + 2u) << 16);
                if (count < ntiles * 4u) { err = MCRAW_FRAME_BAD_META_COUNT; break; }
            }
            // ---- next pointer of every even offset of the window: p + 2 + payload length of the header at p (:419), stored as
            //      the shared-memory ADDRESS of that offset's own entry (a CTA's window is < 64 KiB, so it fits 16 bits and a hop of
            //      the chain is one dependent load and nothing else); NONE when the block does not fit the frame
            const unsigned long long room = len > base ? len - base : 0ull;
            const uint32_t lim = (uint32_t)(room < 0xFFF0ull ? room : 0xFFF0ull);
#pragma unroll
            for (int k = 0; k < KW::W / 16 / 32; k++) {
                const uint32_t v = lane + 32u * k;
                const uint4 d = lds128(stage_s + 16u * v);
                const uint32_t w[4] = {d.x, d.y, d.z, d.w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t pr[2];
#pragma unroll
                    for (int hlf = 0; hlf < 2; hlf++) {
                        const uint32_t q = 16u * v + 4u * i + 2u * hlf + 2u + 8u * cur_len8_nib((w[i] >> (16 * hlf + 4)) & 15u);
                        pr[hlf] = q > lim ? KW::NONE : nxt_s + q;
                    }
                    o[i] = pr[0] | (pr[1] << 16);
                }
                sts128(nxt_s + 16u * v, o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
            // ---- rounds of up to 32 blocks
            bool stop = false;
            const uint32_t a_end = nxt_s + (uint32_t)KW::W;          // entries at or behind it belong to the next window
            uint32_t a = nxt_s + p;
            while (a < a_end) {
                const uint32_t want = min(32u, need_mb - done);
                uint32_t my_a = 0, cnt = 0;
#pragma unroll 4
                for (; cnt < want; cnt++) {                        // the same hop in every lane (one broadcast load each)
                    if (a >= a_end) break;                         // (NONE ends the walk here as well)
                    if (lane == cnt) my_a = a;
                    a = lds_u16(a);
                }
                if (a == KW::NONE) { err = MCRAW_FRAME_TRUNCATED; stop = true; break; }   // the chain ran into the end of the frame
                uint32_t unit_len8 = 0, bad = 0;
                const uint32_t unit = done + lane;
                if (lane < cnt) {
                    const uint32_t my_start = my_a - nxt_s;
                    uint32_t* rec = reinterpret_cast<uint32_t*>(metarec + unit);
                    rec[stream] = (uint32_t)(base + my_start);
                    rec[2 + stream] = lds_u16(stage_s + my_start);
                    if (stream == 0) unit_len8 = bits_block_len8(stage, my_start, unit, ntiles, bad);
                }
                if (stream == 0) {
                    uint32_t incl = unit_len8;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= (uint32_t)d) incl += o;
                    }
                    if (lane < cnt) unitoff[unit] = carry + 8u * (incl - unit_len8);
                    carry += 8u * __shfl_sync(0xFFFFFFFFu, incl, 31);
                    if (__any_sync(0xFFFFFFFFu, bad != 0u)) { err = MCRAW_FRAME_BAD_BITS; stop = true; break; }
                }
                done += cnt;
                if (done >= need_mb) { stop = true; break; }
            }
            p = a - nxt_s;
            if (stop) break;
            p -= (uint32_t)KW::W;
            __syncwarp();
        }
        if (pending >= 0) wait_bulk((uint32_t)pending);            // nothing may still be landing in shared memory when the warp leaves
    }
    if (lane == 0) {
        if (stream == 0) {
            if (!err && (unsigned long long)carry > len) err = MCRAW_FRAME_TRUNCATED;   // RawData.cpp:419
            if (!(err & (MCRAW_FRAME_BAD_HEADER | MCRAW_FRAME_GEOMETRY))) F.unitoff[need_mb] = carry;
        }
        S.status[stream] = err;
    }
    // publish: the warp barrier orders every lane's writes before lane 0, whose fence (cumulative) and epoch store form the release
    __syncwarp();
    if (lane == 0) meta_done_store(S, stream, epoch);
}

// --------------------------------------------------------------------------------------------------------
// k_meta_split: the chain of ONE stream resolved by several CTAs at once (a handful of big frames: k_meta's windows are
// serial, 6 rounds of 48 KiB for the refs stream of a 4080x3072 frame).  CTA (frame, stream, w) owns window w = stream bytes
// [W0 + w C, W0 + (w + 1) C), C = 8 KiB.  A block that starts inside a window belongs to it (it ends at most 130 bytes behind it: the
// window is staged with that margin), so a chain enters window w + 1 at one of 66 even offsets, and the effect of a window
// on the chain is a map  entry -> (exit, blocks walked).  Phases:
//   1. stage, next pointers for every candidate, doubled twice (as k_meta); 66 lanes walk the 66 entries through the
//      window -> the window's map, published with a flag;
//   2. wait for the maps of the windows before (all CTAs of a launch are resident: the host only picks this kernel when
//      they fit), compose them in shared memory -> this window's true entry and the number of blocks before it;
//   3. the true chain: anchors + per-thread positions + unit records (+ bits stream: decode, unit payload lengths) exactly
//      as k_meta does per round; the payload offsets of a window's units are first written relative to the window;
//   4. bits stream: the windows' payload totals are exchanged the same way and every window shifts its units' offsets;
//      the last window collects the error bits and publishes the stream (its done word) for the frame.
// Chains that merge or not makes no difference here: maps are composed, never guessed.  Flags carry the launch epoch of the
// plan (nothing is zeroed between launches).
// --------------------------------------------------------------------------------------------------------
struct KS {
#ifndef MCRAW_KS_WINDOW
#define MCRAW_KS_WINDOW 8192
#endif
    static constexpr int C = MCRAW_KS_WINDOW;    // window bytes
    static constexpr int EXT = 160;              // staged behind the window: <= 130 bytes of a straddling block + group-fetch slack
#ifndef MCRAW_KS_THREADS
#define MCRAW_KS_THREADS 512
#endif
    static constexpr int THREADS = MCRAW_KS_THREADS;
    static constexpr int NE = 66;                // entry states: even offsets 0 .. 130
    static constexpr int MAXW = 32 * (16384 / C);   // windows per stream (512 KiB of stream)
    static constexpr int STAGE_BYTES = C + EXT;
    static constexpr int NXT_BYTES = C + EXT + 16;                          // u16 per even offset, byte offset == window offset
    static constexpr uint32_t SENT = C + EXT;                               // dead chain (a block that does not fit the frame)
    static constexpr int SMEM = STAGE_BYTES + 4 * NXT_BYTES + 4 * MAXW * NE;   // stage, next pointers after 1 / 4 / 16 / 64 blocks, maps
    static constexpr int STAGE_PER_THREAD = (STAGE_BYTES / 16 + THREADS - 1) / THREADS;
    static_assert(C % (16 * THREADS) == 0 && SENT < 65535 && STAGE_BYTES % 16 == 0, "window shape");
};
// scratch of one (frame, stream): maps [nw][NE] u32 (exit offset behind the window | blocks << 16; 0xFFFF = dead), then
// three flag words per window (u64: epoch << 32 | payload): map ready, payload total, done + error bits
__host__ __device__ inline size_t ks_stream_scratch(uint32_t nw) { return (size_t)nw * (KS::NE * 4 + 3 * 8); }

__device__ __forceinline__ void ks_flag_store(unsigned long long* p, uint32_t epoch, uint32_t payload) {
    const unsigned long long v = ((unsigned long long)epoch << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ks_flag_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// all threads: wait until flags[0 .. n) carry `epoch` (n <= THREADS); returns false if it gave up (never expected)
__device__ __forceinline__ bool ks_wait_flags(const unsigned long long* flags, uint32_t n, uint32_t epoch, uint32_t tid, uint32_t& payload) {
    payload = 0;
    for (uint32_t spins = 0;; spins++) {
        bool ok = true;
        if (tid < n) {
            const unsigned long long v = ks_flag_load(flags + tid);
            ok = (uint32_t)(v >> 32) == epoch;
            payload = (uint32_t)v;
        }
        if (__syncthreads_and(ok)) break;
        if (spins > (1u << 20)) return false;
        __nanosleep(100);
    }
    __threadfence();                                           // the flags before what they announce
    return true;
}

// dst = src o src o src o src for the candidates inside the window (positions behind it are fixed points of every table):
// src o src first (kept in dst), then squared in place
__device__ __forceinline__ void ks_quadruple(const uint32_t src_s, const uint32_t dst_s, const int tid) {
    constexpr int K = KS::C / 16 / KS::THREADS;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t v = tid + k * KS::THREADS;
        const uint4 d = lds128(src_s + 16u * v);
        const uint32_t wd[4] = {d.x, d.y, d.z, d.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = lds_u16(src_s + (wd[i] & 0xFFFFu)) | (lds_u16(src_s + (wd[i] >> 16)) << 16);
        sts128(dst_s + 16u * v, o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    uint32_t o[K][4];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t v = tid + k * KS::THREADS;
        const uint4 d = lds128(dst_s + 16u * v);
        const uint32_t wd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int i = 0; i < 4; i++) o[k][i] = lds_u16(dst_s + (wd[i] & 0xFFFFu)) | (lds_u16(dst_s + (wd[i] >> 16)) << 16);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) sts128(dst_s + 16u * (tid + k * KS::THREADS), o[k][0], o[k][1], o[k][2], o[k][3]);
    __syncthreads();
}

// grid = frames * 2 * nw, block = KS::THREADS, dynamic smem = KS::SMEM
__global__ void __launch_bounds__(KS::THREADS, 1) k_meta_split(const FrameDev* __restrict__ frames, FrameState* __restrict__ states,
                                                               const uint32_t nw, const uint32_t epoch) {
    constexpr int C = KS::C, THREADS = KS::THREADS, NE = KS::NE;
    constexpr uint32_t SENT = KS::SENT;
    extern __shared__ __align__(16) uint8_t k1_smem[];
    uint8_t* stage = k1_smem;
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t nxt1_s = stage_s + KS::STAGE_BYTES;
    const uint32_t nxt4_s = nxt1_s + KS::NXT_BYTES;
    const uint32_t nxt16_s = nxt4_s + KS::NXT_BYTES;
    const uint32_t nxt64_s = nxt16_s + KS::NXT_BYTES;
    uint32_t* cmap = reinterpret_cast<uint32_t*>(k1_smem + KS::STAGE_BYTES + 4 * KS::NXT_BYTES);   // [w][NE] maps of the windows before this one
    __shared__ uint32_t warp_sums[THREADS / 32];
    __shared__ uint32_t sh_err, sh_bad, sh_entry, sh_cb;
    __shared__ uint32_t sh_hdr[4];

    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
#ifdef MCRAW_KS_DEBUG
    long long dbg_t[8]; int dbg_n = 0;
#define KS_STAMP() do { if (threadIdx.x == 0 && dbg_n < 8) dbg_t[dbg_n++] = clock64(); } while (0)
#else
#define KS_STAMP() do { } while (0)
#endif
    KS_STAMP();
    const uint32_t w = blockIdx.x % nw;
    const uint32_t fs = blockIdx.x / nw;
    const int f = (int)(fs >> 1);
    const int stream = (int)(fs & 1u);   // 0 = bits, 1 = refs
    const FrameDev& F = frames[f];
    FrameState& S = states[f];
    if (F.type != MCRAW_COMPRESSION_CURRENT) return;
    const int tid = threadIdx.x;
    const uint8_t* __restrict__ src = F.src;
    const unsigned long long len = F.len;
    uint8_t* const sbase = reinterpret_cast<uint8_t*>(F.sp_scratch) + (size_t)stream * ks_stream_scratch(nw);
    uint32_t* const gmap = reinterpret_cast<uint32_t*>(sbase);
    unsigned long long* const flag1 = reinterpret_cast<unsigned long long*>(sbase + (size_t)nw * NE * 4);
    unsigned long long* const flag2 = flag1 + nw;
    unsigned long long* const flag3 = flag2 + nw;

    if (tid == 0) {                                            // the frame header, as k_meta reads it (every window does)
        uint32_t err = 0;
        uint32_t ew = 0, eh = 0, boff = 0, roff = 0;
        if (len < 16) err = MCRAW_FRAME_BAD_HEADER;
        else {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));                // RawData.cpp:500-524
            ew = h.x; eh = h.y; boff = h.z; roff = h.w;
            if (boff > len || roff > len) err |= MCRAW_FRAME_BAD_HEADER;            // :547
            if (ew % 64u) err |= MCRAW_FRAME_BAD_HEADER;                            // :550
            if (F.width <= 0 || ew < (uint32_t)F.width) err |= MCRAW_FRAME_BAD_HEADER;  // :553
            if (ew == 0 || eh == 0) err |= MCRAW_FRAME_BAD_HEADER;
            if (!err) {
                if (ew / 64u != F.tiles_x) err |= MCRAW_FRAME_GEOMETRY;
                if ((eh + 3u) / 4u > F.tile_rows) err |= MCRAW_FRAME_GEOMETRY;
            }
        }
        const uint32_t tr = err ? 0u : (eh + 3u) / 4u;
        const unsigned long long pos0 = stream ? roff : boff;
        if (!err) {
            const uint32_t nb = (ew / 64u) * tr * 4u;
            if (pos0 + 4 > len) err = MCRAW_FRAME_TRUNCATED;
            else if (ld_u32le(src + pos0) < nb) err = MCRAW_FRAME_BAD_META_COUNT;   // RawData.cpp:470-476
        }
        sh_hdr[0] = ew; sh_hdr[1] = eh; sh_hdr[2] = boff; sh_hdr[3] = roff;
        sh_err = err;
        sh_bad = 0;
        if (stream == 0 && w == 0) {
            unsigned long long fit = F.dst_cap / (unsigned long long)(F.width > 0 ? F.width : 1);
            if (fit > 4ull * tr) fit = 4ull * tr;
            S.tile_rows_dev = tr;
            S.rows_fit = (uint32_t)fit;
        }
    }
    __syncthreads();
    const uint32_t hdr_err = sh_err;
    const uint32_t tiles_x = sh_hdr[0] / 64u;
    const uint32_t tile_rows = (sh_hdr[1] + 3u) / 4u;
    const uint32_t ntiles = tiles_x * tile_rows;
    const uint32_t need_mb = hdr_err ? 0u : (ntiles * 4u + 63u) / 64u;      // = number of units
    const unsigned long long pos0 = (unsigned long long)sh_hdr[2 + stream] + 4;
    const unsigned long long par = pos0 & 1ull;
    const unsigned long long W0 = ((pos0 - par) & ~15ull) + par;
    const unsigned long long base = W0 + (unsigned long long)w * C;
    uint32_t* __restrict__ unitoff = F.unitoff;
    uint32_t done = 0, first = 0, local_carry = 0;            // blocks finished up to here, blocks before this window, payload bytes / 8 of this window's units
    bool dead = false;                                         // the chain ran into the end of the frame inside this window

    if (!hdr_err) {
        // ---- 1a. stage [base, base + C + EXT) (zero past len); an odd stream offset is staged byte by byte
        {
            uint4 q[KS::STAGE_PER_THREAD];
#pragma unroll
            for (int k = 0; k < KS::STAGE_PER_THREAD; k++) {
                const int v = tid + k * THREADS;
                const unsigned long long o = base + (unsigned long long)v * 16;
                q[k] = make_uint4(0, 0, 0, 0);
                if (v < KS::STAGE_BYTES / 16) {
                    if (!par && o + 16 <= len) q[k] = __ldg(reinterpret_cast<const uint4*>(src + o));
                    else if (o < len) {
                        uint32_t t4[4] = {0, 0, 0, 0};
                        for (int e = 0; e < 16; e++)
                            if (o + e < len) t4[e >> 2] |= (uint32_t)src[o + e] << (8 * (e & 3));
                        q[k] = make_uint4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < KS::STAGE_PER_THREAD; k++) {
                const int v = tid + k * THREADS;
                if (v < KS::STAGE_BYTES / 16) *reinterpret_cast<uint4*>(stage + v * 16) = q[k];
            }
        }
        __syncthreads();
        KS_STAMP();
        // ---- 1b. nxt1[p] = p + 2 + payload length for every candidate p < C whose block fits the frame (RawData.cpp:419), else
        //      SENT; positions behind the window (where a block that starts inside can end) and SENT loop on themselves
        const unsigned long long room = len > base ? len - base : 0ull;
#pragma unroll
        for (int k = 0; k < C / 16 / THREADS; k++) {
            const uint32_t v = tid + k * THREADS;
            const uint4 d = lds128(stage_s + 16u * v);
            const uint32_t wd[4] = {d.x, d.y, d.z, d.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t pr[2];
#pragma unroll
                for (int hlf = 0; hlf < 2; hlf++) {
                    const uint32_t p = 16u * v + 4u * i + 2u * hlf;
                    const uint32_t hb = (wd[i] >> (16 * hlf + 4)) & 15u;
                    uint32_t q = p + 2u + 8u * cur_len8_nib(hb);
                    if ((unsigned long long)q > room) q = SENT;
                    pr[hlf] = q;
                }
                o[i] = pr[0] | (pr[1] << 16);
            }
            sts128(nxt1_s + 16u * v, o[0], o[1], o[2], o[3]);
        }
        for (uint32_t v = C / 16 + tid; v < (uint32_t)KS::NXT_BYTES / 16; v += THREADS) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { const uint32_t p = 16u * v + 4u * i; o[i] = p | ((p + 2u) << 16); }
            sts128(nxt1_s + 16u * v, o[0], o[1], o[2], o[3]);
            sts128(nxt4_s + 16u * v, o[0], o[1], o[2], o[3]);
            sts128(nxt16_s + 16u * v, o[0], o[1], o[2], o[3]);
            sts128(nxt64_s + 16u * v, o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        // ---- 1c. pointer doubling: next position after 4, 16 and 64 blocks.  A dense window (2- to 12-byte blocks: the bits
        //      stream of a smooth image) holds well over a thousand blocks; with these tables no walk through it is longer than
        //      a few dozen dependent loads, and every thread finds its own block without a serial pass
        ks_quadruple(nxt1_s, nxt4_s, tid);
        ks_quadruple(nxt4_s, nxt16_s, tid);
        ks_quadruple(nxt16_s, nxt64_s, tid);
        KS_STAMP();
        // ---- 1d. the window's map: entry 2 t -> (where the chain leaves the window, blocks it walked inside)
        if (tid < NE && w + 1 < nw) {                          // (nobody reads the last window's map)
            uint32_t p = 2u * (uint32_t)tid, cnt = 0;
            for (;;) { const uint32_t q = lds_u16(nxt64_s + p); if (q >= (uint32_t)C) break; p = q; cnt += 64; }
            for (;;) { const uint32_t q = lds_u16(nxt16_s + p); if (q >= (uint32_t)C) break; p = q; cnt += 16; }
            for (;;) { const uint32_t q = lds_u16(nxt4_s + p); if (q >= (uint32_t)C) break; p = q; cnt += 4; }
            while (p < (uint32_t)C) {                              // (a block that does not fit the frame is not counted)
                p = lds_u16(nxt1_s + p);
                if (p != SENT) cnt++;
            }
            gmap[(size_t)w * NE + tid] = (p == SENT ? 0xFFFFu : p - (uint32_t)C) | (cnt << 16);
        }
        __syncthreads();
        if (tid == 0 && w + 1 < nw) { __threadfence(); ks_flag_store(flag1 + w, epoch, 0u); }

        KS_STAMP();
        // ---- 2. this window's entry and the blocks before it
        uint32_t dummy;
        bool ok = true;
        if (w > 0) ok = ks_wait_flags(flag1, w, epoch, (uint32_t)tid, dummy);
        for (uint32_t i = tid; i < w * NE; i += THREADS) cmap[i] = __ldcg(gmap + i);
        __syncthreads();
        if (tid == 0) {
            uint32_t e = (uint32_t)(pos0 - W0), cb = 0;
            for (uint32_t k = 0; k < w && e != 0xFFFFu; k++) {
                const uint32_t m = cmap[k * NE + (e >> 1)];
                cb += m >> 16;
                e = m & 0xFFFFu;
            }
            sh_entry = e; sh_cb = cb;
            if (!ok) sh_err = MCRAW_FRAME_INTERNAL;
        }
        __syncthreads();
        first = sh_cb;
        done = first;
        uint32_t pstart = sh_entry;                            // 0xFFFF: the chain died in an earlier window
        if (pstart == 0xFFFFu) { dead = true; pstart = SENT; }

        // ---- 3. the true chain through this window, THREADS blocks per round: thread t owns blocks t, t + THREADS, ... of the
        //      window's chain and reaches them by itself (t = 64 a + 16 b + 4 c + d hops through the four tables, then
        //      THREADS / 64 long hops per round); positions behind the window and SENT are fixed points
        uint32_t mypos = pstart;
        {
            const uint32_t t = (uint32_t)tid;
            for (uint32_t k = 0; k < (t >> 6); k++) mypos = lds_u16(nxt64_s + mypos);
            for (uint32_t k = 0; k < ((t >> 4) & 3u); k++) mypos = lds_u16(nxt16_s + mypos);
            for (uint32_t k = 0; k < ((t >> 2) & 3u); k++) mypos = lds_u16(nxt4_s + mypos);
            for (uint32_t k = 0; k < (t & 3u); k++) mypos = lds_u16(nxt1_s + mypos);
        }
        while (done < need_mb && pstart < (uint32_t)C && !sh_err) {
            const uint32_t limit = min((uint32_t)THREADS, need_mb - done);
            const uint32_t my_start = mypos;
            bool valid = false, hit_end = false;
            if ((uint32_t)tid < limit) {
                const uint32_t p = my_start;
                const uint32_t e = lds_u16(nxt1_s + p);
                valid = p < (uint32_t)C && e != SENT;
                hit_end = p == SENT || (p < (uint32_t)C && e == SENT);          // a block that does not fit the frame
            }
            const uint32_t cnt = (uint32_t)__syncthreads_count(valid);         // validity is monotone along the chain
            if (__syncthreads_or(hit_end)) dead = true;
            uint32_t unit_len8 = 0;
            const uint32_t unit = done + (uint32_t)tid;
            if ((uint32_t)tid < cnt) {
                const uint32_t p = my_start;
                const uint32_t hdr = (uint32_t)stage[p] | ((uint32_t)stage[p + 1] << 8);
                uint32_t* rec = reinterpret_cast<uint32_t*>(F.metarec + unit);
                rec[stream] = (uint32_t)(base + p);
                rec[2 + stream] = hdr;
                if (stream == 0) {
                    uint32_t bad;
                    unit_len8 = bits_block_len8(stage, p, unit, ntiles, bad);
                    if (bad) sh_bad = 1;
                }
            }
            if (stream == 0) {
                uint32_t incl = unit_len8;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if ((tid & 31) >= d) incl += o;
                }
                if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
                __syncthreads();
                uint32_t wbase = 0, all = 0;
#pragma unroll
                for (int ww = 0; ww < THREADS / 32; ww++) {
                    const uint32_t sv = warp_sums[ww];
                    if (ww < (tid >> 5)) wbase += sv;
                    all += sv;
                }
                if ((uint32_t)tid < cnt) unitoff[unit] = 8u * (local_carry + wbase + incl - unit_len8);   // relative to the window for now
                local_carry += all;
                if (sh_bad) { if (tid == 0) sh_err = MCRAW_FRAME_BAD_BITS; }
            }
            done += cnt;
#pragma unroll
            for (int k = 0; k < THREADS / 64; k++) mypos = lds_u16(nxt64_s + mypos);   // this thread's block of the next round
            __syncthreads();
            if (cnt < limit) break;                                // the window (or the chain) is exhausted
        }
        __syncthreads();                                           // (every thread has read sh_err in the loop condition)
        {
            // every thread reads the flag, a barrier, then one thread may set it: the compiler is free to hoist the load of a
            // predicated read-modify-write to all threads, which racecheck (rightly) reports against thread 0's store
            const uint32_t err_now = sh_err;
            __syncthreads();
            if (tid == 0 && dead && done < need_mb && !err_now) sh_err = MCRAW_FRAME_TRUNCATED;
        }
        __syncthreads();

        KS_STAMP();
        KS_STAMP();
        // ---- 4. bits stream: absolute payload offsets = METADATA_OFFSET + the totals of the windows before + the local ones
        if (stream == 0) {
            if (tid == 0) ks_flag_store(flag2 + w, epoch, local_carry);   // (a plain number: nothing to fence)
            uint32_t mine = 0;
            bool ok2 = true;
            if (w > 0) ok2 = ks_wait_flags(flag2, w, epoch, (uint32_t)tid, mine);
            if ((uint32_t)tid >= w) mine = 0;
            // sum of the totals before this window (w <= 32 values, one per thread)
            uint32_t sum = mine;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
            if ((tid & 31) == 0) warp_sums[tid >> 5] = sum;
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int ww = 0; ww < THREADS / 32; ww++) before += warp_sums[ww];
            const uint32_t carry_abs = 16u + 8u * before;                                  // METADATA_OFFSET (RawData.cpp:25,562)
            for (uint32_t u = first + tid; u < done; u += THREADS) unitoff[u] += carry_abs;
            if (tid == 0) {
                if (!ok2) sh_err = MCRAW_FRAME_INTERNAL;
                if (done >= need_mb && first < need_mb) {                                  // the window that holds the last unit
                    const unsigned long long endoff = (unsigned long long)carry_abs + 8ull * local_carry;
                    unitoff[need_mb] = (uint32_t)endoff;
                    if (!sh_err && endoff > len) sh_err = MCRAW_FRAME_TRUNCATED;            // RawData.cpp:419
                }
            }
            __syncthreads();
        }
    }

    // ---- done: every window says so (with its error bits); the last one answers for the stream
    __syncthreads();
    KS_STAMP();
#ifdef MCRAW_KS_DEBUG
    if (tid == 0 && epoch == 3u)
        printf("ks f%d s%d w%2u first %5u done %5u | stage %6lld nxt %6lld map %6lld wait+compose %6lld chain %6lld offsets %6lld (cycles)\n", f, stream, w, first, done,
               dbg_t[1] - dbg_t[0], dbg_t[2] - dbg_t[1], dbg_t[3] - dbg_t[2], dbg_t[4] - dbg_t[3], dbg_t[5] - dbg_t[4], dbg_t[6] - dbg_t[5]);
#endif
    const uint32_t my_err = sh_err;
    if (w + 1 < nw) {
        if (tid == 0) { __threadfence(); ks_flag_store(flag3 + w, epoch, my_err); }
        return;
    }
    uint32_t e3 = 0;
    const bool ok3 = nw > 1 ? ks_wait_flags(flag3, nw - 1, epoch, (uint32_t)tid, e3) : true;
    if ((uint32_t)tid >= nw - 1) e3 = 0;
    const uint32_t any_err = (uint32_t)__syncthreads_or((int)e3);   // (non-zero if any window reported an error; the bits themselves below)
    if (tid == 0) {
        uint32_t err = my_err;
        if (any_err)
            for (uint32_t k = 0; k + 1 < nw; k++) err |= (uint32_t)ks_flag_load(flag3 + k);
        if (!ok3) err |= MCRAW_FRAME_INTERNAL;
        if (!err && done < need_mb) err = MCRAW_FRAME_TRUNCATED;  // the windows of this launch do not reach the end of the chain
        S.status[stream] = err;
    }
    meta_publish(S, (uint32_t)stream, epoch);
}

// --------------------------------------------------------------------------------------------------------
// k_units
// --------------------------------------------------------------------------------------------------------
constexpr int KU_WARPS = 4;                       // warps per CTA
#ifndef MCRAW_KU_UPW
#define MCRAW_KU_UPW 24
#endif
constexpr int KU_UPW = MCRAW_KU_UPW;              // consecutive units handled by one warp (software pipelined); <= 31: a warp keeps the unit offsets (+ end) in its lanes
constexpr int KU_IN_BYTES = 16 * 512 + 256;       // worst case unit payload (16-bit blocks) + alignment slack, 128-multiple
constexpr int KU_SLOT_PITCH = 144;                // bytes per tile slot in an output row: 128 + 16 (bank skew)
constexpr int KU_ROW_PITCH = 16 * KU_SLOT_PITCH + 64;   // 2368: (pitch/16) % 8 == 4 -> the two pair rows hit disjoint banks
constexpr int KU_OUT_BYTES = 2 * KU_ROW_PITCH;    // two rows at a time (planes 0..3, then planes 4..7)
constexpr int KU_META_BLOCK = 160;                // a staged metadata block: <= 15 bytes of alignment + 2 + 128, in 16-byte chunks
constexpr int KU_META_BYTES = 2 * 2 * KU_META_BLOCK;   // (this unit, next unit) x (bits block, refs block)
// Payload staging: ONE cp.async.bulk (TMA, 1-D) per unit, issued by lane 0 and tracked by the warp's mbarrier.  A bulk copy
// is linear, so the staged payload is not swizzled: the decode reads meet whatever bank conflicts the block lengths
// produce -- measured faster all the same (profiles/README.md: k_units 260.6 -> 251.3 us on C2).  -DMCRAW_KU_LDGSTS builds
// the round-1 staging (16-byte cp.async per lane into an XOR-swizzled buffer) for the A/B.
#ifndef MCRAW_KU_LDGSTS
#define MCRAW_KU_BULK 1
#endif
#ifdef MCRAW_KU_BULK
constexpr int KU_BAR_BYTES = 128;                 // the warp's mbarrier (8 bytes used)
#else
constexpr int KU_BAR_BYTES = 0;
#endif
constexpr int KU_WARP_SMEM = ((KU_IN_BYTES + KU_OUT_BYTES + KU_META_BYTES + KU_BAR_BYTES + 127) / 128) * 128;
constexpr int KU_SMEM = KU_WARPS * KU_WARP_SMEM;
static_assert(KU_IN_BYTES % 128 == 0 && KU_WARP_SMEM % 128 == 0, "swizzle rows are 128 bytes");
static_assert(KU_UPW >= 1 && KU_UPW <= 31, "lane i holds unitoff[u0 + i], i = 0 .. units");
static_assert((KU_ROW_PITCH / 16) % 8 == 4, "row pitch must skew pair rows by four 16-byte bank groups");

// 128-byte rows, 16-byte chunks XOR-ed with the row index: lanes reading at a 128-byte stride stay (nearly) conflict free
#ifdef MCRAW_KU_BULK
__device__ __forceinline__ uint32_t swz(uint32_t o) { return o; }
#else
__device__ __forceinline__ uint32_t swz(uint32_t o) { return o ^ ((o >> 3) & 0x70u); }
#endif

struct SwzFetch {
    uint32_t base;   // shared-window address of the staged unit (128-byte aligned)
    uint32_t a;      // logical byte offset of the block payload (8-byte aligned)
    __device__ __forceinline__ uint2 operator()(int g) const { return lds64(base + swz(a + 8u * (uint32_t)g)); }
};

// Planes 4*HALF .. 4*HALF+3 of an even/odd block pair -> 64 pixels of output row (pair row) + 2*HALF:
// interleave the two Bayer columns (RawData.cpp:581-593), widen to u16, add the references mod 2^16.
template <bool WITH_H, int HALF>
__device__ __forceinline__ void emit_half(const uint32_t (&LE)[16], const uint32_t (&HE)[16], const uint32_t (&LO)[16],
                                          const uint32_t (&HO)[16], const uint32_t refs, const uint32_t out_lane) {
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        const int j = 4 * HALF + jj;
        uint32_t w[8];
#pragma unroll
        for (int hw = 0; hw < 2; hw++) {
            const uint32_t le = LE[2 * j + hw], lo = LO[2 * j + hw];
            const uint32_t x0 = __byte_perm(le, lo, 0x5140), x1 = __byte_perm(le, lo, 0x7362);   // E0 O0 E1 O1 | E2 O2 E3 O3 (low bytes)
            if (WITH_H) {
                const uint32_t he = HE[2 * j + hw], ho = HO[2 * j + hw];
                const uint32_t y0 = __byte_perm(he, ho, 0x5140), y1 = __byte_perm(he, ho, 0x7362);
                w[4 * hw + 0] = __byte_perm(x0, y0, 0x5140);
                w[4 * hw + 1] = __byte_perm(x0, y0, 0x7362);
                w[4 * hw + 2] = __byte_perm(x1, y1, 0x5140);
                w[4 * hw + 3] = __byte_perm(x1, y1, 0x7362);
            } else {
                w[4 * hw + 0] = __byte_perm(x0, 0u, 0x4140);
                w[4 * hw + 1] = __byte_perm(x0, 0u, 0x4342);
                w[4 * hw + 2] = __byte_perm(x1, 0u, 0x4140);
                w[4 * hw + 3] = __byte_perm(x1, 0u, 0x4342);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = __vadd2(w[k], refs);                                  // mod 2^16 per sample
        sts128(out_lane + jj * 32, w[0], w[1], w[2], w[3]);
        sts128(out_lane + jj * 32 + 16, w[4], w[5], w[6], w[7]);
    }
}

// Optional epilogue on one word = two horizontally adjacent pixels (even column | odd column << 16) of a row with parity
// `rowpar`: what the reference's consumer does with the container's blackLevel[4] / whiteLevel (example.cpp:66-67,89-91 hands
// them to the DNG; a raw processor subtracts and scales).  mode 1: integers, min(max(v - black, 0), white - black);
// mode 2: IEEE half, ((float)v - black) * (1 / (white - black)) clamped to [0, 1], fp32 arithmetic, round to nearest.
// The frame's epilogue constants, fetched ONCE per work item / tile into registers: read through the FrameDev reference
// they would be loaded again after every store (the compiler cannot rule out that the stores alias them).
struct EpiRegs {
    unsigned black2[2], range2[2];
    float blackf[4], scalef[4];
};
template <class Frame>
__device__ __forceinline__ EpiRegs epilogue_regs(const Frame& F, const unsigned mode) {
    EpiRegs E;
#pragma unroll
    for (int i = 0; i < 2; i++) { E.black2[i] = 0; E.range2[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { E.blackf[i] = 0.f; E.scalef[i] = 0.f; }
    if (mode == MCRAW_OUT_BLACK_SUB) {
#pragma unroll
        for (int i = 0; i < 2; i++) { E.black2[i] = F.epi_black2[i]; E.range2[i] = F.epi_range2[i]; }
    } else if (mode != 0u) {
#pragma unroll
        for (int i = 0; i < 4; i++) { E.blackf[i] = F.epi_blackf[i]; E.scalef[i] = F.epi_scalef[i]; }
    }
    return E;
}
// ... and the constants of one row parity (a compile-time index in k_units, a select in the legacy kernel: never a
// dynamically indexed array, which would live in local memory)
struct EpiRow {
    unsigned black2, range2;
    float blackf0, blackf1, scalef0, scalef1;
};
__device__ __forceinline__ EpiRow epilogue_row(const EpiRegs& E, const bool odd) {
    EpiRow R;
    R.black2 = odd ? E.black2[1] : E.black2[0];
    R.range2 = odd ? E.range2[1] : E.range2[0];
    R.blackf0 = odd ? E.blackf[2] : E.blackf[0];
    R.blackf1 = odd ? E.blackf[3] : E.blackf[1];
    R.scalef0 = odd ? E.scalef[2] : E.scalef[0];
    R.scalef1 = odd ? E.scalef[3] : E.scalef[1];
    return R;
}
__device__ __forceinline__ uint32_t epilogue_word(const uint32_t v, const unsigned mode, const EpiRow& F) {
    if (mode == MCRAW_OUT_BLACK_SUB) {
        const uint32_t b = F.black2;
        return __vminu2(__vsub2(__vmaxu2(v, b), b), F.range2);
    }
    // u16 -> float without the conversion pipe (I2F runs at a quarter of the FMA rate and was what the half mode queued on):
    // PRMT drops the 16 bits into the mantissa of 2^23, one exact subtraction leaves float(v) -- the same value I2F gives
    const float vx = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7610)) - 8388608.0f;
    const float vy = __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7632)) - 8388608.0f;
    const float x = __saturatef((vx - F.blackf0) * F.scalef0);
    const float y = __saturatef((vy - F.blackf1) * F.scalef1);
    const __half2 h = __floats2half2_rn(x, y);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// Copy-out role of a lane: 16-byte chunk (lane & 7) of the 128-byte row segments of tile slots (lane >> 3) + 4k.
struct CopyOut {
    uint16_t* ptr[4];    // address of the lane's chunk in row 4*ty of the tile of slot group k
    uint32_t nrows[4];   // rows of that tile that may be written (0 = tile not live for this lane)
};

template <bool VEC, int HALF, unsigned MODE>   // MODE: the epilogue as a compile-time constant (one test per row pair instead of one per word)
__device__ __forceinline__ void copy_half(const CopyOut& co, const uint32_t out_base, const uint32_t lane, const int width,
                                          const EpiRegs& F) {
    constexpr unsigned epi = MODE;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t s = out_base + ((lane >> 3) + 4u * k) * KU_SLOT_PITCH + (lane & 7u) * 16u;
#pragma unroll
        for (int qq = 0; qq < 2; qq++) {
            const uint32_t r = qq + 2 * HALF;
            if (r < (VEC ? co.nrows[k] : (co.nrows[k] & 0xFFu))) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s + qq * KU_ROW_PITCH));
                if (epi) {                                   // output row 4 ty + r: its parity is qq; the chunk starts at an even column
                    const EpiRow R = epilogue_row(F, qq != 0);
                    v.x = epilogue_word(v.x, epi, R); v.y = epilogue_word(v.y, epi, R);
                    v.z = epilogue_word(v.z, epi, R); v.w = epilogue_word(v.w, epi, R);
                }
                uint16_t* orow = co.ptr[k] + (size_t)r * (size_t)width;
                if (VEC) *reinterpret_cast<uint4*>(orow) = v;
                else {
                    // width % 8 != 0 or unaligned dst: element stores, cropped at the row end
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
                    const int n = (int)(co.nrows[k] >> 8);   // pixels of this chunk inside the row (see setup)
                    for (int e = 0; e < 8; e++)
                        if (e < n) orow[e] = (uint16_t)(wv[e >> 1] >> (16 * (e & 1)));
                }
            }
        }
    }
}

template <bool VEC>
__device__ __forceinline__ void emit_and_copy(const uint32_t (&LE)[16], const uint32_t (&HE)[16], const uint32_t (&LO)[16],
                                              const uint32_t (&HO)[16], const uint32_t refs, const bool with_h, const CopyOut& co,
                                              const uint32_t out_base, const uint32_t lane, const int width, const EpiRegs& F,
                                              const unsigned epi) {
    const uint32_t out_lane = out_base + (lane & 1u) * KU_ROW_PITCH + (lane >> 1) * KU_SLOT_PITCH;
#ifdef MCRAW_KU_EXP_NOEMIT        // experiment (wrong pixels): staging and copy-out only -- the memory system's share
    sts128(out_lane, LE[0], LO[0], refs, HE[0] | HO[0]);
    __syncwarp();
    if (epi == MCRAW_OUT_RAW) { copy_half<VEC, 0, MCRAW_OUT_RAW>(co, out_base, lane, width, F); __syncwarp(); copy_half<VEC, 1, MCRAW_OUT_RAW>(co, out_base, lane, width, F); }
    __syncwarp();
    return;
#endif
    if (with_h) emit_half<true, 0>(LE, HE, LO, HO, refs, out_lane); else emit_half<false, 0>(LE, HE, LO, HO, refs, out_lane);
    __syncwarp();
    if (epi == MCRAW_OUT_RAW) copy_half<VEC, 0, MCRAW_OUT_RAW>(co, out_base, lane, width, F);
    else if (epi == MCRAW_OUT_BLACK_SUB) copy_half<VEC, 0, MCRAW_OUT_BLACK_SUB>(co, out_base, lane, width, F);
    else copy_half<VEC, 0, MCRAW_OUT_NORM_F16>(co, out_base, lane, width, F);
    __syncwarp();
    if (with_h) emit_half<true, 1>(LE, HE, LO, HO, refs, out_lane); else emit_half<false, 1>(LE, HE, LO, HO, refs, out_lane);
    __syncwarp();
    if (epi == MCRAW_OUT_RAW) copy_half<VEC, 1, MCRAW_OUT_RAW>(co, out_base, lane, width, F);
    else if (epi == MCRAW_OUT_BLACK_SUB) copy_half<VEC, 1, MCRAW_OUT_BLACK_SUB>(co, out_base, lane, width, F);
    else copy_half<VEC, 1, MCRAW_OUT_NORM_F16>(co, out_base, lane, width, F);
    __syncwarp();
}

// The pixel work of one item: the calling WARP decodes units [u0, u0 + upw) of the frame.
// The per-unit records written by k_meta are single-use: they are read with ld.global.cg (L2 only).
// smem_warp: the warp's KU_WARP_SMEM bytes.
// The two values a lane needs from a staged metadata block (samples 2*lane and 2*lane + 1: the lane's even- and
// odd-column block): (unpacked + header reference) mod 2^16 (RawData.cpp:491-492), as two 16-bit lanes of one word
// (even block | odd block << 16).  blk: shared-memory address of the staged block, whose header sits at byte (pos & 15).
// For header values <= 10 up to three byte PAIRS are picked by the table rows of the lane's plane; the 16-bit layout
// holds the samples themselves.
__device__ __forceinline__ uint32_t meta_values(const uint32_t blk, const uint32_t pos, const uint32_t hdr, const uint32_t* s_terms,
                                                const uint32_t lane) {
    const uint32_t hb = (hdr >> 4) & 15u;
    const uint32_t ref = ((hdr & 15u) << 8) | ((hdr >> 8) & 0xFFu);                // RawData.cpp:106-110
    const uint32_t pay = blk + (pos & 15u) + 2u;
    // a stream that starts at an odd offset (never written by an encoder, accepted by the reference) puts every byte pair
    // at an odd address: two byte loads instead of one 16-bit load, decided once per block (warp-uniform)
    const bool odd = (pos & 1u) != 0u;
    auto pair_at = [&](const uint32_t a) { return odd ? (lds_u8(a) | (lds_u8(a + 1u) << 8)) : lds_u16(a); };
    uint32_t v = 0;
    if (hb > 10u) {                                                               // RawData.cpp:376-408
        v = pair_at(pay + 4u * lane) | (pair_at(pay + 4u * lane + 2u) << 16);
    } else {
        const uint32_t* row = s_terms + (hb * 8u + (lane >> 2)) * 3u;
        const uint32_t b = 2u * (lane & 3u);
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const uint32_t term = row[t];
            if (term >> 16) v |= mcraw_meta_term_pair(term, pair_at(pay + 8u * mcraw_meta_term_group(term) + b));
        }
    }
    return __vadd2(v, ref | (ref << 16));
}

// what the pixel kernel needs of a frame's descriptor (read through L2, see plan_wait)
struct FramePix {
    const uint8_t* src;
    unsigned long long len;
    uint16_t* dst;
    int width;
    unsigned tiles_x, flags, inv_tiles_x, epi_mode;
    uint32_t* unitoff;
    uint4* metarec;
    unsigned epi_black2[2], epi_range2[2];
    float epi_blackf[4], epi_scalef[4];
};
template <bool EPI>
__device__ __forceinline__ FramePix frame_pix_cg(const FrameDev* p) {
    FramePix f;
    f.src = ldcg_ptr(&p->src); f.len = __ldcg(&p->len); f.dst = ldcg_ptr(&p->dst);
    f.width = __ldcg(&p->width); f.tiles_x = __ldcg(&p->tiles_x); f.flags = __ldcg(&p->flags); f.inv_tiles_x = __ldcg(&p->inv_tiles_x);
    f.unitoff = ldcg_ptr(&p->unitoff); f.metarec = ldcg_ptr(&p->metarec);
    f.epi_mode = EPI ? __ldcg(&p->epi_mode) : 0u;
#pragma unroll
    for (int i = 0; i < 2; i++) { f.epi_black2[i] = 0; f.epi_range2[i] = 0; }
#pragma unroll
    for (int i = 0; i < 4; i++) { f.epi_blackf[i] = 0.f; f.epi_scalef[i] = 0.f; }
    if (EPI && f.epi_mode != 0u) {
#pragma unroll
        for (int i = 0; i < 2; i++) { f.epi_black2[i] = __ldcg(&p->epi_black2[i]); f.epi_range2[i] = __ldcg(&p->epi_range2[i]); }
#pragma unroll
        for (int i = 0; i < 4; i++) { f.epi_blackf[i] = __ldcg(&p->epi_blackf[i]); f.epi_scalef[i] = __ldcg(&p->epi_scalef[i]); }
    }
    return f;
}

template <bool EPI>
__device__ __forceinline__ void units_task(const FrameDev* Fd, const FrameState& S, Result* __restrict__ result,
                                           const uint32_t u0, const uint32_t upw, uint8_t* smem_warp, const uint32_t* s_terms,
                                           uint32_t& bulk_phase) {
    const uint32_t lane = threadIdx.x & 31;
    const unsigned status = __ldcg(&S.status[0]) | __ldcg(&S.status[1]);
    const uint32_t rows_fit = __ldcg(&S.rows_fit);
    const FramePix F = frame_pix_cg<EPI>(Fd);
    const int width = F.width;
    if (u0 == 0 && lane == 0) {
        Result r;
        r.written = status ? 0ull : (unsigned long long)rows_fit * (unsigned long long)width;      // RawData.cpp:611
        r.status = status;
        r.pad = 0;
        *result = r;
    }
    if (status) return;
    const uint32_t tiles_x = F.tiles_x;
    const uint32_t ntiles = tiles_x * __ldcg(&S.tile_rows_dev);
    const uint32_t nunits = (ntiles + 15u) / 16u;
    if (u0 >= nunits) return;
    const uint32_t nu = min(upw, nunits - u0);

    const uint32_t in_base = smem_u32(smem_warp);
    const uint32_t out_base = in_base + KU_IN_BYTES;
    const uint8_t* __restrict__ src = F.src;
    const unsigned long long len = F.len;
    const uint32_t inv = F.inv_tiles_x;
    const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
    const unsigned epi = EPI ? F.epi_mode : 0u;         // EPI = false: the epilogue code is not even in the kernel
    const EpiRegs ER = epilogue_regs(F, epi);
    uint16_t* __restrict__ dst = F.dst;

    // payload offsets of this warp's units (+ end): lane i holds unitoff[u0 + i]; and their metadata records
    const uint32_t my_off = (lane <= nu) ? __ldcg(F.unitoff + u0 + lane) : 0u;
    uint4 my_rec = make_uint4(0, 0, 0, 0);
    if (lane < nu) my_rec = __ldcg(F.metarec + u0 + lane);
    auto rec_of = [&](uint32_t i) {
        return make_uint4(__shfl_sync(0xFFFFFFFFu, my_rec.x, i), __shfl_sync(0xFFFFFFFFu, my_rec.y, i),
                          __shfl_sync(0xFFFFFFFFu, my_rec.z, i), __shfl_sync(0xFFFFFFFFu, my_rec.w, i));
    };
    // stage the two metadata blocks of a unit (16-byte chunks around them; nothing is read past the end of the buffer):
    // lanes 0..9 the bits block, lanes 16..25 the refs block
    const uint32_t meta_base = out_base + KU_OUT_BYTES;
    auto stage_meta = [&](const uint4 r, uint32_t slot) {
        const uint32_t k = lane & 15u;
        if (k < (uint32_t)(KU_META_BLOCK / 16)) {
            const uint32_t pos = lane < 16u ? r.x : r.y;
            const unsigned long long a = (unsigned long long)(pos & ~15u) + 16ull * k;
            const uint32_t n = a >= len ? 0u : (uint32_t)min(16ull, len - a);
            const uint32_t d = meta_base + (slot * 2u + (lane >> 4)) * KU_META_BLOCK + 16u * k;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src + (n ? a : 0ull)), "r"(n) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

#ifdef MCRAW_KU_BULK
    const uint32_t bar = meta_base + KU_META_BYTES;
#endif
    // stage unit payload [a0, a1) (rounded out to 16 bytes)
    auto stage_unit = [&](uint32_t a0, uint32_t a1) {
        const uint32_t s0 = a0 & ~15u;
        const uint32_t nchunks = (a1 - s0 + 15u) >> 4;                            // <= (8192 + 8 + 15) / 16
        if ((unsigned long long)s0 + 16ull * nchunks <= len) {
#ifdef MCRAW_KU_BULK
            if (nchunks == 0) {
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
            } else if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(16u * nchunks) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                             ::"r"(in_base), "l"(src + s0), "r"(16u * nchunks), "r"(bar) : "memory");
            }
            return;
#else
            const uint8_t* g = src + s0 + 16u * lane;
            for (uint32_t c = lane; c < nchunks; c += 32, g += 512) cp_async16(in_base + swz(16u * c), g);
#endif
        } else {                                                                   // tail of the buffer: bytes, zero filled
            for (uint32_t c = lane; c < nchunks; c += 32) {
                const unsigned long long o = (unsigned long long)s0 + 16ull * c;
                uint32_t t4[4] = {0, 0, 0, 0};
                for (int k = 0; k < 16; k++)
                    if (o + k < len) t4[k >> 2] |= (uint32_t)src[o + k] << (8 * (k & 3));
                sts128(in_base + swz(16u * c), t4[0], t4[1], t4[2], t4[3]);
            }
#ifdef MCRAW_KU_BULK
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");          // the next bulk copy overwrites these generic stores
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
            return;
#endif
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    uint32_t a0 = __shfl_sync(0xFFFFFFFFu, my_off, 0), a1 = __shfl_sync(0xFFFFFFFFu, my_off, 1);
    uint4 rec = rec_of(0);
    stage_meta(rec, 0);           // committed BEFORE the payload: wait_group 1 then means "the metadata is in"
    stage_unit(a0, a1);

    for (uint32_t i = 0; i < nu; i++) {
        const uint32_t unit = u0 + i;
        // ---- this lane's block pair: header bits values, references, payload offset inside the unit
#ifdef MCRAW_KU_BULK
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");                    // only the metadata travels in cp.async groups
#else
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
#endif
        __syncwarp();
        const uint32_t mb = meta_base + (i & 1u) * 2u * KU_META_BLOCK;
        const uint32_t vb = meta_values(mb, rec.x, rec.z, s_terms, lane);
        const uint32_t refs = meta_values(mb + KU_META_BLOCK, rec.y, rec.w, s_terms, lane);
        const bool live = unit * 16u + (lane >> 1) < ntiles;                        // padding values of the last unit are ignored
        const uint32_t bE = live ? vb & 0xFFu : 0u, bO = live ? (vb >> 16) & 0xFFu : 0u;   // k_meta has checked them: <= 16
        const uint32_t len8 = cur_len8(bE) + cur_len8(bO);
        uint32_t rel8 = len8;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, rel8, d);
            if (lane >= (uint32_t)d) rel8 += o;
        }
        const uint32_t aE = (a0 & 15u) + 8u * (rel8 - len8);

        // copy-out bookkeeping for this unit (independent of the staged data)
        CopyOut co;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t tile = unit * 16u + (lane >> 3) + 4u * k;
            const uint32_t ty = inv ? __umulhi(tile, inv) : tile / tiles_x;
            const uint32_t tx = tile - ty * tiles_x;
            const int xpix = (int)(64u * tx + 8u * (lane & 7u));
            uint32_t nr = 0;
            if (tile < ntiles && xpix < width && 4u * ty < rows_fit) nr = min(4u, rows_fit - 4u * ty);
            if (!vec) nr |= (uint32_t)min(8, max(0, width - xpix)) << 8;
            co.nrows[k] = nr;
            co.ptr[k] = dst + (size_t)(4u * ty) * (size_t)width + xpix;
        }

#ifdef MCRAW_KU_BULK
        {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(bar), "r"(bulk_phase & 1u) : "memory");
            }
            bulk_phase++;
        }
#else
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
        __syncwarp();

        // ---- decode the even-column and the odd-column block of the pair
        uint32_t LE[16], HE[16], LO[16], HO[16];
#ifdef MCRAW_KU_EXP_NODECODE      // experiment (wrong pixels): what the kernel costs without the block decode
#pragma unroll
        for (int q = 0; q < 16; q++) { LE[q] = aE + q; LO[q] = bE + q; HE[q] = 0; HO[q] = 0; }
        { const uint2 t = lds64(in_base + 8u * lane); LE[0] ^= t.x; LO[0] ^= t.y; }
#else
        const uint32_t lenE = decode_block(bE, SwzFetch{in_base, aE}, LE, HE);
        decode_block(bO, SwzFetch{in_base, aE + lenE}, LO, HO);
#endif
        __syncwarp();

        // ---- the input buffer is free again: start fetching the next unit behind the emit / copy-out phases
        if (i + 1 < nu) {
            rec = rec_of(i + 1);
            stage_meta(rec, (i + 1u) & 1u);
            a0 = a1;
            a1 = __shfl_sync(0xFFFFFFFFu, my_off, i + 2);
            stage_unit(a0, a1);
        }

        const bool with_h = __any_sync(0xFFFFFFFFu, (bE > 8u) | (bO > 8u));
        if (vec) emit_and_copy<true>(LE, HE, LO, HO, refs, with_h, co, out_base, lane, width, ER, epi);
        else emit_and_copy<false>(LE, HE, LO, HO, refs, with_h, co, out_base, lane, width, ER, epi);
    }
}

// --------------------------------------------------------------------------------------------------------
// k_units: the pixel work of the whole batch in one persistent launch (grid = resident CTAs, or fewer for small
// batches).  Every WARP takes items (frame, first unit, units) from a queue in global memory, so there is no wave
// quantisation: the host ends the item list with progressively smaller items (build_items()) and warps never wait for
// each other.  Runs after k_meta on the same stream.
// block = KD_THREADS, dynamic smem = KU_SMEM
// --------------------------------------------------------------------------------------------------------
__constant__ uint32_t c_meta_terms[MCRAW_META_ROWS][8][3] = MCRAW_META_TERMS_INIT;

struct WorkItem {
    uint32_t frame;
    uint32_t what;     // bits 0..26 = first unit, bits 27..31 = units - 1 (<= 31: a warp keeps the unit offsets in its lanes)
};
constexpr int KD_THREADS = 32 * KU_WARPS;

// counters[0]: next item; counters[1]: warps that have left the kernel (the last one resets both for the next launch).
// flag_target != 0: launched as a programmatic dependent of k_meta, i.e. possibly while k_meta's last wave is still
// running -- before touching a frame, lane 0 polls the frame's two done words for this launch's epoch (relaxed loads served
// by L2) and closes the wait with one acquire load; the other lanes wait at the warp barrier.  The per-unit records are then read with
// ld.global.cg (L2), so they see everything k_meta released before bumping the counter.
// Three CTAs per SM (168 registers): the fourth buys 2 % of bandwidth and leaves no room for this logic without spills.
#ifndef MCRAW_KU_EPI_CTAS
#define MCRAW_KU_EPI_CTAS 0      // experiment: CTAs per SM the epilogue variant of k_units is compiled for (0 = like the raw one)
#endif
template <bool EPI>
__global__ void __launch_bounds__(KD_THREADS, (EPI && MCRAW_KU_EPI_CTAS) ? MCRAW_KU_EPI_CTAS : 3)
k_units(const FrameDev* frames, const FrameState* __restrict__ states, Result* __restrict__ results,
        const WorkItem* items, const uint32_t nitems, uint32_t* __restrict__ counters, const uint32_t flag_target,
        const uint32_t* plan_ready, const uint32_t plan_epoch) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint32_t s_terms[MCRAW_META_ROWS * 8 * 3];
    // The NEXT batch's k_meta may be launched as a programmatic dependent of this kernel (mcraw_capi.cu, "chain"): it touches
    // nothing this kernel uses (its own slot's scratch, its own frames), so it may start as soon as it finds room.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    for (int i = threadIdx.x; i < MCRAW_META_ROWS * 8 * 3; i += KD_THREADS) s_terms[i] = (&c_meta_terms[0][0][0])[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    uint8_t* smem_warp = smem_raw + (threadIdx.x >> 5) * KU_WARP_SMEM;
    uint32_t bulk_phase = 0;
#ifdef MCRAW_KU_BULK
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(smem_warp) + KU_IN_BYTES + KU_OUT_BYTES + KU_META_BYTES) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
#endif
    uint32_t it = 0;
    if (lane == 0) {
        plan_wait(plan_ready, plan_epoch);                 // the work list and the descriptors are in place
        it = atomicAdd(&counters[0], 1u);
    }
    __syncwarp();
    it = __shfl_sync(0xFFFFFFFFu, it, 0);
    while (it < nitems) {
        WorkItem w;
        {
            const unsigned long long raw = __ldcg(reinterpret_cast<const unsigned long long*>(items + it));
            w.frame = (uint32_t)raw; w.what = (uint32_t)(raw >> 32);
        }
        uint32_t nxt = 0;
        if (lane == 0) {
            nxt = atomicAdd(&counters[0], 1u);             // take the next item early
            if (flag_target) {
                const unsigned* flag = &states[w.frame].done[0];   // both streams' words as one 64-bit load
                unsigned long long v;
                const unsigned long long want = ((unsigned long long)flag_target << 32) | flag_target;
                for (;;) {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(flag) : "memory");
                    if (v == want) break;
                    __nanosleep(256);
                }
                // a word that carries this launch's epoch stays: one acquire load now pairs with the index kernel's releases (relaxed
                // polls keep the L1 invalidations of an acquire out of the wait loop); the warp barrier below extends the order to the other lanes
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(flag) : "memory");
            }
        }
        __syncwarp();
        units_task<EPI>(frames + w.frame, states[w.frame], results + w.frame, w.what & 0x07FFFFFFu, ((w.what >> 27) & 31u) + 1u, smem_warp, s_terms,
                   bulk_phase);
        __syncwarp();
        it = __shfl_sync(0xFFFFFFFFu, nxt, 0);
    }
    if (lane == 0 && atomicAdd(&counters[1], 1u) == gridDim.x * KU_WARPS - 1u) { counters[0] = 0; counters[1] = 0; }
}

}  // namespace mcraw
