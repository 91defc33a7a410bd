// mcraw_legacy.cuh -- sm_100a kernel for the legacy frame format (compressionType 6).
//
// Reference: /root/reference/lib/RawData_Legacy.cpp:445-495 (raw::DecodeLegacy) and :372-442 (DecodeHeader /
// DecodeBlock).  A frame is one chain of 16-sample blocks, each with an inline 2-byte header
// (bits nibble, 12-bit reference) followed by 2*bits payload bytes (32 for nibbles 11..15): block k+1 starts
// where block k ends, so the reference finds the blocks with 750 000 dependent steps per 4000x3000 frame.
//
// Here the chain is resolved in parallel, in ONE pass over the stream.  Two facts carry it:
//   (1) all block lengths are even and <= 34 bytes, so the chain enters any fixed byte range at one of 17 even offsets;
//   (2) two chains that ever share a block start are identical from there on, and with blocks of varying length chains
//       started anywhere MERGE within a few blocks.  Nothing relies on (2) for correctness -- only for speed.
//
// k_legacy_warp: one WARP per CTA, up to 32 CTAs per SM, persistent.  Warps take (frame, tile) tickets in a host-built
// order (tile index major, frame minor: neighbouring tickets belong to different frames, so every frame's chain only has
// to advance a few tiles per generation of warps).  A tile is 32 segments, one per lane.  Per tile:
//   1. stage the tile (+ overrun) with ONE bulk copy (cp.async.bulk, TMA 1-D, mbarrier transaction count) -- the only
//      time the stream is read;
//   2. chain C0 (tile entry offset 0): every lane walks a GUESSED chain from the start of its segment and marks its block
//      starts in a bitmap; then every lane re-walks from where its left neighbour's chain actually ends, only until it
//      meets its own marks (self-synchronisation), repeated until no entry changes.  The other 16 possible entry offsets
//      walk until they meet C0: the tile's transfer map  entry -> (exit offset, block count);
//   3. publish the map (LOCAL), then DECOUPLED LOOK-BACK over the previous tiles of the frame: the nearest predecessor
//      whose inclusive state (exit offset, blocks so far) is known, composed with the maps of the tiles in between
//      (a window of status words per poll; one lane chases the concrete entry state through the staged maps);
//      publish this tile's inclusive state (INCL) right away, so that successors can go on;
//   4. patch the bitmap for the true entry (the few blocks before the merge point; a chain that never meets C0 --
//      blocks of one constant width -- is re-walked from its entry);
//   5. a prefix popcount gives every lane the ordinal of the first block start in ITS segment; it decodes the block PAIRS
//      led from there (even-column block + odd-column block, :480-481) straight from shared memory: MSB-first bit
//      extraction with rotates (:38-370), + reference mod 2^16, column interleave (:483-486) in registers, 16-byte
//      stores, crop at width (:490).
// No CTA-wide barrier anywhere: a warp never waits for another warp of its SM, only (bounded) for the status words of
// earlier tickets, whose owners are resident and never wait for a later ticket -- so the waits always end; they are
// bounded all the same (MCRAW_FRAME_INTERNAL instead of a hang).  Status words carry the launch epoch of the slot,
// so nothing has to be zeroed between launches.
#pragma once
#include "mcraw_kernels.cuh"

namespace mcraw {

#ifndef MCRAW_LGW_WARPS
#define MCRAW_LGW_WARPS 4
#endif
constexpr int LGW_WARPS = MCRAW_LGW_WARPS;     // warps per CTA
constexpr int LGW_THREADS = 32 * LGW_WARPS;
constexpr int LGW_SEG = 124;                   // bytes of the tile a thread owns: 62 candidate block starts, two words of marks.
                                               // NOT a multiple of 128: threads at the same offset of their segments hit 32
                                               // different shared-memory banks (the walks are byte loads at thread * LGW_SEG + d)
constexpr int LGW_TILE = LGW_THREADS * LGW_SEG;   // bytes per tile (one CTA at a time): 15 872 for four warps
constexpr int LGW_PRE = LGW_SEG;               // bytes a thread's guessed chain runs in front of its segment to fall in step
constexpr int LGW_PAIR_CHUNK = 1024;           // pairs listed and decoded per pass (a tile holds ~700 for typical images,
                                               // up to LGW_TILE / 4 when every block is 2 bytes: then several passes)
constexpr int LG_STATES = 17;                  // entry offsets 0, 2, ..., 32
constexpr uint32_t LG_DEAD = 31;               // exit code of a chain that ran into the end of the buffer
constexpr uint32_t LG_NO_MERGE = 0xFFFFu;      // merge point of an entry whose chain never meets C0 inside the tile
constexpr int LG_OVERRUN = 80;                 // a pair led inside the tile ends at most 2 + 34 + 34 bytes past it (+ word reads)
constexpr int LGW_DATA = LGW_TILE + LG_OVERRUN;
constexpr int LGW_LB = 32;                     // look-back window: status words read per poll
constexpr int LGW_SMEM = LGW_DATA + LGW_THREADS * 8 /*marks*/ + LGW_LB * LG_STATES * 4 + LGW_THREADS * 4 /*exits, prefix*/ + LGW_PAIR_CHUNK * 2;
constexpr uint32_t LGW_ST_LOCAL = 1u, LGW_ST_INCL = 2u;
constexpr uint32_t LGW_ERR_BIT = 1u << 5;      // sticky: a wait gave up somewhere up the chain
constexpr uint32_t LGW_SPIN_LIMIT = 1u << 18;  // polls of ~0.2 us: a wait that long means something is broken, not slow
constexpr uint32_t LGW_NONE = 0xFFFFFFFFu;     // "no chain arrives here" (it ended at an undecodable block)
static_assert(LGW_DATA % 16 == 0 && LGW_TILE % 16 == 0, "bulk copies work in 16-byte granules");
static_assert(LGW_SEG % 2 == 0 && LGW_SEG / 2 <= 64 && LGW_SEG >= 34, "two words of marks per segment; a block never skips a segment");

struct LgWork { uint32_t frame, tile; };

// payload bytes of a 16-sample block for header nibble b (RawData_Legacy.cpp:13-32, min(16, bits) at :395)
__device__ __forceinline__ uint32_t leg_len(uint32_t b) { return b <= 10u ? 2u * b : 32u; }
// total length of the block whose header byte is hb
__device__ __forceinline__ uint32_t leg_step(uint32_t hb) {
    const uint32_t b = hb >> 4;
    return 2u + (b > 10u ? 32u : 2u * b);
}
// The 2-byte header at byte offset o (even) of the staged tile, as the low 16 bits of the result (byte o first).
__device__ __forceinline__ uint32_t leg_header(const uint8_t* data, uint32_t o) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(data);
    const uint32_t i0 = o >> 2;
    return (o & 2u) ? w[i0] >> 16 : w[i0];
}
__device__ __forceinline__ uint32_t leg_hdr_bits(uint32_t h) { return (h >> 4) & 15u; }                          // RawData_Legacy.cpp:372-375
__device__ __forceinline__ uint32_t leg_hdr_ref(uint32_t h) { return ((h & 15u) << 8) | ((h >> 8) & 0xFFu); }

// stage [tile_off, tile_off + nbytes) of the frame into shared memory, zero past len (16-byte granules): the tail of a buffer
__device__ __forceinline__ void lg_stage_tail(uint8_t* sm, const uint8_t* __restrict__ src, unsigned long long len,
                                              unsigned long long tile_off, int nbytes, int tid) {
    for (int v = tid; v < nbytes / 16; v += LGW_THREADS) {
        const unsigned long long o = tile_off + 16ull * (unsigned)v;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (o + 16 <= len) q = __ldg(reinterpret_cast<const uint4*>(src + o));
        else if (o < len) {
            uint32_t t4[4] = {0, 0, 0, 0};
            for (int k = 0; k < 16; k++)
                if (o + k < len) t4[k >> 2] |= (uint32_t)src[o + k] << (8 * (k & 3));
            q = make_uint4(t4[0], t4[1], t4[2], t4[3]);
        }
        *reinterpret_cast<uint4*>(sm + 16 * v) = q;
    }
}

// status word of a tile: blocks up to the end of the tile << 32 | epoch (24 bits) << 8 | state << 6 | error << 5 | exit offset / 2
__device__ __forceinline__ unsigned long long lgw_pack(uint32_t count, uint32_t epoch, uint32_t st, uint32_t low6) {
    return ((unsigned long long)count << 32) | ((unsigned long long)(epoch & 0xFFFFFFu) << 8) | (st << 6) | low6;
}
__device__ __forceinline__ void lgw_store_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void lgw_store_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lgw_load_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void lgw_cta_sync() {
    if (LGW_WARPS == 1) __syncwarp(); else __syncthreads();
}
__device__ __forceinline__ bool lgw_cta_any(const bool v) {
    if (LGW_WARPS == 1) return __any_sync(0xFFFFFFFFu, v) != 0;
    return __syncthreads_or(v ? 1 : 0) != 0;
}

// One thread, its segment [seg0, seg0 + LGW_SEG) of the staged tile: walk from tile-relative byte offset p to the end of the
// segment.  Block starts go into (m0, m1): one bit per even offset.  With MERGE the walk stops at the first position
// already marked -- from there on the old marks are this chain's own -- and older marks before it are dropped.
// rel: bytes from the tile start to the end of the buffer (a block is decoded only if it ends before the last byte,
// RawData_Legacy.cpp:387,398).  Returns true if the chain ended at an undecodable block.
template <bool MERGE>
__device__ __forceinline__ bool lgw_walk_segment(const uint8_t* data, const uint32_t seg0, uint32_t& p, uint32_t& m0, uint32_t& m1, const uint32_t rel) {
    const uint32_t old0 = m0, old1 = m1;
    uint32_t a0 = 0, a1 = 0;
    bool dead = false;
    const uint32_t end = seg0 + (uint32_t)LGW_SEG;
    while (p < end) {
        const uint32_t pos = (p - seg0) >> 1;                         // 0 .. 63
        const uint32_t bit = 1u << (pos & 31u);
        const bool hi = pos >= 32u;
        if (MERGE && ((hi ? old1 : old0) & bit)) {                    // met the old chain: its marks from here on stay
            if (hi) a1 |= old1 & ~(bit - 1u);
            else { a0 |= old0 & ~(bit - 1u); a1 = old1; }
            break;
        }
        const uint32_t q = p + leg_step(data[p]);
        if (q >= rel) { dead = true; break; }
        if (hi) a1 |= bit; else a0 |= bit;
        p = q;
    }
    m0 = a0; m1 = a1;
    return dead;
}

// CTA-wide: the exact chain that enters the tile at byte offset entry0 (even, <= 32), as a bitmap of block starts in
// shared memory (thread t owns words 2t, 2t+1).  Thread 0 knows where its chain starts.  The others guess -- but not
// blindly: a chain started anywhere falls in step with the true one within a few blocks, so the guess runs LGW_PRE bytes
// in front of the segment first (no marks, five instructions a block) and nearly always enters the segment where the
// true chain does.  Then every thread takes its left neighbour's real exit as its entry: if that is what it guessed,
// nothing is left to do; otherwise it re-walks only until it meets its own marks (self-synchronisation).  Repeated
// until no entry changes.  Returns (all threads) the chain's exit from the tile -- byte offset into the next tile, or
// LGW_NONE if the chain died -- and its block count in `total`.
__device__ __forceinline__ uint32_t lgw_chain(const uint8_t* data, uint32_t* bitmap, uint32_t* sh_exit, uint32_t* sh_sums,
                                              const uint32_t entry0, const uint32_t tile_rel, const uint32_t tid, uint32_t& total) {
    const uint32_t seg0 = tid * (uint32_t)LGW_SEG;
    uint32_t p = entry0;
    if (tid != 0) {
        p = seg0 > (uint32_t)LGW_PRE + entry0 ? seg0 - (uint32_t)LGW_PRE : entry0;
        while (p < seg0) p += leg_step(data[p]);
    }
    uint32_t entry = p, m0 = 0, m1 = 0;
    bool dead = lgw_walk_segment<false>(data, seg0, p, m0, m1, tile_rel);
    uint32_t exitv = dead ? LGW_NONE : p;                // tile-relative position in the next segment (steps are <= 34 bytes)
    for (;;) {
        sh_exit[tid] = exitv;
        lgw_cta_sync();
        const uint32_t e = tid == 0 ? entry0 : sh_exit[tid - 1];
        const bool upd = e != entry;
        if (upd) {
            entry = e;
            if (e == LGW_NONE) { m0 = 0; m1 = 0; exitv = LGW_NONE; }
            else {
                p = e;
                dead = lgw_walk_segment<true>(data, seg0, p, m0, m1, tile_rel);
                if (dead) exitv = LGW_NONE;
                else if (p >= seg0 + (uint32_t)LGW_SEG) exitv = p;     // walked to the end without meeting the old chain
                // else: merged -> the old exit stands
            }
        }
        if (!lgw_cta_any(upd)) break;                    // (a CTA barrier: every read of sh_exit above is done)
    }
    bitmap[2 * tid] = m0;
    bitmap[2 * tid + 1] = m1;
    // exclusive prefix of the block counts per segment (kept in sh_exit for lgw_marks_before) and the total
    const uint32_t c = __popc(m0) + __popc(m1);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((tid & 31u) >= (uint32_t)d) incl += o;
    }
    const uint32_t ex = sh_exit[LGW_THREADS - 1];
    uint32_t before = 0, all = incl;
    if (LGW_WARPS > 1) {
        if ((tid & 31u) == 31u) sh_sums[tid >> 5] = incl;
        __syncthreads();
        all = 0;
#pragma unroll
        for (int w = 0; w < LGW_WARPS; w++) {
            const uint32_t v = sh_sums[w];
            if ((uint32_t)w < (tid >> 5)) before += v;
            all += v;
        }
    } else {
        all = __shfl_sync(0xFFFFFFFFu, incl, 31);
        __syncwarp();                                    // every lane has read the last exit
    }
    sh_exit[tid] = before + incl - c;
    total = all;
    lgw_cta_sync();
    return ex == LGW_NONE ? LGW_NONE : ex - (uint32_t)LGW_TILE;
}

// marks are kept per segment: words 2s, 2s+1 hold the block starts of segment s, one bit per even offset from its start
__device__ __forceinline__ bool lgw_marked(const uint32_t* bitmap, const uint32_t q) {
    const uint32_t seg = q / (uint32_t)LGW_SEG, pos = (q - seg * (uint32_t)LGW_SEG) >> 1;
    return (bitmap[2u * seg + (pos >> 5)] >> (pos & 31u)) & 1u;
}
// marked block starts in front of tile-relative position q (prefix: lgw_chain's per-segment counts)
__device__ __forceinline__ uint32_t lgw_marks_before(const uint32_t* bitmap, const uint32_t* prefix, const uint32_t q) {
    const uint32_t seg = q / (uint32_t)LGW_SEG, pos = (q - seg * (uint32_t)LGW_SEG) >> 1;
    uint32_t n = prefix[seg];
    if (pos >= 32u) n += __popc(bitmap[2u * seg]) + __popc(bitmap[2u * seg + 1] & ((1u << (pos - 32u)) - 1u));
    else n += __popc(bitmap[2u * seg] & ((1u << pos) - 1u));
    return n;
}

// OR the 16 samples of the block at byte offset o (header nibble `bits`) into px[], at bit ADJ of each word (0: even-column
// block, 16: odd-column block, RawData_Legacy.cpp:483-486).  d32: the staged tile as words.  The payload is a contiguous
// MSB-first bit stream (:38-358), so 8 samples are exactly `bits` bytes: per group of 8 the bytes are fetched as
// big-endian words (PRMT with a runtime selector does alignment and byte order in one go) and every sample is one
// rotate + one mask.  Three lane-uniform formulations instead of one code path per width: widths 0..8 (two 4-sample
// windows per group), 9..10 (four 2-sample windows), and 16-bit big-endian samples (:360-370, nibbles 11..15, :395).
template <int ADJ>
__device__ __forceinline__ void lgw_block(const uint32_t* __restrict__ d32, const uint32_t o, const uint32_t bits, uint32_t (&px)[16]) {
    const uint32_t a = o + 2u;
    if (bits <= 8u) {
        const uint32_t w = bits;
        const uint32_t mask = ((1u << w) - 1u) << ADJ;
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r[j] = (32u - ADJ - (uint32_t)(j + 1) * w) & 31u;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t ag = a + g * w;
            const uint32_t i = ag >> 2, sel = 0x0123u + 0x1111u * (ag & 3u);
            const uint32_t W0 = d32[i], W1 = d32[i + 1], W2 = d32[i + 2];
            const uint32_t G0 = __byte_perm(W0, W1, sel), G1 = __byte_perm(W1, W2, sel);
            const uint32_t A1 = __funnelshift_lc(G1, G0, 4u * w);                      // the window of samples 4..7
#pragma unroll
            for (int j = 0; j < 4; j++) {
                px[8 * g + j] |= __funnelshift_r(G0, G0, r[j]) & mask;
                px[8 * g + 4 + j] |= __funnelshift_r(A1, A1, r[j]) & mask;
            }
        }
    } else if (bits <= 10u) {
        const uint32_t w = bits;
        const uint32_t mask = ((1u << w) - 1u) << ADJ;
        const uint32_t r0 = (32u - ADJ - w) & 31u, r1 = (32u - ADJ - 2u * w) & 31u;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t ag = a + g * w;
            const uint32_t i = ag >> 2, sel = 0x0123u + 0x1111u * (ag & 3u);
            const uint32_t W0 = d32[i], W1 = d32[i + 1], W2 = d32[i + 2], W3 = d32[i + 3];
            const uint32_t G0 = __byte_perm(W0, W1, sel), G1 = __byte_perm(W1, W2, sel), G2 = __byte_perm(W2, W3, sel);
            uint32_t win[4];
            win[0] = G0;                                                               // samples 2m, 2m+1 start at bit 2mw
            win[1] = __funnelshift_l(G1, G0, 2u * w);
            win[2] = __funnelshift_l(G2, G1, 4u * w - 32u);
            win[3] = __funnelshift_l(G2, G1, 6u * w - 32u);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                px[8 * g + 2 * m] |= __funnelshift_r(win[m], win[m], r0) & mask;
                px[8 * g + 2 * m + 1] |= __funnelshift_r(win[m], win[m], r1) & mask;
            }
        }
    } else {
        const uint32_t i = a >> 2, k0 = a & 3u;                                        // the payload starts 2-byte aligned
        // sample k = bytes (2k, 2k+1), big-endian: byte index (k0 + 2 (k & 1)) of the word pair (k / 2, k / 2 + 1);
        // the selector puts (high byte, low byte) at result bytes (1, 0) for ADJ 0 and (3, 2) for ADJ 16
        const uint32_t selA = ((k0 + 1u) | (k0 << 4)) << (ADJ / 2), selB = ((k0 + 3u) | ((k0 + 2u) << 4)) << (ADJ / 2);
        uint32_t lo = d32[i];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const uint32_t hi = d32[i + m + 1];
            px[2 * m] |= __byte_perm(lo, hi, selA) & (0xFFFFu << ADJ);
            px[2 * m + 1] |= __byte_perm(lo, hi, selB) & (0xFFFFu << ADJ);
            lo = hi;
        }
    }
}

template <bool EPI>
__global__ void __launch_bounds__(LGW_THREADS) k_legacy_warp(const FrameDev* __restrict__ frames, Result* __restrict__ results,
                                                             const LgWork* __restrict__ work, const uint32_t nwork,
                                                             uint32_t* __restrict__ counters, const uint32_t epoch) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    uint8_t* data = lg_smem;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LGW_DATA);                     // marks: two words per segment
    uint32_t* lbmaps = bitmap + 2 * LGW_THREADS;                                            // look-back: [LGW_LB][LG_STATES]
    uint32_t* sh_exit = lbmaps + LGW_LB * LG_STATES;                                        // chain walk: exits, then the count prefix
    uint16_t* plist = reinterpret_cast<uint16_t*>(sh_exit + LGW_THREADS);                   // pair list: [LGW_PAIR_CHUNK]
    __shared__ __align__(8) unsigned long long bar_storage;
    __shared__ uint32_t sh_map[LG_STATES], sh_merge[LG_STATES], warp_sums[LGW_WARPS];
    __shared__ uint32_t sh_ticket, sh_base, sh_skip, sh_rewalk;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t bar = smem_u32(&bar_storage);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const uint32_t* d32 = reinterpret_cast<const uint32_t*>(data);
    uint32_t bulk_uses = 0;                                  // bulk copies this CTA has waited for: the mbarrier's phase

    for (;;) {
        lgw_cta_sync();                                       // every thread is done with the previous tile's shared memory
        // ---- 1. ticket; stage: ONE bulk copy for a tile that lies wholly inside the buffer
        if (tid == 0) {
            const uint32_t t = atomicAdd(&counters[2], 1u);
            sh_ticket = t;
            if (t < nwork) {
                const LgWork w0 = work[t];
                const FrameDev& F0 = frames[w0.frame];
                const unsigned long long off0 = (unsigned long long)w0.tile * LGW_TILE;
                if (off0 + (unsigned long long)LGW_DATA <= F0.len) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)LGW_DATA) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 ::"r"(smem_u32(data)), "l"(F0.src + off0), "r"((uint32_t)LGW_DATA), "r"(bar) : "memory");
                }
            }
        }
        lgw_cta_sync();
        const uint32_t ticket = sh_ticket;
        if (ticket >= nwork) break;
        const LgWork wk = work[ticket];
        const FrameDev& F = frames[wk.frame];
        const unsigned long long len = F.len;
        const uint32_t ntile = (uint32_t)max((len + LGW_TILE - 1) / LGW_TILE, 1ull);
        const uint32_t tile = wk.tile;
        const unsigned long long tile_off = (unsigned long long)tile * LGW_TILE;
        const uint32_t tile_rel = (uint32_t)min(len > tile_off ? len - tile_off : 0ull, (unsigned long long)(1u << 30));
        const bool last_tile = tile + 1 == ntile;
        const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;                            // pairs per row (RawData_Legacy.cpp:34-36)
        const uint32_t need_pairs = ppr * (uint32_t)F.height;                            // < 2^26: width * height <= 2^30 (prepare())
        const unsigned long long need = 2ull * need_pairs;                               // blocks of the image (:478-482)
        const bool fits = F.dst_cap >= (unsigned long long)F.width * (unsigned long long)F.height;
        if (tile_off + (unsigned long long)LGW_DATA <= len) {
            for (;;) {
                uint32_t ok;
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(bar), "r"(bulk_uses & 1u) : "memory");
                if (ok) break;
                __nanosleep(32);                              // the issue slots belong to the warps that have their data
            }
            bulk_uses++;
        } else {                                              // the tail of the buffer: 16-byte granules with zero fill
            lg_stage_tail(data, F.src, len, tile_off, LGW_DATA, (int)tid);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");           // a later bulk copy overwrites these generic stores
            lgw_cta_sync();
        }

        // ---- 2. chain C0 (all threads), then the transfer map (warp 0)
        uint32_t total0;
        const uint32_t ex0 = lgw_chain(data, bitmap, sh_exit, warp_sums, 0u, tile_rel, tid, total0);
        if (warp == 0) {
            const uint32_t exit0 = (ex0 == LGW_NONE || last_tile) ? LG_DEAD : ex0 >> 1;
            if (lane < LG_STATES) {
                uint32_t q = 2u * lane, pre = 0, m = LG_NO_MERGE, ex = exit0, count;
                bool d2 = false;
                if (lane == 0) { m = 0; count = total0; }
                else {
                    for (;;) {
                        if (q >= (uint32_t)LGW_TILE) { ex = (q - LGW_TILE) >> 1; break; }               // never met C0 in this tile
                        if (lgw_marked(bitmap, q)) { m = q >> 1; break; }
                        const uint32_t nq = q + leg_step(data[q]);
                        if (nq >= tile_rel) { d2 = true; break; }
                        q = nq;
                        pre++;
                    }
                    if (m != LG_NO_MERGE) {
                        // blocks of C0 from the merge point on = total0 - (marks before it)
                        count = pre + total0 - lgw_marks_before(bitmap, sh_exit, 2u * m);
                    } else {
                        count = pre;
                        if (d2 || last_tile) ex = LG_DEAD;
                    }
                }
                const uint32_t mapv = ex | (count << 5);
                sh_map[lane] = mapv;
                sh_merge[lane] = m;
                F.lg_tilemap[(size_t)tile * LG_STATES + lane] = mapv;
            }
            __syncwarp();
            // ---- 3. publish (release: the map entries the other lanes wrote are ordered before it by the warp barrier), look back
            if (lane == 0) lgw_store_release(F.lg_status + tile, lgw_pack(0u, epoch, LGW_ST_LOCAL, 0u));
            uint32_t entry = 0, base = 0, errbit = 0;
            if (tile > 0) {
                const uint32_t jhi = tile - 1;
                uint32_t spins = 0;
                for (;;) {
                    const int j = (int)jhi - (int)lane;
                    unsigned long long sw = 0;
                    if (lane < (uint32_t)LGW_LB && j >= 0) sw = lgw_load_acquire(F.lg_status + j);
                    const uint32_t lo32 = (uint32_t)sw;
                    const uint32_t st = ((lo32 >> 8) & 0xFFFFFFu) == (epoch & 0xFFFFFFu) ? (lo32 >> 6) & 3u : 0u;
                    const unsigned incl = __ballot_sync(0xFFFFFFFFu, st == LGW_ST_INCL);
                    const unsigned any = __ballot_sync(0xFFFFFFFFu, st != 0u);
                    if (incl) {
                        const int d = __ffs(incl) - 1;            // nearest predecessor with a known inclusive state: tile jhi - d
                        const unsigned between = (1u << d) - 1u;  // tiles jhi - d + 1 .. jhi must have published their maps
                        if ((any & between) == between) {
                            uint32_t state = __shfl_sync(0xFFFFFFFFu, lo32 & 31u, d);
                            uint32_t count = __shfl_sync(0xFFFFFFFFu, (uint32_t)(sw >> 32), d);
                            errbit = __shfl_sync(0xFFFFFFFFu, lo32 & LGW_ERR_BIT, d);
                            const uint32_t first = jhi - (uint32_t)d + 1u;                          // maps of tiles first .. jhi
                            __syncwarp();                        // every lane's acquire load before any lane's map loads
                            for (uint32_t idx = lane; idx < (uint32_t)d * LG_STATES; idx += 32)
                                lbmaps[idx] = __ldcg(F.lg_tilemap + (size_t)first * LG_STATES + idx);
                            __syncwarp();
                            if (lane == 0) {
                                for (int m = 0; m < d; m++) {
                                    if (state == LG_DEAD) break;
                                    const uint32_t v = lbmaps[m * LG_STATES + state];
                                    count += v >> 5;
                                    state = v & 31u;
                                }
                            }
                            entry = __shfl_sync(0xFFFFFFFFu, state, 0);
                            base = __shfl_sync(0xFFFFFFFFu, count, 0);
                            break;
                        }
                    }
                    if (++spins > LGW_SPIN_LIMIT) { entry = LG_DEAD; errbit = LGW_ERR_BIT; break; }   // never expected
                    __nanosleep(spins < 8 ? 40 : 200);
                }
            }
            uint32_t exitv = LG_DEAD, total = base;
            if (entry != LG_DEAD) {
                const uint32_t v = sh_map[entry];
                exitv = v & 31u;
                total = base + (v >> 5);
            }
            const bool skip = entry == LG_DEAD || !fits || (unsigned long long)base >= need;     // nothing of the image starts here
            if (lane == 0) {
                lgw_store_relaxed(F.lg_status + tile, lgw_pack(total, epoch, LGW_ST_INCL, exitv | errbit));   // self-contained word
                if (last_tile) {
                    unsigned status = 0;
                    if (!fits) status |= MCRAW_FRAME_GEOMETRY;
                    if ((unsigned long long)total < need) status |= MCRAW_FRAME_TRUNCATED;     // reference: stale samples (:387,398)
                    if (errbit) status |= MCRAW_FRAME_INTERNAL;
                    Result r;
                    r.written = status ? 0ull : (unsigned long long)F.width * (unsigned long long)F.height;   // :494
                    r.status = status;
                    r.pad = 0;
                    results[wk.frame] = r;
                }
                sh_base = base;
                sh_skip = skip ? 1u : 0u;
                sh_rewalk = 0;
            }
            // ---- 4. the bitmap for the true entry
            if (!skip && entry != 0) {
                const uint32_t m = sh_merge[entry];
                __syncwarp();
                if (m == LG_NO_MERGE) {
                    if (lane == 0) sh_rewalk = 2u * entry;                         // blocks of one constant width: see below
                } else {
                    const uint32_t mseg = (2u * m) / (uint32_t)LGW_SEG, mpos = (2u * m - mseg * (uint32_t)LGW_SEG) >> 1;
                    for (uint32_t w = lane; w < 2u * mseg; w += 32) bitmap[w] = 0;   // C0's marks before the merge point go
                    __syncwarp();
                    if (lane == 0) {
                        if (mpos >= 32u) { bitmap[2u * mseg] = 0; bitmap[2u * mseg + 1] &= ~((1u << (mpos - 32u)) - 1u); }
                        else bitmap[2u * mseg] &= ~((1u << mpos) - 1u);
                        uint32_t p = 2u * entry;
                        while (p < 2u * m) {
                            const uint32_t sg = p / (uint32_t)LGW_SEG, ps = (p - sg * (uint32_t)LGW_SEG) >> 1;
                            bitmap[2u * sg + (ps >> 5)] |= 1u << (ps & 31u);
                            p += leg_step(data[p]);
                        }
                    }
                }
            }
        }
        lgw_cta_sync();
        if (sh_skip) continue;
        const uint32_t base = sh_base;
        if (sh_rewalk) {                                      // the entry's chain never meets C0 in this tile: walk it from its entry
            uint32_t t2;
            lgw_chain(data, bitmap, sh_exit, warp_sums, sh_rewalk, tile_rel, tid, t2);
        }

        // ---- 5. pair list of the tile: every block with an even ordinal leads a pair (even-column block, then odd-column
        //      block, RawData_Legacy.cpp:480-481); plist[q] = (tile-relative offset of the leader) / 2 for pair ordinal
        //      p_first + q.  Ordinals come from prefix popcounts of the marks.  Consecutive lanes then decode consecutive
        //      pairs: their reads of the staged tile are a few words apart (different banks), their stores adjacent.
        const uint32_t wv0 = bitmap[2 * tid], wv1 = bitmap[2 * tid + 1];
        const uint32_t c = __popc(wv0) + __popc(wv1);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        uint32_t before = 0, total = incl;
        if (LGW_WARPS > 1) {
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            total = 0;
#pragma unroll
            for (int w = 0; w < LGW_WARPS; w++) {
                const uint32_t v = warp_sums[w];
                if ((uint32_t)w < warp) before += v;
                total += v;
            }
        } else {
            total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        const uint32_t p_first = (base + 1u) >> 1;
        uint32_t npairs = ((base + total + 1u) >> 1) - p_first;
        npairs = min(npairs, need_pairs > p_first ? need_pairs - p_first : 0u);
        const uint32_t ord0 = base + before + incl - c;          // ordinal of the first block start in this thread's segment
        const int width = F.width;
        uint16_t* __restrict__ dst = F.dst;
        const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
        const unsigned epi = EPI ? F.epi_mode : 0u;             // EPI = false: the epilogue code is not even in the kernel
        for (uint32_t c0 = 0; c0 < npairs; c0 += LGW_PAIR_CHUNK) {
            const uint32_t cn = min((uint32_t)LGW_PAIR_CHUNK, npairs - c0);
            {
                uint32_t ord = ord0;
                unsigned long long marks = (unsigned long long)wv0 | ((unsigned long long)wv1 << 32);
                while (marks) {
                    const uint32_t b = (uint32_t)__ffsll((long long)marks) - 1u;
                    marks &= marks - 1;
                    const uint32_t q = (ord >> 1) - p_first - c0;       // wraps to a huge value for earlier passes' pairs
                    if (!(ord & 1u) && q < cn) plist[q] = (uint16_t)(((uint32_t)LGW_SEG / 2u) * tid + b);
                    ord++;
                }
            }
            lgw_cta_sync();
            uint32_t P = p_first + c0 + tid;
            uint32_t y = P / ppr, xq = P - y * ppr;
            for (uint32_t q = tid; q < cn; q += LGW_THREADS) {
                const uint32_t oE = 2u * (uint32_t)plist[q];
                const uint32_t hE = leg_header(data, oE), bitsE = leg_hdr_bits(hE);
                const uint32_t oO = oE + 2u + leg_len(bitsE);
                const uint32_t hO = leg_header(data, oO), bitsO = leg_hdr_bits(hO);
                uint32_t px[16];
#pragma unroll
                for (int i = 0; i < 16; i++) px[i] = 0;
                lgw_block<0>(d32, oE, bitsE, px);
                lgw_block<16>(d32, oO, bitsO, px);
                const uint32_t refs = leg_hdr_ref(hE) | (leg_hdr_ref(hO) << 16);
#pragma unroll
                for (int i = 0; i < 16; i++) px[i] = __vadd2(px[i], refs);                    // :483-486, + reference mod 2^16
                if (epi) {                                                                    // optional black / white level epilogue
#pragma unroll
                    for (int i = 0; i < 16; i++) px[i] = epilogue_word(px[i], epi, F, (int)(y & 1u));
                }
                const int x = (int)(32u * xq);
                uint16_t* orow = dst + (size_t)y * (size_t)width + x;
                if (vec && x + 32 <= width) {
                    uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
                    for (int i = 0; i < 4; i++) o4[i] = make_uint4(px[4 * i], px[4 * i + 1], px[4 * i + 2], px[4 * i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {                                            // crop at width (:490)
                        if (x + 2 * i < width) orow[2 * i] = (uint16_t)px[i];
                        if (x + 2 * i + 1 < width) orow[2 * i + 1] = (uint16_t)(px[i] >> 16);
                    }
                }
                xq += LGW_THREADS;                                                            // the pair LGW_THREADS further on
                while (xq >= ppr) { xq -= ppr; y++; }
            }
            lgw_cta_sync();
        }
    }
    // the last CTA to leave resets the ticket counters for the next launch
    if (tid == 0 && atomicAdd(&counters[3], 1u) == gridDim.x - 1u) { counters[2] = 0; counters[3] = 0; }
}

}  // namespace mcraw
