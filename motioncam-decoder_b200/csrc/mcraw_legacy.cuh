// mcraw_legacy.cuh -- sm_100a kernel for the legacy frame format (compressionType 6).
//
// Reference: /root/reference/lib/RawData_Legacy.cpp:445-495 (raw::DecodeLegacy) and :372-442 (DecodeHeader /
// DecodeBlock).  A frame is one chain of 16-sample blocks, each with an inline 2-byte header
// (bits nibble, 12-bit reference) followed by 2*bits payload bytes (32 for nibbles 11..15): block k+1 starts
// where block k ends, so the reference finds the blocks with 750 000 dependent steps per 4000x3000 frame.
//
// Here the chain is resolved in parallel, EXACTLY and in ONE pass over the stream.  What carries it: all block lengths
// are even and <= 34 bytes, so a chain enters any fixed byte range at one of 17 even offsets -- the effect of a byte
// range on the chain is a map  entry (17 values) -> exit (17 values), maps compose, and the map of a short range is
// cheap to compute for ALL entries at once by going through its candidate block starts backwards.
//
// k_legacy_warp: persistent CTAs of LGW_THREADS threads take (frame, tile) tickets in a host-built order (tile index
// major, frame minor: neighbouring tickets belong to different frames).  A tile is one segment of LGW_SEG bytes per
// thread.  Per tile:
//   1. stage the tile (+ overrun) with ONE bulk copy (cp.async.bulk, TMA 1-D, mbarrier transaction count) -- the only
//      time the stream is read; the tile of the ticket one grid further on is prefetched into L2;
//   2. every thread: exit table of its segment.  For candidate block start c (even offsets, last to first)
//      exit[c] = exit[c + length of the block whose header sits at c]; positions past the segment are their own exit.
//      62 uniform steps, no guessing, no divergence; the first 17 entries are the segment's map;
//   3. 17 lanes of every warp chase the 17 possible entries through the warp's 32 segment maps and leave, per segment,
//      the entry each of them arrives with; composing the warps' maps gives the tile's map  entry -> exit;
//   4. exits across tiles: nearly always the 17 exits of a tile are one value (chains started anywhere fall in step
//      within a few hundred bytes), so the tile publishes "exit X whatever the entry" and its successor knows its entry
//      after ONE hop -- no dependency chain across tiles.  Otherwise (runs of one constant block length) the tile
//      publishes its map and, once it knows its own entry, its actual exit; successors compose the maps back to the
//      nearest known exit;
//   5. with its exact entry every thread walks its segment once: block starts (a bitmap in two registers) and their count;
//      the tile's count goes through a plain DECOUPLED LOOK-BACK prefix sum over the tiles of the frame (aggregate /
//      inclusive words): block ordinals give pair parity and the output position (:478-482);
//   6. pair list from the bitmaps (even-ordinal block + the block behind it, :480-481), then consecutive threads decode
//      consecutive pairs straight from shared memory: MSB-first bit extraction with rotates (:38-370), + reference
//      mod 2^16, column interleave (:483-486) in registers, 16-byte stores, crop at width (:490).
// Waits are only ever for the status words of EARLIER tickets, whose owners are resident or done and never wait for a
// later ticket -- so they end; they are bounded all the same (MCRAW_FRAME_INTERNAL instead of a hang).  Status words
// carry the launch epoch of the slot, so nothing has to be zeroed between launches.
#pragma once
#include "mcraw_kernels.cuh"

namespace mcraw {

#ifndef MCRAW_LGW_WARPS
#define MCRAW_LGW_WARPS 4
#endif
constexpr int LGW_WARPS = MCRAW_LGW_WARPS;     // warps per CTA
constexpr int LGW_THREADS = 32 * LGW_WARPS;
constexpr int LGW_SEG = 124;                   // bytes of the tile a thread owns: 62 candidate block starts, two words of marks.
                                               // 31 words: threads at the same offset of their segments hit 32 different banks
constexpr int LGW_NC = LGW_SEG / 2;            // candidate block starts per segment
constexpr int LGW_TILE = LGW_THREADS * LGW_SEG;   // bytes a CTA stages and indexes at a time (the WINDOW): 15 872 for four warps
#ifndef MCRAW_LGW_RUN
#define MCRAW_LGW_RUN 8
#endif
constexpr int LGW_RUN = MCRAW_LGW_RUN;                     // the first LGW_RUN segments of a window are the RUN-UP: they belong to the tile before
                                               // (for tile 0 they are its own); chains started at all 17 offsets of the window start
                                               // have nearly always become one by the end of the run-up, and then the tile knows its
                                               // entry without waiting for anybody
constexpr int LGW_RUNB = LGW_RUN * LGW_SEG;
constexpr int LGW_STRIDE = LGW_TILE - LGW_RUNB;   // window t starts at byte t * LGW_STRIDE; tile t owns window bytes [LGW_RUNB (0 for t = 0), LGW_TILE)
constexpr int LGW_TAB = 84;                    // bytes of exit table per thread: LGW_NC candidates + 17 positions behind the segment
                                               // (+ pad); 21 words, odd, for the same reason
constexpr int LGW_ENT = 32;                    // table offset of the 17 entries a segment is reached with (one per warp entry):
                                               // candidates 17.. of the table are dead once the segment's map is complete
constexpr int LGW_PAIR_CHUNK = LGW_WARPS >= 4 ? 1024 : 512;   // pairs listed and decoded per pass (a tile of four warps holds ~700 for
                                               // typical images, up to LGW_TILE / 4 when every block is 2 bytes: then several passes)
constexpr int LGW_OUT = 2048;                  // bytes of decoded pixels a warp hands to one bulk store: 32 pairs x 64 bytes
constexpr int LG_STATES = 17;                  // entry offsets 0, 2, ..., 32
constexpr uint32_t LG_DEAD = 31;               // exit code of a chain that ran into the end of the buffer
constexpr int LG_OVERRUN = 128;                // a pair led inside the tile ends at most 2 + 34 + 34 bytes past it (+ word reads); keeps what follows 64-byte aligned
constexpr int LGW_DATA = LGW_TILE + LG_OVERRUN;
#ifndef MCRAW_LGW_PF_DIV
#define MCRAW_LGW_PF_DIV 1
#endif
constexpr int LGW_PF_DIV = MCRAW_LGW_PF_DIV;   // L2 prefetch distance: (resident CTAs) / LGW_PF_DIV tickets ahead
constexpr int LGW_LB = 32;                     // look-back window: status words read per poll
constexpr int LGW_SMEM = LGW_DATA + LGW_THREADS * LGW_TAB;      // the pair list and the output staging reuse the tables
constexpr uint32_t LGW_ST_AGG = 1u, LGW_ST_INCL = 2u;          // count words: this tile's blocks / all blocks up to its end
constexpr uint32_t LGW_EX_CONV = 1u, LGW_EX_MAP = 2u, LGW_EX_FINAL = 3u;   // exit words, see 4. above
constexpr uint32_t LGW_ERR_BIT = 1u << 5;      // sticky: a wait gave up somewhere up the chain
constexpr uint32_t LGW_SPIN_LIMIT = 1u << 18;  // polls of ~0.2 us: a wait that long means something is broken, not slow
static_assert(LGW_DATA % 64 == 0 && LGW_TILE % 16 == 0 && LGW_STRIDE % 16 == 0 && (2 * LGW_PAIR_CHUNK) % 64 == 0 && LGW_OUT % 64 == 0,
              "bulk copies work in 16-byte granules; the output staging slots are 64-byte aligned");
static_assert(LGW_RUN >= 1 && LGW_RUN < 32 && LGW_RUNB >= 34, "the run-up lies inside warp 0 and is at least one block long");
static_assert(LGW_SEG % 4 == 0 && LGW_NC <= 64 && LGW_SEG >= 34, "two words of marks per segment; a block never skips a segment");
static_assert(LGW_NC + LG_STATES <= LGW_TAB && LGW_ENT >= LG_STATES && LGW_ENT + LG_STATES <= LGW_NC && LGW_TAB % 4 == 0, "exit table layout");
static_assert(LGW_PAIR_CHUNK * 2 + LGW_WARPS * LGW_OUT <= LGW_THREADS * LGW_TAB && (LGW_PAIR_CHUNK * 2) % 16 == 0,
              "the pair list and the warps' output staging reuse the tables");

struct LgWork { uint32_t frame, tile; };
// tiles of a frame buffer of len bytes
__host__ __device__ inline unsigned long long lgw_ntiles(const unsigned long long len) {
    return len > (unsigned long long)LGW_RUNB ? (len - LGW_RUNB + LGW_STRIDE - 1) / LGW_STRIDE : 1ull;
}

// payload bytes of a 16-sample block for header nibble b (RawData_Legacy.cpp:13-32, min(16, bits) at :395)
__device__ __forceinline__ uint32_t leg_len(uint32_t b) { return b <= 10u ? 2u * b : 32u; }
// total length of the block whose header byte is hb
__device__ __forceinline__ uint32_t leg_step(uint32_t hb) {
    const uint32_t b = hb >> 4;
    return 2u + (b > 10u ? 32u : 2u * b);
}
// The 2-byte header at byte offset o (even) of the staged tile, as the low 16 bits of the result (byte o first).
__device__ __forceinline__ uint32_t leg_header(const uint8_t* data, uint32_t o) {
    return *reinterpret_cast<const uint16_t*>(data + o);
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

// status words of a tile (two per tile: [2 * tile] count word, [2 * tile + 1] exit word):
//   value << 32 | epoch (24 bits) << 8 | state << 6 | error << 5 | exit offset / 2
__device__ __forceinline__ unsigned long long lgw_pack(uint32_t count, uint32_t epoch, uint32_t st, uint32_t low6) {
    return ((unsigned long long)count << 32) | ((unsigned long long)(epoch & 0xFFFFFFu) << 8) | (st << 6) | low6;
}
__device__ __forceinline__ uint32_t lgw_state(const unsigned long long w, const uint32_t epoch) {
    const uint32_t lo32 = (uint32_t)w;
    return ((lo32 >> 8) & 0xFFFFFFu) == (epoch & 0xFFFFFFu) ? (lo32 >> 6) & 3u : 0u;
}
__device__ __forceinline__ void lgw_store_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lgw_load_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One thread, its segment [seg0, seg0 + LGW_SEG) of the staged tile: walk from half offset `pos` of the segment (byte
// seg0 + 2 pos) to the end of the segment; block starts go into (m0, m1), one bit per even offset.  TAIL: the buffer may
// end inside or right behind the window; rel: bytes from the window start to the end of the buffer (a block is decoded
// only if it ends before the last byte, RawData_Legacy.cpp:387,398); the chain stops at the first block that is not.
template <bool TAIL>
__device__ __forceinline__ void lgw_walk_segment(const uint8_t* data, const uint32_t seg0, uint32_t pos, uint32_t& m0, uint32_t& m1, const uint32_t rel) {
    unsigned long long m = 0;
    const uint8_t* sp = data + seg0;
    while (pos < (uint32_t)LGW_NC) {
        const uint32_t b = (uint32_t)sp[2u * pos] >> 4;
        const uint32_t h = b > 10u ? 16u : b;                         // payload / 2
        if (TAIL) { if (seg0 + 2u * pos + 2u + 2u * h >= rel) break; }
        m |= 1ull << pos;
        pos += 1u + h;
    }
    m0 = (uint32_t)m; m1 = (uint32_t)(m >> 32);
}

// One lane, one of the 17 states a warp's first segment can be entered with: follow it through the warp's 32 segment maps
// (Tw: the warp's tables) and leave, in every segment's table, the state that segment is entered with.  Returns the state
// the warp is left with; eR: the state the first segment behind the run-up is entered with (only warp 0's is used).
// All offsets are immediates: a store, an add and a load per segment.  TAIL: chains can end (LG_DEAD) inside the window.
template <bool TAIL>
__device__ __forceinline__ uint32_t lgw_chase(uint8_t* const Tw, const uint32_t lane, uint32_t& eR) {
    uint32_t e = lane;
    uint8_t* const Te = Tw + LGW_ENT + lane;
#pragma unroll
    for (int s = 0; s < 32; s++) {
        if (s == LGW_RUN) eR = e;
        Te[s * LGW_TAB] = (uint8_t)e;
        if (!TAIL || e != LG_DEAD) e = Tw[s * LGW_TAB + e];
    }
    return e;
}

// One thread: the exit table of its segment.  T[c], c = 0 .. LGW_NC - 1: where the chain that has a block start at byte
// 2c of the segment enters the next segment (half offset 0 .. 16), or LG_DEAD if it ends at a block that cannot be decoded.
// dw: the segment as words.  TAIL: the buffer may end inside or right behind the tile (rel as above, seg0 = the segment's
// tile-relative offset); interior tiles skip that test.  The two candidates of a word go together: both look-ups are
// issued before either result is stored (the second one can only need the first one's result when its block is 2 bytes
// long -- then it is taken from the register), which halves the chain of dependent shared-memory round trips.
template <bool TAIL>
__device__ __forceinline__ void lgw_exit_table(const uint32_t* __restrict__ dw, uint8_t* T, const uint32_t seg0, const uint32_t rel) {
    // positions behind the segment are their own exit: T[LGW_NC + k] = k.  Written as words from T[LGW_NC - 2] on (the table
    // is word-aligned and LGW_NC is even); the two candidates caught by the first word are written before they are read.
    uint32_t* tw = reinterpret_cast<uint32_t*>(T + LGW_NC - 2);
    tw[0] = 0x01000000u;
#pragma unroll
    for (int k = 1; k < 5; k++) tw[k] = 0x03020100u + 0x04040404u * (uint32_t)k - 0x02020202u;
#pragma unroll
    for (int w = LGW_SEG / 4 - 1; w >= 0; --w) {
        const uint32_t v = dw[w];
        const int c0 = 2 * w, c1 = 2 * w + 1;
        const uint32_t b1 = (v >> 20) & 15u, b0 = (v >> 4) & 15u;
        const uint32_t h1 = b1 > 10u ? 16u : b1, h0 = b0 > 10u ? 16u : b0;                  // payload / 2
        uint32_t x1 = T[c1 + 1 + h1];
        uint32_t x0 = T[c0 + 1 + h0];                                 // = T[c1] (not yet written) when h0 == 0
        if (TAIL) { if (seg0 + 2u * (uint32_t)c1 + 2u + 2u * h1 >= rel) x1 = LG_DEAD; }
        if (h0 == 0u) x0 = x1;
        if (TAIL) { if (seg0 + 2u * (uint32_t)c0 + 2u + 2u * h0 >= rel) x0 = LG_DEAD; }
        *reinterpret_cast<uint16_t*>(T + c0) = (uint16_t)(x0 | (x1 << 8));
    }
}

__device__ __forceinline__ uint32_t lg_prmt(const uint32_t a, const uint32_t b, const uint32_t sel) {   // selector nibbles must be 0 .. 7
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// The 16 samples of the block at byte offset o (header nibble `bits`) go into the low (SECOND = false: even-column block)
// or high (SECOND = true: odd-column block) halves of px[] (RawData_Legacy.cpp:483-486); returns the mask of a sample's
// valid bits -- what lands above a sample inside its half is left for the caller to clear (one AND per word for both
// blocks).  d32: the staged tile as words.  The payload is a contiguous MSB-first bit stream (:38-358), so 8 samples
// are exactly `bits` bytes: per group of 8 the bytes are fetched as big-endian words (PRMT with a runtime selector does
// alignment and byte order in one go) into windows with a few samples at the top, and a sample is the window shifted
// right.  The arithmetic pipe is this kernel's bottleneck, so three of four shifts are multiplications
// (mul.hi by a power of two, on the FMA pipe) and the halves are merged by PRMT.  Three lane-uniform formulations
// instead of one code path per width: widths 0..8 (two 4-sample windows per group), 9..10 (four 2-sample windows),
// and 16-bit big-endian samples (:360-370, nibbles 11..15, :395).
// PIECE ORDER: px[k] receives sample k ^ (4 * rot), rot = 0 .. 3 (rot1 = rot & 1, rot2 = rot & 2 as flags) -- the
// lane's four 16-byte output pieces, exchanged, which is what lets a warp write its 32 x 64 bytes into LINEAR shared memory
// without bank conflicts (see the bulk store in the kernel).  The exchange is applied to the windows, not to the samples.
template <bool SECOND>
__device__ __forceinline__ uint32_t lgw_block(const uint32_t* __restrict__ d32, const uint32_t o, const uint32_t bits, uint32_t (&px)[16],
                                              const bool rot1, const bool rot2, const uint32_t rot) {
    const uint32_t a = o + 2u;
#define LGW_PUT(k, v) px[k] = SECOND ? lg_prmt(px[k], (v), 0x5410u) : (v)
    if (bits <= 8u) {
        const uint32_t w = bits;
        uint32_t win[4];                                                               // samples 4t .. 4t+3 sit at the top of win[t]
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t ag = a + g * w;
            const uint32_t i = ag >> 2, sel = 0x0123u + 0x1111u * (ag & 3u);
            const uint32_t W0 = d32[i], W1 = d32[i + 1], W2 = d32[i + 2];
            const uint32_t G0 = lg_prmt(W0, W1, sel), G1 = lg_prmt(W1, W2, sel);
            win[2 * g] = G0;
            win[2 * g + 1] = __funnelshift_lc(G1, G0, 4u * w);                         // the window of samples 4..7
        }
        {
            const uint32_t t0 = rot1 ? win[1] : win[0], t1 = rot1 ? win[0] : win[1], t2 = rot1 ? win[3] : win[2], t3 = rot1 ? win[2] : win[3];
            win[0] = rot2 ? t2 : t0; win[1] = rot2 ? t3 : t1; win[2] = rot2 ? t0 : t2; win[3] = rot2 ? t1 : t3;
        }
        const uint32_t m1 = 1u << w, m2 = 1u << (2u * w), m3 = 1u << (3u * w);         // x >> (32 - k w) = mul.hi(x, 2^(k w)), k w <= 24
        const uint32_t s4 = (32u - 4u * w) & 31u;                                       // w = 8: the window itself (w = 0: anything, mask 0)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            LGW_PUT(4 * t, __umulhi(win[t], m1));
            LGW_PUT(4 * t + 1, __umulhi(win[t], m2));
            LGW_PUT(4 * t + 2, __umulhi(win[t], m3));
            LGW_PUT(4 * t + 3, win[t] >> s4);
        }
        return m1 - 1u;
    } else if (bits <= 10u) {
        const uint32_t w = bits;
        uint32_t win[8];                                                               // samples 2n, 2n+1 sit at the top of win[n]
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const uint32_t ag = a + g * w;
            const uint32_t i = ag >> 2, sel = 0x0123u + 0x1111u * (ag & 3u);
            const uint32_t W0 = d32[i], W1 = d32[i + 1], W2 = d32[i + 2], W3 = d32[i + 3];
            const uint32_t G0 = lg_prmt(W0, W1, sel), G1 = lg_prmt(W1, W2, sel), G2 = lg_prmt(W2, W3, sel);
            win[4 * g] = G0;                                                           // samples 2m, 2m+1 start at bit 2mw
            win[4 * g + 1] = __funnelshift_l(G1, G0, 2u * w);
            win[4 * g + 2] = __funnelshift_l(G2, G1, 4u * w - 32u);
            win[4 * g + 3] = __funnelshift_l(G2, G1, 6u * w - 32u);
        }
        {
            uint32_t t[8];
#pragma unroll
            for (int n = 0; n < 8; n++) t[n] = rot1 ? win[n ^ 2] : win[n];
#pragma unroll
            for (int n = 0; n < 8; n++) win[n] = rot2 ? t[n ^ 4] : t[n];
        }
        const uint32_t m1 = 1u << w, m2 = 1u << (2u * w);
#pragma unroll
        for (int n = 0; n < 8; n++) {
            LGW_PUT(2 * n, __umulhi(win[n], m1));
            LGW_PUT(2 * n + 1, __umulhi(win[n], m2));
        }
        return m1 - 1u;
    } else {
        const uint32_t i = a >> 2, k0 = a & 3u;                                        // the payload starts 2-byte aligned
        // sample k = bytes (2k, 2k+1), big-endian: byte index (k0 + 2 (k & 1)) of the word pair (k / 2, k / 2 + 1);
        // the selector puts (high byte, low byte) at result bytes (1, 0)
        const uint32_t selA = (k0 + 1u) | (k0 << 4), selB = (k0 + 3u) | ((k0 + 2u) << 4);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const uint32_t ms = (uint32_t)m ^ (2u * rot);
            const uint32_t lo = d32[i + ms], hi = d32[i + ms + 1u];
            LGW_PUT(2 * m, lg_prmt(lo, hi, selA));
            LGW_PUT(2 * m + 1, lg_prmt(lo, hi, selB));
        }
        return 0xFFFFu;
    }
#undef LGW_PUT
}

template <bool EPI>
__global__ void __launch_bounds__(LGW_THREADS, EPI ? 5 : 8) k_legacy_warp(const FrameDev* __restrict__ frames, Result* __restrict__ results,
                                                             const LgWork* __restrict__ work, const uint32_t nwork,
                                                             uint32_t* __restrict__ counters, const uint32_t epoch) {
    extern __shared__ __align__(128) uint8_t lg_smem[];
    // The NEXT batch's kernel may be launched as a programmatic dependent of this one (mcraw_capi.cu, "chain": its own slot's
    // tickets, status words and counters; outputs that are the same or disjoint): its CTAs move in as this grid's CTAs run out
    // of tickets, instead of waiting for the last of them.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    uint8_t* data = lg_smem;
    uint8_t* tables = lg_smem + LGW_DATA;                                                   // [LGW_THREADS][LGW_TAB]
    uint16_t* plist = reinterpret_cast<uint16_t*>(tables);                                  // pair list: [LGW_PAIR_CHUNK], after the tables
    __shared__ __align__(8) unsigned long long bar_storage;
    __shared__ uint8_t sh_wmap[LGW_WARPS][LG_STATES + 3];                                   // map of each warp's 32 segments
    __shared__ uint32_t warp_sums[LGW_WARPS];
    __shared__ uint32_t sh_ticket, sh_base, sh_skip, sh_err, sh_runconv, sh_cw0;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t bar = smem_u32(&bar_storage);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const uint32_t* d32 = reinterpret_cast<const uint32_t*>(data);
    const uint32_t seg0 = tid * (uint32_t)LGW_SEG;
    uint8_t* T = tables + tid * (uint32_t)LGW_TAB;
    uint32_t bulk_uses = 0;                                  // bulk copies this CTA has waited for: the mbarrier's phase
    const uint32_t rot = (lane >> 1) & 3u;                   // exchange of this lane's output pieces (lgw_block)
    const bool rot1 = (rot & 1u) != 0, rot2 = (rot & 2u) != 0;
    const uint32_t out_s = smem_u32(tables) + 2u * (uint32_t)LGW_PAIR_CHUNK + (uint32_t)LGW_OUT * warp;   // this warp's output staging

    uint32_t next_ticket = 0;                                // thread 0: the ticket drawn during the last decode pass of the previous tile
    bool have_next = false;                                  // (one pass early hides the atomic's round trip; earlier than that, tiles
                                                             // that wait for this one's block count would wait longer)
    for (;;) {
        __syncthreads();                                      // every thread is done with the previous tile's shared memory
        // ---- 1. ticket; stage: ONE bulk copy for a tile that lies wholly inside the buffer
        if (tid == 0) {
            const uint32_t t = have_next ? next_ticket : atomicAdd(&counters[2], 1u);
            have_next = false;
            sh_ticket = t;
            if (t < nwork) {
                const LgWork w0 = work[t];
                const FrameDev& F0 = frames[w0.frame];
                const unsigned long long off0 = (unsigned long long)w0.tile * LGW_STRIDE;
                if (off0 + (unsigned long long)LGW_DATA <= F0.len) {
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic accesses of the last tile before the async write
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)LGW_DATA) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 ::"r"(smem_u32(data)), "l"(F0.src + off0), "r"((uint32_t)LGW_DATA), "r"(bar) : "memory");
                }
            }
        }
        __syncthreads();
        const uint32_t ticket = sh_ticket;
        if (ticket >= nwork) break;
        if (tid == 0) {
            // the tile some CTA will take about one round of the grid from now: have it in L2 by then
            const uint32_t tn = ticket + gridDim.x / (uint32_t)LGW_PF_DIV;
            if (tn < nwork) {
                const LgWork w1 = work[tn];
                const FrameDev& F1 = frames[w1.frame];
                const unsigned long long off1 = (unsigned long long)w1.tile * LGW_STRIDE;
                if (off1 + (unsigned long long)LGW_DATA <= F1.len)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(F1.src + off1), "r"((uint32_t)LGW_DATA) : "memory");
            }
        }
        const LgWork wk = work[ticket];
        const FrameDev& F = frames[wk.frame];
        const unsigned long long len = F.len;
        const uint32_t ntile = (uint32_t)lgw_ntiles(len);
        const uint32_t tile = wk.tile;
        const unsigned long long tile_off = (unsigned long long)tile * LGW_STRIDE;   // where the window starts
        const uint32_t tile_rel = (uint32_t)min(len > tile_off ? len - tile_off : 0ull, (unsigned long long)(1u << 30));
        const bool last_tile = tile + 1 == ntile;
        const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;                            // pairs per row (RawData_Legacy.cpp:34-36)
        const uint32_t need_pairs = ppr * (uint32_t)F.height;                            // < 2^26: width * height <= 2^30 (prepare())
        const unsigned long long need = 2ull * need_pairs;                               // blocks of the image (:478-482)
        const bool fits = F.dst_cap >= (unsigned long long)F.width * (unsigned long long)F.height;
        unsigned long long* const cntw = F.lg_status;                                    // [2 * tile]: count word, [2 * tile + 1]: exit word
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
            __syncthreads();
        }

        // ---- 2. exit table of this thread's segment
        const bool interior = tile_rel > (uint32_t)LGW_TILE + 34u;     // no block that starts in this window reaches the end of the buffer
        if (interior) lgw_exit_table<false>(d32 + tid * (uint32_t)(LGW_SEG / 4), T, seg0, tile_rel);
        else lgw_exit_table<true>(d32 + tid * (uint32_t)(LGW_SEG / 4), T, seg0, tile_rel);
        __syncwarp();                                         // the chase stays inside the warp's own 32 tables
        // ---- 3. the 17 entries of this warp's first segment, chased through its 32 segments
        uint32_t eR = LG_DEAD;                                // warp 0: what the first segment behind the run-up is entered with
        if (lane < (uint32_t)LG_STATES) {
            uint8_t* const Tw = tables + (32u * warp) * (uint32_t)LGW_TAB;
            const uint32_t e = interior ? lgw_chase<false>(Tw, lane, eR) : lgw_chase<true>(Tw, lane, eR);
            sh_wmap[warp][lane] = (uint8_t)e;
        }
        if (warp == 0) {                                      // have all 17 chains become one inside the run-up?
            const uint32_t e0 = __shfl_sync(0xFFFFFFFFu, eR, 0);
            const bool one = __all_sync(0xFFFFFFFFu, lane >= (uint32_t)LG_STATES || eR == e0);
            if (lane == 0) { sh_runconv = (tile == 0u || one) ? 1u : 0u; sh_err = 0u; }
        }
        __syncthreads();
        // ---- 4. warp 0: the tile's exit word.  Tile 0 starts with a block; every other tile whose run-up merged all chains
        //      knows its entry already (any of the 17 chains is the true one from there on).  Only a tile whose run-up did not
        //      (runs of one constant block length) has to look back for the exit of the tile before it.
        const bool runconv = sh_runconv != 0u;
        if (warp == 0) {
            uint32_t x = lane < (uint32_t)LG_STATES ? lane : 0u;
#pragma unroll
            for (int w = 0; w < LGW_WARPS; w++)
                if (x != LG_DEAD) x = sh_wmap[w][x];          // where the chain that enters the window with `lane` leaves it
            if (last_tile) x = LG_DEAD;                       // nothing follows the last tile
            const uint32_t x0 = __shfl_sync(0xFFFFFFFFu, x, 0);
            const bool conv = tile == 0u || __all_sync(0xFFFFFFFFu, x == x0);   // tile 0: chain 0 is the chain
            // Status words are self-contained (value, epoch and state in one 64-bit store), so they travel as relaxed
            // gpu-scope accesses; only a published MAP refers to other memory (the 17 map entries): fence, then the word
            // on this side, the word, then a fence before the entries are read on the other.
            if (!conv) {                                      // map: entry of the TILE (behind the run-up) -> exit
                uint32_t* const tm = F.lg_tilemap + (size_t)tile * LG_STATES;
                if (lane < (uint32_t)LG_STATES) tm[lane] = LG_DEAD;           // entries no chain arrives with are never asked for
                __syncwarp();
                if (lane < (uint32_t)LG_STATES && eR != LG_DEAD) tm[eR] = x;  // chains that share eR share x
                __threadfence();
                __syncwarp();
            }
            if (lane == 0) lgw_store_relaxed(cntw + 2 * (size_t)tile + 1, lgw_pack(0u, epoch, conv ? LGW_EX_CONV : LGW_EX_MAP, x0));
            if (!runconv) {
                uint32_t entry = LG_DEAD, errbit = 0;
                const uint32_t jhi = tile - 1;                // (tile 0 never gets here)
                uint32_t spins = 0;
                for (;;) {
                    const int j = (int)jhi - (int)lane;
                    unsigned long long sw = 0;
                    if (j >= 0) sw = lgw_load_relaxed(cntw + 2 * (size_t)j + 1);
                    const uint32_t st = lgw_state(sw, epoch);
                    const unsigned known = __ballot_sync(0xFFFFFFFFu, st == LGW_EX_CONV || st == LGW_EX_FINAL);
                    const unsigned any = __ballot_sync(0xFFFFFFFFu, st != 0u);
                    if (known) {
                        const int d = __ffs(known) - 1;           // nearest predecessor whose exit is known: tile jhi - d
                        const unsigned between = (1u << d) - 1u;  // tiles jhi - d + 1 .. jhi must have published their maps
                        if ((any & between) == between) {
                            uint32_t state = __shfl_sync(0xFFFFFFFFu, (uint32_t)sw & 31u, d);
                            errbit = __shfl_sync(0xFFFFFFFFu, (uint32_t)sw & LGW_ERR_BIT, d);
                            if (d > 0) __threadfence();          // the status loads before the map loads
                            for (int m = d - 1; m >= 0 && state != LG_DEAD; m--)          // (rare: runs of one constant block length)
                                state = __ldcg(F.lg_tilemap + (size_t)(jhi - (uint32_t)m) * LG_STATES + state);
                            entry = state;
                            break;
                        }
                    }
                    if (++spins > LGW_SPIN_LIMIT) { entry = LG_DEAD; errbit = LGW_ERR_BIT; break; }   // never expected
                    __nanosleep(spins < 8 ? 40 : 200);
                }
                // the chain of this window that arrives behind the run-up with `entry` (the true chain is one of the 17)
                const unsigned match = __ballot_sync(0xFFFFFFFFu, lane < (uint32_t)LG_STATES && eR == entry && entry != LG_DEAD);
                const uint32_t cw0 = match ? (uint32_t)__ffs((int)match) - 1u : LG_DEAD;
                // the actual exit of a tile that published a map: successors stop composing here
                const uint32_t x_act = cw0 == LG_DEAD ? LG_DEAD : __shfl_sync(0xFFFFFFFFu, x, cw0);
                if (!conv && lane == 0) lgw_store_relaxed(cntw + 2 * (size_t)tile + 1, lgw_pack(0u, epoch, LGW_EX_FINAL, x_act | errbit));
                if (lane == 0) { sh_cw0 = cw0; sh_err = errbit; }
            }
        }
        if (!runconv) __syncthreads();                        // (uniform)
        // ---- 5. the exact chain: every thread walks its segment from the entry it is reached with
        uint32_t wv0 = 0, wv1 = 0;
        {
            uint32_t cw = runconv ? 0u : sh_cw0;             // what the window is entered with (any chain, if they merged in the run-up)
#pragma unroll
            for (int w = 0; w < LGW_WARPS - 1; w++)
                if ((uint32_t)w < warp && cw != LG_DEAD) cw = sh_wmap[w][cw];     // ... and this thread's warp
            const bool owned = tile == 0u || tid >= (uint32_t)LGW_RUN;            // run-up segments belong to the tile before
            const uint32_t e = (cw == LG_DEAD || !owned) ? LG_DEAD : (uint32_t)T[LGW_ENT + cw];
            if (e != LG_DEAD) {
                if (interior) lgw_walk_segment<false>(data, seg0, e, wv0, wv1, tile_rel);
                else lgw_walk_segment<true>(data, seg0, e, wv0, wv1, tile_rel);
            }
        }
        const uint32_t c = __popc(wv0) + __popc(wv1);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        uint32_t before = 0, total = incl;
        const uint32_t errbit0 = sh_err;
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
        // ---- blocks before this tile: decoupled look-back over the count words (warp 0)
        if (warp == 0) {
            uint32_t base = 0, errbit = errbit0;
            if (tile > 0) {
                if (lane == 0) lgw_store_relaxed(cntw + 2 * (size_t)tile, lgw_pack(total, epoch, LGW_ST_AGG, errbit));
                int jhi = (int)tile - 1;
                uint32_t spins = 0;
                for (;;) {
                    const int j = jhi - (int)lane;
                    unsigned long long sw = lgw_pack(0u, epoch, LGW_ST_INCL, 0u);        // in front of tile 0: nothing
                    if (j >= 0) sw = lgw_load_relaxed(cntw + 2 * (size_t)j);
                    const uint32_t st = lgw_state(sw, epoch);
                    const unsigned inclm = __ballot_sync(0xFFFFFFFFu, st == LGW_ST_INCL);
                    const unsigned any = __ballot_sync(0xFFFFFFFFu, st != 0u);
                    const int d = inclm ? __ffs(inclm) - 1 : 32;  // nearest predecessor with an inclusive count: tile jhi - d
                    const unsigned upto = d >= 31 ? 0xFFFFFFFFu : (2u << d) - 1u;          // lanes 0 .. d (0 .. 31 if there is none)
                    if ((any & upto) == upto) {
                        uint32_t v = lane <= (uint32_t)d ? (uint32_t)(sw >> 32) : 0u;
                        uint32_t eb = lane <= (uint32_t)d ? (uint32_t)sw & LGW_ERR_BIT : 0u;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                            eb |= __shfl_xor_sync(0xFFFFFFFFu, eb, o);
                        }
                        base += v;
                        errbit |= eb;
                        if (inclm) break;
                        jhi -= 32;                                // 32 aggregates and no inclusive count yet: further back
                        continue;
                    }
                    if (++spins > LGW_SPIN_LIMIT) { errbit = LGW_ERR_BIT; break; }        // never expected
                    __nanosleep(spins < 8 ? 40 : 200);
                }
            }
            const uint32_t all = base + total;
            const bool skip = !fits || (unsigned long long)base >= need || total == 0u;   // nothing of the image starts here
            if (lane == 0) {
                lgw_store_relaxed(cntw + 2 * (size_t)tile, lgw_pack(all, epoch, LGW_ST_INCL, errbit));   // self-contained word
                if (last_tile) {
                    unsigned status = 0;
                    if (!fits) status |= MCRAW_FRAME_GEOMETRY;
                    if ((unsigned long long)all < need) status |= MCRAW_FRAME_TRUNCATED;       // reference: stale samples (:387,398)
                    if (errbit) status |= MCRAW_FRAME_INTERNAL;
                    Result r;
                    r.written = status ? 0ull : (unsigned long long)F.width * (unsigned long long)F.height;   // :494
                    r.status = status;
                    r.pad = 0;
                    results[wk.frame] = r;
                }
                sh_base = base;
                sh_skip = skip ? 1u : 0u;
            }
        }
        __syncthreads();
        if (sh_skip) continue;
        const uint32_t base = sh_base;

        // ---- 6. pair list of the tile: every block with an even ordinal leads a pair (even-column block, then odd-column
        //      block, RawData_Legacy.cpp:480-481); plist[q] = (tile-relative offset of the leader) / 2 for pair ordinal
        //      p_first + q.  Consecutive threads then decode consecutive pairs: their reads of the staged tile are a few
        //      words apart (different banks), their stores adjacent.
        const uint32_t p_first = (base + 1u) >> 1;
        uint32_t npairs = ((base + total + 1u) >> 1) - p_first;
        npairs = min(npairs, need_pairs > p_first ? need_pairs - p_first : 0u);
        const uint32_t ord0 = base + before + incl - c;          // ordinal of the first block start in this thread's segment
        const int width = F.width;
        uint16_t* __restrict__ dst = F.dst;
        const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
        const bool lin = vec && (width & 31) == 0;              // rows of whole pairs: consecutive pairs are consecutive memory
        const unsigned epi = EPI ? F.epi_mode : 0u;             // EPI = false: the epilogue code is not even in the kernel
        const EpiRegs ER = epilogue_regs(F, epi);
        // the pair leaders among this thread's block starts: marks whose ordinal is even.  x = inclusive prefix parity of the
        // marks (bit i: parity of the number of marks at or below i), so the k-th mark (k = 0, 1, ..) has x = (k + 1) & 1
        uint32_t lead0, lead1;
        {
            uint32_t x0 = wv0, x1 = wv1;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) { x0 ^= x0 << sh; x1 ^= x1 << sh; }
            const uint32_t odd0 = 0u - (ord0 & 1u), odd1 = 0u - ((ord0 + __popc(wv0)) & 1u);   // all ones: the first mark has an odd ordinal
            lead0 = wv0 & (x0 ^ odd0);
            lead1 = wv1 & (x1 ^ odd1);
        }
        for (uint32_t c0 = 0; c0 < npairs; c0 += LGW_PAIR_CHUNK) {
            const uint32_t cn = min((uint32_t)LGW_PAIR_CHUNK, npairs - c0);
            {
                uint32_t q = ((ord0 + 1u) >> 1) - p_first - c0;              // this thread's first pair; wraps to a huge value for earlier passes' pairs
#pragma unroll
                for (int hw = 0; hw < 2; hw++) {
                    uint32_t marks = hw ? lead1 : lead0;
                    while (marks) {
                        const uint32_t b = (uint32_t)__ffs((int)marks) - 1u;
                        marks &= marks - 1u;
                        if (q < cn) plist[q] = (uint16_t)(((uint32_t)LGW_SEG / 2u) * tid + 32u * hw + b);
                        q++;
                    }
                }
            }
            __syncthreads();
            if (lin) {
                // Whole rows of pairs (width = 32 * pairs per row) and an aligned buffer: pair P occupies bytes [64 P, 64 P + 64)
                // of the output, so the 32 pairs of a warp pass are 2 KiB in a row.  They go out as ONE bulk store (TMA, 1-D)
                // from a linear staging buffer; lane l writes its piece i ^ rot in round i, rot = (l / 2) mod 4, which
                // spreads a quarter warp's 16-byte stores over all 32 banks.
                uint32_t leader;                                                            // one lane of the warp owns its bulk stores
                asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(leader));
                const uint32_t slot_x = (out_s + 64u * lane) | (16u * rot);               // piece i of this lane goes to slot_x ^ 16 i
                unsigned long long gdst = reinterpret_cast<unsigned long long>(dst) + 64ull * (unsigned long long)(p_first + c0 + 32u * warp);
                for (uint32_t q0 = 32u * warp; q0 < cn; q0 += LGW_THREADS, gdst += 64ull * LGW_THREADS) {
                    if (tid == 0 && q0 + LGW_THREADS >= cn && c0 + LGW_PAIR_CHUNK >= npairs) {   // this tile's last pass
                        next_ticket = atomicAdd(&counters[2], 1u);
                        have_next = true;
                    }
                    const uint32_t q = q0 + lane;
                    uint32_t px[16];
                    if (q < cn) {
                        const uint32_t oE = 2u * (uint32_t)plist[q];
                        const uint32_t hE = leg_header(data, oE), bitsE = leg_hdr_bits(hE);
                        const uint32_t oO = oE + 2u + leg_len(bitsE);
                        const uint32_t hO = leg_header(data, oO), bitsO = leg_hdr_bits(hO);
                        const uint32_t mE = lgw_block<false>(d32, oE, bitsE, px, rot1, rot2, rot);
                        const uint32_t mO = lgw_block<true>(d32, oO, bitsO, px, rot1, rot2, rot);
                        const uint32_t cm = mE | (mO << 16);
                        const uint32_t refs = leg_hdr_ref(hE) | (leg_hdr_ref(hO) << 16);
#pragma unroll
                        for (int i = 0; i < 16; i++) px[i] = __vadd2(px[i] & cm, refs);           // :483-486, + reference mod 2^16
                        if (epi) {                                                                // optional black / white level epilogue
                            const uint32_t y = (p_first + c0 + q) / ppr;
                            const EpiRow R = epilogue_row(ER, (y & 1u) != 0u);
#pragma unroll
                            for (int i = 0; i < 16; i++) px[i] = epilogue_word(px[i], epi, R);
                        }
                    }
                    // lane 0: the last store has read the buffer
                    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
                    __syncwarp();
                    if (q < cn) {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n"
                                         ::"r"(slot_x ^ (16u * (uint32_t)i)), "r"(px[4 * i]), "r"(px[4 * i + 1]), "r"(px[4 * i + 2]), "r"(px[4 * i + 3]) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");               // generic stores before the async read
                    __syncwarp();
                    if (leader) {
                        const uint32_t nb = 64u * min(32u, cn - q0);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
                                     "cp.async.bulk.commit_group;\n" ::"l"(gdst), "r"(out_s), "r"(nb) : "memory");
                    }
                }
                if (leader) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // before anyone reuses the tables
            } else {
                uint32_t P = p_first + c0 + tid;
                uint32_t y = P / ppr, xq = P - y * ppr;
                for (uint32_t q = tid; q < cn; q += LGW_THREADS) {
                    if (tid == 0 && q + LGW_THREADS >= cn && c0 + LGW_PAIR_CHUNK >= npairs) {     // this tile's last pass
                        next_ticket = atomicAdd(&counters[2], 1u);
                        have_next = true;
                    }
                    const uint32_t oE = 2u * (uint32_t)plist[q];
                    const uint32_t hE = leg_header(data, oE), bitsE = leg_hdr_bits(hE);
                    const uint32_t oO = oE + 2u + leg_len(bitsE);
                    const uint32_t hO = leg_header(data, oO), bitsO = leg_hdr_bits(hO);
                    uint32_t px[16];
                    const uint32_t mE = lgw_block<false>(d32, oE, bitsE, px, false, false, 0u);
                    const uint32_t mO = lgw_block<true>(d32, oO, bitsO, px, false, false, 0u);
                    const uint32_t cm = mE | (mO << 16);
                    const uint32_t refs = leg_hdr_ref(hE) | (leg_hdr_ref(hO) << 16);
#pragma unroll
                    for (int i = 0; i < 16; i++) px[i] = __vadd2(px[i] & cm, refs);               // :483-486, + reference mod 2^16
                    if (epi) {                                                                    // optional black / white level epilogue
                        const EpiRow R = epilogue_row(ER, (y & 1u) != 0u);
#pragma unroll
                        for (int i = 0; i < 16; i++) px[i] = epilogue_word(px[i], epi, R);
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
            }
            __syncthreads();
        }
    }
    // the last CTA to leave resets the ticket counters for the next launch
    if (tid == 0 && atomicAdd(&counters[3], 1u) == gridDim.x - 1u) { counters[2] = 0; counters[3] = 0; }
}

}  // namespace mcraw
