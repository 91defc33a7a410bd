// mcraw_legacy.cuh -- sm_100a kernels for the legacy frame format (compressionType 6).
//
// Reference: /root/reference/lib/RawData_Legacy.cpp:445-495 (raw::DecodeLegacy) and :372-442 (DecodeHeader /
// DecodeBlock).  A frame is one chain of 16-sample blocks, each with an inline 2-byte header
// (bits nibble, 12-bit reference) followed by 2*bits payload bytes (32 for nibbles 11..15): block k+1 starts
// where block k ends, so the reference finds the blocks with 750 000 dependent steps per 4000x3000 frame.
//
// Here the chain is resolved in parallel.  All block lengths are even and <= 34 bytes, so the true chain enters
// any fixed byte SEGMENT at one of 17 even offsets (0, 2, ..., 32):
//
//   k_legacy_maps    one CTA per TILE of 32 segments of 1 KiB.  The tile is staged in shared memory; for every
//                    segment, 17 lanes walk the 17 candidate chains and record the TRANSFER MAP
//                    entry offset -> (exit offset into the next segment, blocks started, chain died).
//                    The CTA then composes its 32 segment maps into one tile map.
//   k_legacy_scan    one CTA per frame: composes the tile maps front to back from entry offset 0 (serial, but only
//                    len / 32 KiB steps in shared memory) -> entry offset and first block ordinal of every tile;
//                    checks that the chain holds the 2 * (paddedWidth / 32) * height blocks the frame needs
//                    (RawData_Legacy.cpp:478-482) and writes the per-frame result.
//   k_legacy_decode  one CTA per tile: re-stages the tile, resolves the entry of each of its segments from the
//                    segment maps, marks the block starts of every segment in a shared-memory bitmap (one lane per
//                    segment), then every lane decodes block PAIRS (even-column block + odd-column block,
//                    :480-481): MSB-first bit extraction with funnel shifts (:38-370), + reference mod 2^16,
//                    column interleave (:483-486) in registers, 16-byte stores, crop at width (:490).
#pragma once
#include "mcraw_kernels.cuh"

namespace mcraw {

constexpr int LG_SEG = 1024;                 // bytes per segment
constexpr int LG_SLOTS = LG_SEG / 2;         // candidate (even) block starts per segment
constexpr int LG_TILE_SEGS = 32;             // segments per tile
constexpr int LG_TILE = LG_SEG * LG_TILE_SEGS;
constexpr int LG_STATES = 17;                // entry offsets 0, 2, ..., 32
constexpr uint32_t LG_DEAD = 31;             // exit code of a chain that ran into the end of the buffer
constexpr int LG_THREADS = 256;
constexpr int LG_OVERRUN = 80;               // a pair led inside the tile ends at most 2 + 34 + 34 bytes past it

// payload bytes of a 16-sample block for header nibble b (RawData_Legacy.cpp:13-32, min(16, bits) at :395)
__device__ __forceinline__ uint32_t leg_len(uint32_t b) { return b <= 10u ? 2u * b : 32u; }

// stage [tile_off, tile_off + nbytes) of the frame into shared memory, zero past len (16-byte granules)
__device__ __forceinline__ void lg_stage(uint8_t* sm, const uint8_t* __restrict__ src, unsigned long long len,
                                         unsigned long long tile_off, int nbytes, int tid) {
    if (tile_off + (unsigned long long)nbytes <= len) {                       // the common case: wholly inside the buffer
        const uint4* g = reinterpret_cast<const uint4*>(src + tile_off);
        uint4* d = reinterpret_cast<uint4*>(sm);
#pragma unroll 4
        for (int v = tid; v < nbytes / 16; v += LG_THREADS) d[v] = __ldg(g + v);
        return;
    }
    for (int v = tid; v < nbytes / 16; v += LG_THREADS) {
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

// Walk a chain from p to its first block start at or beyond `stop`.  CHECK: the segment lies near the end of the buffer,
// so a block may fail the reference's bound (RawData_Legacy.cpp:387,398: decoded only if offset + 2 + payload < len).
template <bool CHECK>
__device__ __forceinline__ void lg_walk(const uint8_t* seg, uint32_t& p, uint32_t& cnt, bool& dead, const uint32_t stop,
                                        const uint32_t rel_len) {
    while (p < stop) {
        const uint32_t b = (uint32_t)seg[p] >> 4;
        const uint32_t q = p + 2u + (b > 10u ? 32u : 2u * b);
        if (CHECK && q >= rel_len) { dead = true; break; }
        p = q;
        cnt++;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_legacy_maps: grid = (max tiles, frames), block = LG_THREADS, dynamic smem = LG_MAPS_SMEM
//
// Phase 1 (a warp per segment, 17 lanes): the 17 candidate chains are walked only until they have MERGED -- chains
//   that meet at one block start are the same chain from there on, and with blocks of varying length that happens
//   within a few blocks.  Merging is tested at barrier lines 64, 128, 256, 512 bytes into the segment: every lane
//   walks to its first block start at or beyond the line; if all 17 agree on it (X), the segment's map is
//   "entry e -> pre[e] blocks, then the common chain from X".  Chains that have not merged by 512 (constant-width
//   regions) are simply walked to the end of the segment: the full map, no shortcut.
// Phase 2 (one lane per segment): the common chain from X to the end of the segment, marking every block start in a
//   bitmap that k_legacy_decode reuses, so nothing past X is ever walked twice or 17-fold.
// ---------------------------------------------------------------------------------------------------------
constexpr int LG_BM_WORDS = LG_SLOTS / 32;                  // bitmap words per segment
constexpr uint32_t LG_NO_X = 0xFFFFu;                       // segment has no merge point (full map, no bitmap)
constexpr int LG_MAPS_SMEM = LG_TILE + LG_TILE_SEGS * 18 * 2 + LG_TILE_SEGS * LG_BM_WORDS * 4 + LG_TILE_SEGS * 2;

__global__ void __launch_bounds__(LG_THREADS) k_legacy_maps(const FrameDev* __restrict__ frames) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY) return;
    const unsigned long long len = F.len;
    const uint32_t nseg = (uint32_t)((len + LG_SEG - 1) / LG_SEG);
    const uint32_t tile = blockIdx.x;
    if ((unsigned long long)tile * LG_TILE_SEGS >= nseg) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* data = lg_smem;
    uint16_t* maps = reinterpret_cast<uint16_t*>(lg_smem + LG_TILE);                                  // [32][18]
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LG_TILE + LG_TILE_SEGS * 36);            // [32][16]
    uint16_t* segx = reinterpret_cast<uint16_t*>(lg_smem + LG_TILE + LG_TILE_SEGS * 36 + LG_TILE_SEGS * LG_BM_WORDS * 4);
    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    lg_stage(data, F.src, len, tile_off, LG_TILE, tid);
    for (int i = tid; i < LG_TILE_SEGS * LG_BM_WORDS; i += LG_THREADS) bitmap[i] = 0;
    __syncthreads();

    const uint32_t segs_here = min((uint32_t)LG_TILE_SEGS, nseg - tile * LG_TILE_SEGS);
    // ---- phase 1
    for (uint32_t s = warp; s < segs_here; s += LG_THREADS / 32) {
        const unsigned long long seg_abs = tile_off + (unsigned long long)s * LG_SEG;
        // RawData_Legacy.cpp:387,398: a block is decoded only if offset + 2 + payload < len; as a segment-relative bound
        const uint32_t rel_len = (uint32_t)min(len - seg_abs, (unsigned long long)(1u << 30));
        const uint8_t* seg = data + s * LG_SEG;
        const bool active = lane < LG_STATES;
        const bool check = rel_len < (uint32_t)(LG_SEG + 40);                     // warp-uniform: only the last segments of a frame
        uint32_t p = active ? 2u * lane : 0xFFFF0000u, cnt = 0;                   // idle lanes never enter the walk
        bool dead = false;
        uint32_t x = LG_NO_X;
        // merge test at 64, 96, ..., 256, 320, ..., 512 bytes; chains still apart at 512 are walked to the end
        for (uint32_t line = 64;; line += (line < 256u ? 32u : 64u)) {
            const uint32_t stop = line > 512u ? (uint32_t)LG_SEG : line;
            if (check) { if (!dead) lg_walk<true>(seg, p, cnt, dead, stop, rel_len); }
            else lg_walk<false>(seg, p, cnt, dead, stop, rel_len);
            if (stop == (uint32_t)LG_SEG) break;                                  // walked to the end: full map
            const uint32_t p0 = __shfl_sync(0xFFFFFFFFu, p, 0);
            const bool agree = __all_sync(0xFFFFFFFFu, !active || (!dead && p == p0));
            if (agree && p0 < (uint32_t)LG_SEG) { x = p0; break; }
        }
        if (active) {
            if (x == LG_NO_X) maps[s * 18 + lane] = (uint16_t)((dead ? LG_DEAD : (p - LG_SEG) >> 1) | (cnt << 5));
            else maps[s * 18 + lane] = (uint16_t)cnt;                             // pre[e]; completed in phase 2
        }
        if (lane == 0) segx[s] = (uint16_t)x;
    }
    __syncthreads();
    // ---- phase 2
    if (warp == 0 && (uint32_t)lane < segs_here && segx[lane] != LG_NO_X) {
        const uint32_t s = lane;
        const unsigned long long seg_abs = tile_off + (unsigned long long)s * LG_SEG;
        const uint32_t rel_len = (uint32_t)min(len - seg_abs, (unsigned long long)(1u << 30));
        const uint8_t* seg = data + s * LG_SEG;
        uint32_t* bm = bitmap + s * LG_BM_WORDS;
        uint32_t p = segx[s], cnt = 0;
        bool dead = false;
        // one bitmap word covers 64 bytes of the segment: walk word by word, collecting the starts in a register
        for (uint32_t wd = p >> 6; wd < (uint32_t)LG_BM_WORDS && !dead; wd++) {
            uint32_t acc = 0;
            const uint32_t stop = 64u * (wd + 1u);
            while (p < stop) {
                const uint32_t b = (uint32_t)seg[p] >> 4;
                const uint32_t q = p + 2u + (b > 10u ? 32u : 2u * b);
                if (q >= rel_len) { dead = true; break; }
                acc |= 1u << ((p >> 1) & 31u);
                p = q;
                cnt++;
            }
            bm[wd] = acc;
        }
        const uint32_t ex = dead ? LG_DEAD : (p - LG_SEG) >> 1;
        for (int e = 0; e < LG_STATES; e++) maps[s * 18 + e] = (uint16_t)(ex | (((uint32_t)maps[s * 18 + e] + cnt) << 5));
    }
    __syncthreads();
    // ---- results of the tile: segment maps, merge points and bitmaps for k_legacy_decode; the composed tile map
    const size_t seg0 = (size_t)tile * LG_TILE_SEGS;
    for (uint32_t i = tid; i < segs_here * LG_STATES; i += LG_THREADS) {
        const uint32_t s = i / LG_STATES, e = i - s * LG_STATES;
        F.lg_segmap[seg0 * LG_STATES + i] = maps[s * 18 + e];
    }
    for (uint32_t i = tid; i < segs_here * LG_BM_WORDS; i += LG_THREADS) F.lg_bitmap[seg0 * LG_BM_WORDS + i] = bitmap[i];
    if ((uint32_t)tid < segs_here) F.lg_segx[seg0 + tid] = segx[tid];
    if (warp == 0 && lane < LG_STATES) {
        uint32_t state = lane, total = 0;
        for (uint32_t s = 0; s < segs_here; s++) {
            const uint32_t m = maps[s * 18 + state];
            total += m >> 5;
            state = m & 31u;
            if (state == LG_DEAD) break;
        }
        // running off the end of the buffer without meeting an undecodable block also ends the chain
        if (state != LG_DEAD && tile * LG_TILE_SEGS + segs_here == nseg) state = LG_DEAD;
        F.lg_tilemap[(size_t)tile * LG_STATES + lane] = state | (total << 5);
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_legacy_scan: grid = frames, block = LG_THREADS
// ---------------------------------------------------------------------------------------------------------
constexpr int LG_SCAN_TILES = 512;    // tile maps staged per round

__global__ void __launch_bounds__(LG_THREADS) k_legacy_scan(const FrameDev* __restrict__ frames, FrameState* __restrict__ states,
                                                            Result* __restrict__ results) {
    __shared__ uint32_t tm[LG_SCAN_TILES * LG_STATES];
    __shared__ uint32_t sh_state, sh_base;
    const FrameDev& F = frames[blockIdx.x];
    if (F.type != MCRAW_COMPRESSION_LEGACY) return;
    const int tid = threadIdx.x;
    const unsigned long long len = F.len;
    const uint32_t nseg = (uint32_t)((len + LG_SEG - 1) / LG_SEG);
    const uint32_t ntile = (nseg + LG_TILE_SEGS - 1) / LG_TILE_SEGS;
    const unsigned long long ppr = ((unsigned long long)F.width + 31ull) / 32ull;          // RawData_Legacy.cpp:34-36,449
    const unsigned long long need = 2ull * ppr * (unsigned long long)F.height;              // :478-482
    unsigned status = 0;
    if (len == 0) status = MCRAW_FRAME_TRUNCATED;
    if (!status && F.dst_cap < (unsigned long long)F.width * (unsigned long long)F.height) status = MCRAW_FRAME_GEOMETRY;
    if (tid == 0) { sh_state = 0; sh_base = 0; }
    __syncthreads();
    if (!status) {
        for (uint32_t t0 = 0; t0 < ntile; t0 += LG_SCAN_TILES) {
            const uint32_t nt = min((uint32_t)LG_SCAN_TILES, ntile - t0);
            for (uint32_t i = tid; i < nt * LG_STATES; i += LG_THREADS) tm[i] = F.lg_tilemap[(size_t)t0 * LG_STATES + i];
            __syncthreads();
            if (tid == 0) {
                uint32_t state = sh_state, base = sh_base;
                for (uint32_t t = 0; t < nt; t++) {
                    F.lg_tilestate[2 * (size_t)(t0 + t)] = state;
                    F.lg_tilestate[2 * (size_t)(t0 + t) + 1] = base;
                    if (state != LG_DEAD) {
                        const uint32_t m = tm[t * LG_STATES + state];
                        base += m >> 5;
                        state = m & 31u;
                    }
                }
                sh_state = state; sh_base = base;
            }
            __syncthreads();
        }
        if ((unsigned long long)sh_base < need) status = MCRAW_FRAME_TRUNCATED;   // reference: stale samples (:387,398)
    }
    if (tid == 0) {
        states[blockIdx.x].status[0] = status;
        states[blockIdx.x].status[1] = 0;
        Result r;
        r.written = status ? 0ull : (unsigned long long)F.width * (unsigned long long)F.height;   // :494
        r.status = status;
        r.pad = 0;
        results[blockIdx.x] = r;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_legacy_decode
// ---------------------------------------------------------------------------------------------------------
// 16 samples of W bits each, MSB-first contiguous (RawData_Legacy.cpp:38-358).  w: the staged tile as words, pi / psh:
// word index and bit shift of the first payload byte (the payload starts 2-byte aligned).
template <int W>
__device__ __forceinline__ void leg_unpack(const uint32_t* w, const uint32_t pi, const uint32_t psh, uint32_t (&v)[16]) {
    constexpr int NW = (W + 1) / 2;                   // payload words: 2 * W bytes
    uint32_t be[NW + 1];
    uint32_t lo = w[pi];
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t hi = w[pi + k + 1];
        be[k] = __byte_perm(__funnelshift_r(lo, hi, psh), 0u, 0x0123);            // big-endian view of payload bytes 4k .. 4k+3
        lo = hi;
    }
    be[NW] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int bit = k * W, i = bit >> 5, sh = bit & 31;
        uint32_t x;
        if (sh + W <= 32) x = be[i] >> (32 - sh - W);
        else x = __funnelshift_l(be[i + 1], be[i], sh) >> (32 - W);
        v[k] = x & ((1u << W) - 1u);
    }
}

// Decode the block whose header sits at byte offset o (even) of the staged tile: returns its total length.
__device__ __forceinline__ uint32_t leg_block(const uint8_t* data, uint32_t o, uint32_t (&v)[16], uint32_t& ref) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(data);
    const uint32_t i0 = o >> 2, sh = (o & 2u) * 8u;
    const uint32_t h = __funnelshift_r(w[i0], w[i0 + 1], sh);                    // bytes o .. o+3
    const uint32_t bits = (h >> 4) & 15u;                                        // RawData_Legacy.cpp:372-375
    ref = ((h & 15u) << 8) | ((h >> 8) & 0xFFu);
    const uint32_t pi = (o + 2u) >> 2, psh = ((o + 2u) & 2u) * 8u;
    switch (bits) {
    case 0:
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = 0;                                   // :402-404
        return 2u;
    case 1: leg_unpack<1>(w, pi, psh, v); return 4u;
    case 2: leg_unpack<2>(w, pi, psh, v); return 6u;
    case 3: leg_unpack<3>(w, pi, psh, v); return 8u;
    case 4: leg_unpack<4>(w, pi, psh, v); return 10u;
    case 5: leg_unpack<5>(w, pi, psh, v); return 12u;
    case 6: leg_unpack<6>(w, pi, psh, v); return 14u;
    case 7: leg_unpack<7>(w, pi, psh, v); return 16u;
    case 8: leg_unpack<8>(w, pi, psh, v); return 18u;
    case 9: leg_unpack<9>(w, pi, psh, v); return 20u;
    case 10: leg_unpack<10>(w, pi, psh, v); return 22u;
    default: leg_unpack<16>(w, pi, psh, v); return 34u;                          // 11..15 -> 16-bit big-endian (:360-370,395)
    }
}

constexpr int LG_DEC_DATA = LG_TILE + LG_OVERRUN;
constexpr int LG_MAX_PAIRS = LG_TILE / 4;             // a pair is at least two 2-byte blocks
constexpr int LG_DEC_SMEM = LG_DEC_DATA + LG_TILE_SEGS * 18 * 2 /*maps*/ + LG_TILE_SEGS * LG_BM_WORDS * 4 /*bitmaps*/ +
                            LG_MAX_PAIRS * 2 /*pair list*/ + (LG_TILE_SEGS + 1) * 8 /*entry, base*/ + LG_TILE_SEGS * 2 /*merge points*/;

__global__ void __launch_bounds__(LG_THREADS, 4) k_legacy_decode(const FrameDev* __restrict__ frames, const FrameState* __restrict__ states) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY || states[blockIdx.y].status[0]) return;
    const unsigned long long len = F.len;
    const uint32_t nseg = (uint32_t)((len + LG_SEG - 1) / LG_SEG);
    const uint32_t tile = blockIdx.x;
    if ((unsigned long long)tile * LG_TILE_SEGS >= nseg) return;
    const uint32_t tile_entry = F.lg_tilestate[2 * (size_t)tile];
    const uint32_t tile_base = F.lg_tilestate[2 * (size_t)tile + 1];
    const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;                            // pairs per row (RawData_Legacy.cpp:34-36)
    const unsigned long long need = 2ull * ppr * (unsigned long long)F.height;       // blocks of the image (:478-482), < 2^33
    if (tile_entry == LG_DEAD || (unsigned long long)tile_base >= need) return;      // nothing of the image starts here

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* data = lg_smem;
    uint16_t* maps = reinterpret_cast<uint16_t*>(lg_smem + LG_DEC_DATA);                          // [32][18]
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LG_DEC_DATA + LG_TILE_SEGS * 36);    // [32][16]
    uint16_t* plist = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(bitmap) + LG_TILE_SEGS * LG_BM_WORDS * 4);
    uint32_t* seg_entry = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(plist) + LG_MAX_PAIRS * 2);   // [33]
    uint32_t* seg_base = seg_entry + LG_TILE_SEGS + 1;                                            // [33]: [32] = end of the tile
    uint16_t* segx = reinterpret_cast<uint16_t*>(seg_base + LG_TILE_SEGS + 1);

    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    const uint32_t segs_here = min((uint32_t)LG_TILE_SEGS, nseg - tile * LG_TILE_SEGS);
    lg_stage(data, F.src, len, tile_off, LG_DEC_DATA, tid);
    for (uint32_t i = tid; i < segs_here * LG_STATES; i += LG_THREADS) {
        const uint32_t s = i / LG_STATES, e = i - s * LG_STATES;
        maps[s * 18 + e] = F.lg_segmap[((size_t)tile * LG_TILE_SEGS + s) * LG_STATES + e];
    }
    // block starts past each segment's merge point come from k_legacy_maps
    for (uint32_t i = tid; i < LG_TILE_SEGS * LG_BM_WORDS; i += LG_THREADS)
        bitmap[i] = i < segs_here * LG_BM_WORDS ? F.lg_bitmap[(size_t)tile * LG_TILE_SEGS * LG_BM_WORDS + i] : 0u;
    if ((uint32_t)tid < LG_TILE_SEGS) segx[tid] = (uint32_t)tid < segs_here ? F.lg_segx[(size_t)tile * LG_TILE_SEGS + tid] : (uint16_t)LG_NO_X;
    __syncthreads();
    // ---- entry offset and first block ordinal of every segment of the tile
    if (tid == 0) {
        uint32_t state = tile_entry, base = tile_base;
        for (uint32_t s = 0; s < LG_TILE_SEGS; s++) {
            seg_entry[s] = s < segs_here ? state : LG_DEAD;
            seg_base[s] = base;
            if (s < segs_here && state != LG_DEAD) {
                const uint32_t m = maps[s * 18 + state];
                base += m >> 5;
                state = m & 31u;
            }
        }
        seg_base[LG_TILE_SEGS] = base;
    }
    __syncthreads();
    // ---- one lane per segment: walk the (now known) chain from its entry to the merge point -- a few blocks; the whole
    //      segment only where the candidate chains never merged -- and add those block starts to the bitmap
    if (warp == 0) {
        const uint32_t s = lane;
        const uint32_t e = seg_entry[s];
        if (e != LG_DEAD) {
            const unsigned long long seg_abs = tile_off + (unsigned long long)s * LG_SEG;
            const uint32_t rel_len = (uint32_t)min(len - seg_abs, (unsigned long long)(1u << 30));
            const uint8_t* seg = data + s * LG_SEG;
            uint32_t* bm = bitmap + s * LG_BM_WORDS;
            const uint32_t stop = segx[s] == LG_NO_X ? (uint32_t)LG_SEG : (uint32_t)segx[s];
            uint32_t p = 2u * e;
            while (p < stop) {
                const uint32_t q = p + 2u + leg_len(seg[p] >> 4);
                if (q >= rel_len) break;
                bm[p >> 6] |= 1u << ((p >> 1) & 31u);
                p = q;
            }
        }
    }
    __syncthreads();
    // ---- pair list of the tile: every block with an even ordinal leads a pair (even-column block, then odd-column block,
    //      RawData_Legacy.cpp:480-481); plist[q] = (tile-relative offset of the leader) / 2 for pair ordinal p_first + q
    const uint32_t p_first = (tile_base + 1u) >> 1;
    uint32_t npairs = ((seg_base[LG_TILE_SEGS] + 1u) >> 1) - p_first;
    npairs = (uint32_t)min((unsigned long long)npairs, (need >> 1) - (unsigned long long)p_first);
    for (uint32_t s = warp; s < segs_here; s += LG_THREADS / 32) {
        if (seg_entry[s] == LG_DEAD) break;
        const uint32_t wordv = lane < (uint32_t)LG_BM_WORDS ? bitmap[s * LG_BM_WORDS + lane] : 0u;
        const uint32_t c = __popc(wordv);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < LG_BM_WORDS; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        uint32_t ord = seg_base[s] + incl - c;               // ordinal of the first block start in this lane's word
        uint32_t wv = wordv;
        while (wv) {
            const uint32_t b = __ffs(wv) - 1;
            wv &= wv - 1;
            const uint32_t q = (ord >> 1) - p_first;
            if (!(ord & 1u) && q < npairs) plist[q] = (uint16_t)(s * (LG_SEG / 2) + 32u * lane + b);
            ord++;
        }
    }
    __syncthreads();
    // ---- decode: a lane takes one block pair at a time -> 32 consecutive pixels
    const int width = F.width;
    uint16_t* __restrict__ dst = F.dst;
    const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
    uint32_t P = p_first + (uint32_t)tid;
    uint32_t y = P / ppr, xq = P - y * ppr;
    for (uint32_t q = tid; q < npairs; q += LG_THREADS) {
        const uint32_t o = 2u * (uint32_t)plist[q];
        uint32_t vE[16], vO[16], refE, refO;
        const uint32_t lenE = leg_block(data, o, vE, refE);
        leg_block(data, o + lenE, vO, refO);
        const uint32_t refs = refE | (refO << 16);
        const int x = (int)(32u * xq);
        uint32_t px[16];
#pragma unroll
        for (int k = 0; k < 16; k++) px[k] = __vadd2(vE[k] | (vO[k] << 16), refs);    // :483-486, + reference mod 2^16
        uint16_t* orow = dst + (size_t)y * (size_t)width + x;
        if (vec && x + 32 <= width) {
            uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
            for (int k = 0; k < 4; k++) o4[k] = make_uint4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) {                                            // crop at width (:490)
                if (x + 2 * k < width) orow[2 * k] = (uint16_t)px[k];
                if (x + 2 * k + 1 < width) orow[2 * k + 1] = (uint16_t)(px[k] >> 16);
            }
        }
        xq += LG_THREADS;                                                             // the pair LG_THREADS further on
        while (xq >= ppr) { xq -= ppr; y++; }
    }
}

}  // namespace mcraw
