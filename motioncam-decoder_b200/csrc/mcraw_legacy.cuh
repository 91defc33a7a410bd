// mcraw_legacy.cuh -- sm_100a kernels for the legacy frame format (compressionType 6).
//
// Reference: /root/reference/lib/RawData_Legacy.cpp:445-495 (raw::DecodeLegacy) and :372-442 (DecodeHeader /
// DecodeBlock).  A frame is one chain of 16-sample blocks, each with an inline 2-byte header
// (bits nibble, 12-bit reference) followed by 2*bits payload bytes (32 for nibbles 11..15): block k+1 starts
// where block k ends, so the reference finds the blocks with 750 000 dependent steps per 4000x3000 frame.
//
// Here the chain is resolved in parallel.  Two facts carry it:
//   (1) all block lengths are even and <= 34 bytes, so the chain enters any fixed byte range at one of 17 even offsets;
//   (2) two chains that ever share a block start are identical from there on, and with blocks of varying length chains
//       started anywhere MERGE within a few blocks.  Nothing relies on (2) for correctness -- only for speed.
//
//   k_legacy_maps    one CTA per TILE of 32 KiB, staged in shared memory.  One lane per 1 KiB segment walks a GUESSED
//                    chain (from the start of its segment) and marks its block starts in a bitmap; then every lane
//                    re-walks from where its left neighbour's chain actually ends, only until it meets its own marks
//                    (self-synchronisation), repeated until no entry changes.  The result is the exact chain C0 of the
//                    tile for tile entry offset 0.  For the other 16 possible entry offsets, 16 lanes walk until they
//                    meet C0: entry e -> (merge point, blocks before it).  Output: the tile's transfer map
//                    entry -> (exit offset, block count), C0's bitmap, the merge points.
//   k_legacy_scan    one CTA per frame: composes the tile maps front to back from entry offset 0 (serial, but only
//                    len / 32 KiB steps in shared memory) -> entry offset and first block ordinal of every tile;
//                    checks that the chain holds the 2 * (paddedWidth / 32) * height blocks the frame needs
//                    (RawData_Legacy.cpp:478-482), writes the per-frame result, and (one thread per tile) patches
//                    each tile's bitmap for its true entry: the few blocks before the merge point.
//   k_legacy_decode  one CTA per tile: stages the tile and its bitmap, turns the bitmap into the list of block PAIRS
//                    (even-column block + odd-column block, :480-481) with prefix popcounts, then every lane decodes one
//                    pair at a time: MSB-first bit extraction with funnel shifts (:38-370), + reference mod 2^16,
//                    column interleave (:483-486) in registers, 16-byte stores, crop at width (:490).
#pragma once
#include "mcraw_kernels.cuh"

namespace mcraw {

constexpr int LG_SEG = 512;                  // bytes per segment (one lane of the index warp)
constexpr int LG_TILE_SEGS = 32;             // segments per tile
constexpr int LG_TILE = LG_SEG * LG_TILE_SEGS;
constexpr int LG_TILE_SLOTS = LG_TILE / 2;   // candidate (even) block starts per tile
constexpr int LG_TILE_WORDS = LG_TILE_SLOTS / 32;   // bitmap words per tile
constexpr int LG_SEG_WORDS = LG_SEG / 64;    // bitmap words per segment
constexpr int LG_STATES = 17;                // entry offsets 0, 2, ..., 32
constexpr uint32_t LG_DEAD = 31;             // exit code of a chain that ran into the end of the buffer
constexpr uint32_t LG_NO_MERGE = 0xFFFFu;    // merge point of an entry whose chain never meets C0 inside the tile
constexpr uint32_t LG_SLOW = 0x100u;         // tile state flag: k_legacy_decode has to walk the tile itself
constexpr int LG_THREADS = 128;
constexpr int LG_OVERRUN = 80;               // a pair led inside the tile ends at most 2 + 34 + 34 bytes past it

// payload bytes of a 16-sample block for header nibble b (RawData_Legacy.cpp:13-32, min(16, bits) at :395)
__device__ __forceinline__ uint32_t leg_len(uint32_t b) { return b <= 10u ? 2u * b : 32u; }

// stage [tile_off, tile_off + nbytes) of the frame into shared memory, zero past len (16-byte granules)
template <int NT>
__device__ __forceinline__ void lg_stage(uint8_t* sm, const uint8_t* __restrict__ src, unsigned long long len,
                                         unsigned long long tile_off, int nbytes, int tid) {
    if (tile_off + (unsigned long long)nbytes <= len) {                       // the common case: wholly inside the buffer
        const uint4* g = reinterpret_cast<const uint4*>(src + tile_off);
        uint4* d = reinterpret_cast<uint4*>(sm);
        int v = tid;
        for (; v + 7 * NT < nbytes / 16; v += 8 * NT) {                        // eight loads in flight per thread
            uint4 q[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = __ldg(g + v + k * NT);
#pragma unroll
            for (int k = 0; k < 8; k++) d[v + k * NT] = q[k];
        }
        for (; v < nbytes / 16; v += NT) d[v] = __ldg(g + v);
        return;
    }
    for (int v = tid; v < nbytes / 16; v += NT) {
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

// total length of the block whose header byte is hb (RawData_Legacy.cpp:13-32,395)
__device__ __forceinline__ uint32_t leg_step(uint32_t hb) {
    const uint32_t b = hb >> 4;
    return 2u + (b > 10u ? 32u : 2u * b);
}

// ---------------------------------------------------------------------------------------------------------
// k_legacy_maps: grid = (max tiles, frames), block = LG_MAPS_THREADS, dynamic smem = LG_MAPS_SMEM
// ---------------------------------------------------------------------------------------------------------
constexpr int LG_MAPS_THREADS = 128;
// The walks never look at payload bytes, only at "how far is the next block if a block started here": the tile is turned
// into a table of half step lengths (1..17), one byte per even offset, while it passes through registers on its way in.
constexpr int LG_NX_SEG = LG_SEG / 2;          // table bytes per segment
constexpr int LG_NX_PITCH = LG_NX_SEG + 8;     // segments 8 bytes apart: lane s walks segment s, and at a pitch of LG_NX_SEG all
                                               // 32 lanes would hit the same bank on every hop
constexpr int LG_MAPS_DATA = LG_TILE_SEGS * LG_NX_PITCH;
constexpr int LG_MAPS_SMEM = LG_MAPS_DATA + LG_TILE_WORDS * 4;

// half step lengths of the eight even offsets of 16 stream bytes (RawData_Legacy.cpp:13-32,395), packed low byte first
__device__ __forceinline__ uint2 lg_half_steps(const uint4 q) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t nib = (w[i] >> (16 * h + 4)) & 15u;
            const uint32_t half = 1u + (nib > 10u ? 16u : nib);
            o[i >> 1] |= half << (8 * (2 * (i & 1) + h));
        }
    }
    return make_uint2(o[0], o[1]);
}

// One lane, one segment (nx: its table of half step lengths): walk from byte offset p (segment-relative) to the end of the segment.  Block starts go into
// bm[] (one word per 64 bytes).  With MERGE, the walk stops at the first position that is already marked in bm[] --
// from there on the old marks are this chain's own -- and the old marks before that position are dropped.
// rel: bytes from the segment start to the end of the buffer (RawData_Legacy.cpp:387,398: a block is decoded only if
// offset + 2 + payload < len).  Returns true if the chain ended at an undecodable block.
template <bool MERGE>
__device__ __forceinline__ bool lg_walk_segment(const uint8_t* nx, uint32_t& p, uint32_t (&bm)[LG_SEG_WORDS], const uint32_t rel) {
    bool merged = false, dead = false;
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) {
        if (merged) continue;                                  // the old marks from the merge point on are this chain's
        const uint32_t stop = 64u * (wd + 1);
        if (dead || p >= stop) { bm[wd] = 0; continue; }       // before the entry, or after the chain ended: no block starts
        const uint32_t old = bm[wd];
        uint32_t acc = 0;
        while (p < stop) {
            const uint32_t bit = 1u << ((p >> 1) & 31u);
            if (MERGE && (old & bit)) { merged = true; acc |= old & ~(bit - 1u); break; }
            const uint32_t q = p + 2u * nx[p >> 1];
            if (q >= rel) { dead = true; break; }
            acc |= bit;
            p = q;
        }
        bm[wd] = acc;
    }
    return dead;
}

__global__ void __launch_bounds__(LG_MAPS_THREADS) k_legacy_maps(const FrameDev* __restrict__ frames) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY) return;
    const unsigned long long len = F.len;
    const uint32_t ntile = (uint32_t)((len + LG_TILE - 1) / LG_TILE);
    const uint32_t tile = blockIdx.x;
    if (tile >= ntile) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* data = lg_smem;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LG_MAPS_DATA);                           // [LG_TILE_WORDS]
    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    const uint32_t tile_rel = (uint32_t)min(len - tile_off, (unsigned long long)(1u << 30));        // bytes to the end of the buffer
    // the tile passes through registers: only its step table lands in shared memory (segment s at s * LG_NX_PITCH)
    if (tile_off + LG_TILE <= len) {
        const uint4* g = reinterpret_cast<const uint4*>(F.src + tile_off);
        constexpr int PER = LG_TILE / 16 / LG_MAPS_THREADS;
        uint4 q[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) q[k] = __ldg(g + tid + k * LG_MAPS_THREADS);
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int v = tid + k * LG_MAPS_THREADS;
            *reinterpret_cast<uint2*>(data + 8 * v + 8 * (v / (LG_SEG / 16))) = lg_half_steps(q[k]);
        }
    } else {
        for (int v = tid; v < LG_TILE / 16; v += LG_MAPS_THREADS) {
            const unsigned long long o = tile_off + 16ull * (unsigned)v;
            uint32_t t4[4] = {0, 0, 0, 0};
            for (int k = 0; k < 16; k++)
                if (o + k < len) t4[k >> 2] |= (uint32_t)F.src[o + k] << (8 * (k & 3));
            *reinterpret_cast<uint2*>(data + 8 * v + 8 * (v / (LG_SEG / 16))) = lg_half_steps(make_uint4(t4[0], t4[1], t4[2], t4[3]));
        }
    }
    __syncthreads();
    if (warp != 0) return;

    // ---- chain C0 (tile entry offset 0): lane s owns segment s
    const uint32_t seg_rel = tile_rel > (uint32_t)lane * LG_SEG ? tile_rel - (uint32_t)lane * LG_SEG : 0u;
    const uint8_t* seg = data + lane * LG_NX_PITCH;
    uint32_t bm[LG_SEG_WORDS];
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) bm[wd] = 0;
    constexpr uint32_t NONE = 0xFFu;               // "no chain arrives here" (a predecessor's chain is dead)
    uint32_t entry = 0, p = 0;
    bool dead = lg_walk_segment<false>(seg, p, bm, seg_rel);
    uint32_t exitv = dead ? NONE : p - LG_SEG;     // byte offset into the next segment
    for (;;) {
        uint32_t e = __shfl_up_sync(0xFFFFFFFFu, exitv, 1);
        if (lane == 0) e = 0;
        const bool upd = e != entry;
        if (upd) {
            entry = e;
            if (e == NONE) {
#pragma unroll
                for (int wd = 0; wd < LG_SEG_WORDS; wd++) bm[wd] = 0;
                exitv = NONE;
            } else {
                p = e;
                // drop the old marks before the entry, then walk until the old chain is met
                dead = lg_walk_segment<true>(seg, p, bm, seg_rel);
                if (dead) exitv = NONE;
                else if (p >= (uint32_t)LG_SEG) exitv = p - LG_SEG;      // walked to the end without meeting the old chain
                // else: merged -> the old exit stands
            }
        }
        if (!__any_sync(0xFFFFFFFFu, upd)) break;
    }
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) bitmap[lane * LG_SEG_WORDS + wd] = bm[wd];
    uint32_t cnt = 0;
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) cnt += __popc(bm[wd]);
    uint32_t total0 = cnt;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) total0 += __shfl_xor_sync(0xFFFFFFFFu, total0, d);
    // exit of C0 from the tile: the last segment's; a dead chain anywhere, or the end of the buffer, ends it
    uint32_t exit0 = __shfl_sync(0xFFFFFFFFu, exitv, 31);
    const bool last_tile = tile + 1 == ntile;
    exit0 = (exit0 == NONE || last_tile) ? LG_DEAD : exit0 >> 1;
    __syncwarp();

    // ---- the other entry offsets: walk until C0 is met
    uint32_t* const bmw = bitmap;
    if (lane < LG_STATES) {
        uint32_t q = 2u * lane, pre = 0, m = LG_NO_MERGE, ex = exit0, count;
        bool d2 = false;
        if (lane == 0) { m = 0; count = total0; }
        else {
            for (;;) {
                if (q >= (uint32_t)LG_TILE) { ex = (q - LG_TILE) >> 1; break; }                       // never met C0 in this tile
                if ((bmw[q >> 6] >> ((q >> 1) & 31u)) & 1u) { m = q >> 1; break; }
                const uint32_t nq = q + 2u * data[(q >> 1) + 8u * (q / (uint32_t)LG_SEG)];
                if (nq >= tile_rel) { d2 = true; break; }
                q = nq;
                pre++;
            }
            if (m != LG_NO_MERGE) {
                // blocks of C0 from the merge point on = total0 - (marks before it)
                uint32_t before = 0;
                for (uint32_t w = 0; w < (m >> 5); w++) before += __popc(bmw[w]);
                before += __popc(bmw[m >> 5] & ((1u << (m & 31u)) - 1u));
                count = pre + total0 - before;
            } else {
                count = pre;
                if (d2 || last_tile) ex = LG_DEAD;
            }
        }
        F.lg_tilemap[(size_t)tile * LG_STATES + lane] = ex | (count << 5);
        F.lg_merge[(size_t)tile * LG_STATES + lane] = (uint16_t)m;
    }
    for (int i = lane; i < LG_TILE_WORDS; i += 32) F.lg_bitmap[(size_t)tile * LG_TILE_WORDS + i] = bitmap[i];
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
    const uint32_t ntile = (uint32_t)((len + LG_TILE - 1) / LG_TILE);
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
                    uint32_t m = 0;
                    if (state != LG_DEAD) m = tm[t * LG_STATES + state];
                    tm[t * LG_STATES] = state;               // entries of this tile are no longer needed: reuse two of them
                    tm[t * LG_STATES + 1] = base;
                    if (state != LG_DEAD) { base += m >> 5; state = m & 31u; }
                }
                sh_state = state; sh_base = base;
            }
            __syncthreads();
            for (uint32_t i = tid; i < 2 * nt; i += LG_THREADS) F.lg_tilestate[2 * (size_t)t0 + i] = tm[(i >> 1) * LG_STATES + (i & 1u)];
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
// k_legacy_fix: grid = (ceil(max tiles / LG_THREADS), frames), one THREAD per tile.  The bitmap holds chain C0; for a tile
// entered at offset e != 0 the blocks before the merge point of e are different: walk those few blocks (headers straight
// from global memory) and patch the words in front of the merge point.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LG_THREADS) k_legacy_fix(const FrameDev* __restrict__ frames, const FrameState* __restrict__ states) {
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY || states[blockIdx.y].status[0]) return;
    const uint32_t ntile = (uint32_t)((F.len + LG_TILE - 1) / LG_TILE);
    const uint32_t t = blockIdx.x * LG_THREADS + threadIdx.x;
    if (t >= ntile) return;
    const uint32_t e = F.lg_tilestate[2 * (size_t)t];
    if (e == 0 || e == LG_DEAD) return;
    const uint32_t m = F.lg_merge[(size_t)t * LG_STATES + e];
    if (m == LG_NO_MERGE) { F.lg_tilestate[2 * (size_t)t] = e | LG_SLOW; return; }
    const uint8_t* __restrict__ tsrc = F.src + (unsigned long long)t * LG_TILE;
    uint32_t* bmw = F.lg_bitmap + (size_t)t * LG_TILE_WORDS;
    uint32_t p = 2u * e;                                    // byte offset inside the tile; the merge point is at byte 2 * m
    for (uint32_t w = 0; w <= (m >> 5); w++) {
        uint32_t acc = 0;
        const uint32_t stop = min(64u * (w + 1u), 2u * m);
        while (p < stop) {
            acc |= 1u << ((p >> 1) & 31u);
            p += leg_step(__ldg(tsrc + p));
        }
        if (w == (m >> 5)) acc |= bmw[w] & ~((1u << (m & 31u)) - 1u);          // C0's marks from the merge point on stay
        bmw[w] = acc;
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

// The 2-byte header at byte offset o (even) of the staged tile, as the low 16 bits of the result (byte o first).
__device__ __forceinline__ uint32_t leg_header(const uint8_t* data, uint32_t o) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(data);
    const uint32_t i0 = o >> 2;
    return (o & 2u) ? w[i0] >> 16 : w[i0];
}
__device__ __forceinline__ uint32_t leg_hdr_bits(uint32_t h) { return (h >> 4) & 15u; }                          // RawData_Legacy.cpp:372-375
__device__ __forceinline__ uint32_t leg_hdr_ref(uint32_t h) { return ((h & 15u) << 8) | ((h >> 8) & 0xFFu); }

// The 16 samples of the block at byte offset o whose header says `bits`.
__device__ __forceinline__ void leg_payload(const uint8_t* data, uint32_t o, uint32_t bits, uint32_t (&v)[16]) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(data);
    const uint32_t pi = (o + 2u) >> 2, psh = ((o + 2u) & 2u) * 8u;
    switch (bits) {
    case 0:
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = 0;                                   // :402-404
        break;
    case 1: leg_unpack<1>(w, pi, psh, v); break;
    case 2: leg_unpack<2>(w, pi, psh, v); break;
    case 3: leg_unpack<3>(w, pi, psh, v); break;
    case 4: leg_unpack<4>(w, pi, psh, v); break;
    case 5: leg_unpack<5>(w, pi, psh, v); break;
    case 6: leg_unpack<6>(w, pi, psh, v); break;
    case 7: leg_unpack<7>(w, pi, psh, v); break;
    case 8: leg_unpack<8>(w, pi, psh, v); break;
    case 9: leg_unpack<9>(w, pi, psh, v); break;
    case 10: leg_unpack<10>(w, pi, psh, v); break;
    default: leg_unpack<16>(w, pi, psh, v); break;                               // 11..15 -> 16-bit big-endian (:360-370,395)
    }
}

constexpr int LG_DEC_DATA = LG_TILE + LG_OVERRUN;
constexpr int LG_PAIR_CHUNK = 1024;                   // pairs listed and decoded per pass (a tile holds ~700 for typical images,
                                                      // up to LG_TILE / 4 when every block is 2 bytes: then several passes)
constexpr int LG_DEC_SMEM = LG_DEC_DATA + LG_TILE_WORDS * 4 /*bitmap*/ + LG_PAIR_CHUNK * 2 /*pair list*/ + 64 /*warp sums*/;
static_assert(LG_TILE_WORDS == 2 * LG_THREADS, "k_legacy_decode gives every thread two bitmap words");

__global__ void __launch_bounds__(LG_THREADS, 9) k_legacy_decode(const FrameDev* __restrict__ frames, const FrameState* __restrict__ states) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY || states[blockIdx.y].status[0]) return;
    const unsigned long long len = F.len;
    const uint32_t ntile = (uint32_t)((len + LG_TILE - 1) / LG_TILE);
    const uint32_t tile = blockIdx.x;
    if (tile >= ntile) return;
    const uint32_t tile_state = F.lg_tilestate[2 * (size_t)tile];
    const uint32_t tile_base = F.lg_tilestate[2 * (size_t)tile + 1];
    const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;                            // pairs per row (RawData_Legacy.cpp:34-36)
    const unsigned long long need = 2ull * ppr * (unsigned long long)F.height;       // blocks of the image (:478-482), < 2^33
    if ((tile_state & 31u) == LG_DEAD || (unsigned long long)tile_base >= need) return;      // nothing of the image starts here

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* data = lg_smem;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LG_DEC_DATA);                        // [LG_TILE_WORDS]
    uint16_t* plist = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(bitmap) + LG_TILE_WORDS * 4);
    uint32_t* warp_sums = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(plist) + LG_PAIR_CHUNK * 2);

    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    lg_stage<LG_THREADS>(data, F.src, len, tile_off, LG_DEC_DATA, tid);
    if (tile_state & LG_SLOW) {
        // the chain entering this tile never meets C0 inside it (blocks of one constant width): walk it here
        for (int i = tid; i < LG_TILE_WORDS; i += LG_THREADS) bitmap[i] = 0;
        __syncthreads();
        if (tid == 0) {
            const uint32_t tile_rel = (uint32_t)min(len - tile_off, (unsigned long long)(1u << 30));
            uint32_t p = 2u * (tile_state & 31u);
            while (p < (uint32_t)LG_TILE) {
                const uint32_t q = p + leg_step(data[p]);
                if (q >= tile_rel) break;
                bitmap[p >> 6] |= 1u << ((p >> 1) & 31u);
                p = q;
            }
        }
    } else {
        for (int i = tid; i < LG_TILE_WORDS; i += LG_THREADS) bitmap[i] = F.lg_bitmap[(size_t)tile * LG_TILE_WORDS + i];
    }
    __syncthreads();
    // ---- pair list of the tile: every block with an even ordinal leads a pair (even-column block, then odd-column block,
    //      RawData_Legacy.cpp:480-481); plist[q] = (tile-relative offset of the leader) / 2 for pair ordinal p_first + q.
    //      Ordinals come from prefix popcounts of the bitmap: thread t owns words 2t and 2t+1.
    const uint32_t w0 = bitmap[2 * tid], w1 = bitmap[2 * tid + 1];
    const uint32_t c = __popc(w0) + __popc(w1);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < LG_THREADS / 32; w++) {
        const uint32_t v = warp_sums[w];
        if (w < warp) before += v;
        total += v;
    }
    const uint32_t p_first = (tile_base + 1u) >> 1;
    uint32_t npairs = ((tile_base + total + 1u) >> 1) - p_first;
    npairs = (uint32_t)min((unsigned long long)npairs, (need >> 1) - (unsigned long long)p_first);
    // ---- decode: a lane takes one block pair at a time -> 32 consecutive pixels
    const int width = F.width;
    uint16_t* __restrict__ dst = F.dst;
    const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
    const uint32_t ord0 = tile_base + before + incl - c;     // ordinal of the first block start in this thread's words
    for (uint32_t c0 = 0; c0 < npairs; c0 += LG_PAIR_CHUNK) {
        const uint32_t cn = min((uint32_t)LG_PAIR_CHUNK, npairs - c0);
        {
            uint32_t ord = ord0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t wv = h ? w1 : w0;
                while (wv) {
                    const uint32_t b = __ffs(wv) - 1;
                    wv &= wv - 1;
                    const uint32_t q = (ord >> 1) - p_first - c0;       // wraps to a huge value for earlier passes' pairs
                    if (!(ord & 1u) && q < cn) plist[q] = (uint16_t)(32u * (2u * tid + h) + b);
                    ord++;
                }
            }
        }
        __syncthreads();
        uint32_t P = p_first + c0 + (uint32_t)tid;
        uint32_t y = P / ppr, xq = P - y * ppr;
        // the leader's offset and header are fetched one pair ahead; both headers of a pair before either payload
        uint32_t o = 0, hE = 0;
        if ((uint32_t)tid < cn) { o = 2u * (uint32_t)plist[tid]; hE = leg_header(data, o); }
        for (uint32_t q = tid; q < cn; q += LG_THREADS) {
            const uint32_t oE = o, bitsE = leg_hdr_bits(hE);
            const uint32_t oO = oE + 2u + leg_len(bitsE);
            const uint32_t hO = leg_header(data, oO);
            const uint32_t refs = leg_hdr_ref(hE) | (leg_hdr_ref(hO) << 16);
            if (q + LG_THREADS < cn) { o = 2u * (uint32_t)plist[q + LG_THREADS]; hE = leg_header(data, o); }
            uint32_t vE[16], vO[16];
            leg_payload(data, oE, bitsE, vE);
            leg_payload(data, oO, leg_hdr_bits(hO), vO);
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
        __syncthreads();
    }
}

// =========================================================================================================
// k_legacy_fused: the whole legacy decode in ONE pass over the stream (replaces maps + scan + fix + decode).
//
// Persistent CTAs take (frame, tile) tickets in a host-built order (tile index major, frame minor: neighbouring
// tickets belong to different frames, so every frame's chain only has to advance a few tiles per generation of CTAs).
// A CTA is a two-stage pipeline over two shared-memory buffers, handed back and forth with mbarriers:
//
//   INDEX WARP (warp 0), one tile ahead of the others:
//   1. stage the tile (+ overrun) with ONE bulk copy (cp.async.bulk, TMA 1-D; mbarrier transaction count) -- the only time
//      the stream is read;
//   2. resolve chain C0 (entry offset 0) with the self-synchronising segment walk, then the other 16 entry offsets up to
//      their merge point with C0: the tile's transfer map  entry -> (exit offset, block count);
//   3. publish the map (LOCAL), then DECOUPLED LOOK-BACK over the previous tiles of the frame: the nearest predecessor
//      whose inclusive state (exit offset, blocks so far) is known, composed with the maps of the tiles in between
//      (a window of 32 status words per poll; one lane chases the concrete entry state through the staged maps);
//      publish this tile's inclusive state (INCL) right away, so that successors can go on;
//   4. patch the bitmap for the true entry (the few blocks before the merge point; a chain that never meets C0 is
//      re-walked with the segment walk from its entry) and hand the buffer to the decode warps.
//   DECODE WARPS (warps 1..4): prefix popcounts over the bitmap give every thread the ordinal of the first block start
//      in its own 128 bytes of the stream; it decodes the block PAIRS led from there (even-column block + odd-column
//      block, RawData_Legacy.cpp:480-481) straight from shared memory and gives the buffer back.
//
// Ticket order guarantees that every predecessor a tile waits for has been started by a resident CTA (which never waits
// for a successor), so the waits always end; they are bounded all the same (MCRAW_FRAME_INTERNAL instead of a hang).
// Status words carry the launch epoch of the slot, so nothing has to be zeroed between launches.
// =========================================================================================================
struct LgWork { uint32_t frame, tile; };

constexpr int LGF_DEC_WARPS = 4;
constexpr int LGF_DEC_THREADS = 32 * LGF_DEC_WARPS;     // thread t of the decode warps owns bitmap words 2t and 2t+1
constexpr int LGF_THREADS = 32 + LGF_DEC_THREADS;       // warp 0: index warp
constexpr int LGF_DATA = LG_TILE + LG_OVERRUN;
constexpr int LGF_BUF = LGF_DATA + LG_TILE_WORDS * 4;   // one pipeline stage: tile bytes, then the bitmap of block starts
constexpr int LGF_LB = 32;                              // look-back window: status words read per poll
constexpr int LGF_SMEM = 2 * LGF_BUF + LGF_LB * LG_STATES * 4;
constexpr uint32_t LGF_ST_LOCAL = 1u, LGF_ST_INCL = 2u;
constexpr uint32_t LGF_ERR_BIT = 1u << 5;               // sticky: a wait gave up somewhere up the chain
constexpr uint32_t LGF_SPIN_LIMIT = 1u << 22;
constexpr uint32_t LGF_STAGE_DONE = 1u, LGF_STAGE_SKIP = 2u;
static_assert(LGF_DATA % 16 == 0 && LGF_BUF % 16 == 0, "bulk copies work in 16-byte granules");
static_assert(LG_TILE_WORDS == 2 * LGF_DEC_THREADS, "every decode thread owns two bitmap words");

struct LgfStage { uint32_t frame, tile, base, flags; };

// status word of a tile: blocks up to the end of the tile << 32 | epoch (24 bits) << 8 | state << 6 | error << 5 | exit offset / 2
__device__ __forceinline__ unsigned long long lgf_pack(uint32_t count, uint32_t epoch, uint32_t st, uint32_t low6) {
    return ((unsigned long long)count << 32) | ((unsigned long long)(epoch & 0xFFFFFFu) << 8) | (st << 6) | low6;
}
__device__ __forceinline__ void lgf_store_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long lgf_load_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}

// One lane, one segment of the staged tile: walk from tile-relative byte offset p to the end of segment `seg`.  Block
// starts go into bm[] (one word per 64 bytes of the segment).  With MERGE the walk stops at the first position already
// marked in bm[] -- from there on the old marks are this chain's own -- and older marks before it are dropped.
// rel: bytes from the tile start to the end of the buffer (a block is decoded only if it ends before the last byte,
// RawData_Legacy.cpp:387,398).  Returns true if the chain ended at an undecodable block.
template <bool MERGE>
__device__ __forceinline__ bool lgf_walk_segment(const uint8_t* data, const uint32_t seg, uint32_t& p, uint32_t (&bm)[LG_SEG_WORDS], const uint32_t rel) {
    bool merged = false, dead = false;
    const uint32_t seg0 = seg * LG_SEG;
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) {
        if (merged) continue;
        const uint32_t stop = seg0 + 64u * (wd + 1);
        if (dead || p >= stop) { bm[wd] = 0; continue; }
        const uint32_t old = bm[wd];
        uint32_t acc = 0;
        while (p < stop) {
            const uint32_t bit = 1u << ((p >> 1) & 31u);
            if (MERGE && (old & bit)) { merged = true; acc |= old & ~(bit - 1u); break; }
            const uint32_t q = p + leg_step(data[p]);
            if (q >= rel) { dead = true; break; }
            acc |= bit;
            p = q;
        }
        bm[wd] = acc;
    }
    return dead;
}

// Warp-wide: the exact chain that enters the tile at byte offset entry0 (even, < LG_SEG), as a bitmap of block starts in
// shared memory.  Lane s owns segment s: a guessed chain from the segment start first, then re-walks from the left
// neighbour's real exit until the lane's own marks are met, repeated until no entry changes.
// Returns (all lanes) the chain's exit from the tile: byte offset into the next tile, or 0xFFFFFFFF if the chain died.
__device__ __forceinline__ uint32_t lgf_chain(const uint8_t* data, uint32_t* bitmap, const uint32_t entry0, const uint32_t tile_rel,
                                              const uint32_t lane, uint32_t& total) {
    constexpr uint32_t NONE = 0xFFFFFFFFu;
    uint32_t bm[LG_SEG_WORDS];
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) bm[wd] = 0;
    const uint32_t seg0 = lane * LG_SEG;
    uint32_t entry = lane == 0 ? entry0 : seg0;          // tile-relative position where this lane's walk starts
    uint32_t p = entry;
    bool dead = lgf_walk_segment<false>(data, lane, p, bm, tile_rel);
    uint32_t exitv = dead ? NONE : p;                    // tile-relative position in the next segment (steps are <= 34 bytes)
    for (;;) {
        uint32_t e = __shfl_up_sync(0xFFFFFFFFu, exitv, 1);
        if (lane == 0) e = entry0;
        const bool upd = e != entry;
        if (upd) {
            entry = e;
            if (e == NONE) {
#pragma unroll
                for (int wd = 0; wd < LG_SEG_WORDS; wd++) bm[wd] = 0;
                exitv = NONE;
            } else {
                p = e;
                dead = lgf_walk_segment<true>(data, lane, p, bm, tile_rel);
                if (dead) exitv = NONE;
                else if (p >= seg0 + (uint32_t)LG_SEG) exitv = p;      // walked to the end without meeting the old chain
                // else: merged -> the old exit stands
            }
        }
        if (!__any_sync(0xFFFFFFFFu, upd)) break;
    }
    uint32_t cnt = 0;
#pragma unroll
    for (int wd = 0; wd < LG_SEG_WORDS; wd++) { bitmap[lane * LG_SEG_WORDS + wd] = bm[wd]; cnt += __popc(bm[wd]); }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    total = cnt;
    const uint32_t ex = __shfl_sync(0xFFFFFFFFu, exitv, 31);
    __syncwarp();
    return ex == NONE ? NONE : ex - (uint32_t)LG_TILE;
}

// OR the 16 samples of the block at byte offset o (header nibble `bits`) into px[], at bit ADJ of each word (0: even-column
// block, 16: odd-column block, RawData_Legacy.cpp:483-486).  d32: the staged tile as words.  The payload is a contiguous
// MSB-first bit stream (:38-358), so 8 samples are exactly `bits` bytes: per group of 8 the bytes are fetched as
// big-endian words (PRMT with a runtime selector does alignment and byte order in one go) and every sample is one
// rotate + one mask.  Three lane-uniform formulations instead of one code path per width: widths 0..8 (two 4-sample
// windows per group), 9..10 (four 2-sample windows), and 16-bit big-endian samples (:360-370, nibbles 11..15, :395).
template <int ADJ>
__device__ __forceinline__ void lgf_block(const uint32_t* __restrict__ d32, const uint32_t o, const uint32_t bits, uint32_t (&px)[16]) {
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

__global__ void __launch_bounds__(LGF_THREADS) k_legacy_fused(const FrameDev* __restrict__ frames, Result* __restrict__ results,
                                                              const LgWork* __restrict__ work, const uint32_t nwork,
                                                              uint32_t* __restrict__ counters, const uint32_t epoch) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    uint32_t* lbmaps = reinterpret_cast<uint32_t*>(lg_smem + 2 * LGF_BUF);                          // look-back: [LGF_LB][LG_STATES]
    __shared__ __align__(8) unsigned long long bars[6];              // per buffer: loaded (bulk copy), full (index -> decode), empty (decode -> index)
    __shared__ LgfStage stage[2];
    __shared__ uint32_t sh_map[LG_STATES], sh_merge[LG_STATES], warp_sums[LGF_DEC_WARPS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) mbar_init(bar0 + 8u * i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ======================================== index warp ========================================
        for (uint32_t k = 0;; k++) {
            const uint32_t b = k & 1u, par = (k >> 1) & 1u;
            uint8_t* data = lg_smem + b * LGF_BUF;
            uint32_t* bitmap = reinterpret_cast<uint32_t*>(data + LGF_DATA);
            const uint32_t bar_ld = bar0 + 8u * b, bar_full = bar0 + 16u + 8u * b, bar_empty = bar0 + 32u + 8u * b;
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(&counters[2], 1u);
            ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
            mbar_wait(bar_empty, par ^ 1u);                   // the decode warps are done with this buffer (passes at once the first time)
            if (ticket >= nwork) {
                if (lane == 0) { stage[b].flags = LGF_STAGE_DONE; mbar_arrive(bar_full); }
                break;
            }
            const LgWork wk = work[ticket];
            const FrameDev& F = frames[wk.frame];
            const unsigned long long len = F.len;
            const uint32_t ntile = (uint32_t)max((len + LG_TILE - 1) / LG_TILE, 1ull);
            const uint32_t tile = wk.tile;
            const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
            const uint32_t tile_rel = (uint32_t)min(len > tile_off ? len - tile_off : 0ull, (unsigned long long)(1u << 30));
            const bool last_tile = tile + 1 == ntile;
            const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;                            // pairs per row (RawData_Legacy.cpp:34-36)
            const unsigned long long need = 2ull * ppr * (unsigned long long)F.height;       // blocks of the image (:478-482)
            const bool fits = F.dst_cap >= (unsigned long long)F.width * (unsigned long long)F.height;

            // ---- 1. stage: one bulk copy for a tile that lies wholly inside the buffer, else 16-byte granules with zero fill
            if (tile_off + (unsigned long long)LGF_DATA <= len) {
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_ld), "r"((uint32_t)LGF_DATA) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                                 ::"r"(smem_u32(data)), "l"(F.src + tile_off), "r"((uint32_t)LGF_DATA), "r"(bar_ld) : "memory");
                }
            } else {
                lg_stage<32>(data, F.src, len, tile_off, LGF_DATA, lane);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // a later bulk copy overwrites these generic stores
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ld);
            }
            mbar_wait(bar_ld, par);

            // ---- 2. transfer map
            uint32_t total0;
            const uint32_t ex0 = lgf_chain(data, bitmap, 0u, tile_rel, lane, total0);
            const uint32_t exit0 = (ex0 == 0xFFFFFFFFu || last_tile) ? LG_DEAD : ex0 >> 1;
            if (lane < LG_STATES) {
                uint32_t q = 2u * lane, pre = 0, m = LG_NO_MERGE, ex = exit0, count;
                bool d2 = false;
                if (lane == 0) { m = 0; count = total0; }
                else {
                    for (;;) {
                        if (q >= (uint32_t)LG_TILE) { ex = (q - LG_TILE) >> 1; break; }                 // never met C0 in this tile
                        if ((bitmap[q >> 6] >> ((q >> 1) & 31u)) & 1u) { m = q >> 1; break; }
                        const uint32_t nq = q + leg_step(data[q]);
                        if (nq >= tile_rel) { d2 = true; break; }
                        q = nq;
                        pre++;
                    }
                    if (m != LG_NO_MERGE) {
                        uint32_t before = 0;                    // blocks of C0 from the merge point on = total0 - (marks before it)
                        for (uint32_t w = 0; w < (m >> 5); w++) before += __popc(bitmap[w]);
                        before += __popc(bitmap[m >> 5] & ((1u << (m & 31u)) - 1u));
                        count = pre + total0 - before;
                    } else {
                        count = pre;
                        if (d2 || last_tile) ex = LG_DEAD;
                    }
                }
                const uint32_t mapv = ex | (count << 5);
                sh_map[lane] = mapv;
                sh_merge[lane] = m;
                F.lg_tilemap[(size_t)tile * LG_STATES + lane] = mapv;
                __threadfence();
            }
            __syncwarp();
            // ---- 3. publish, look back, publish again
            if (lane == 0) {
                __threadfence();                                 // release: the map before the status word
                lgf_store_status(F.lg_status + tile, lgf_pack(0u, epoch, LGF_ST_LOCAL, 0u));
            }
            uint32_t entry = 0, base = 0, errbit = 0;
            if (tile > 0) {
                const uint32_t jhi = tile - 1;
                uint32_t spins = 0;
                for (;;) {
                    const int j = (int)jhi - lane;
                    unsigned long long sw = 0;
                    if (j >= 0) sw = lgf_load_status(F.lg_status + j);
                    const uint32_t lo32 = (uint32_t)sw;
                    const uint32_t st = ((lo32 >> 8) & 0xFFFFFFu) == (epoch & 0xFFFFFFu) ? (lo32 >> 6) & 3u : 0u;
                    const unsigned incl = __ballot_sync(0xFFFFFFFFu, st == LGF_ST_INCL);
                    const unsigned any = __ballot_sync(0xFFFFFFFFu, st != 0u);
                    if (incl) {
                        const int d = __ffs(incl) - 1;            // nearest predecessor with a known inclusive state: tile jhi - d
                        const unsigned between = (1u << d) - 1u;  // tiles jhi - d + 1 .. jhi must have published their maps
                        if ((any & between) == between) {
                            uint32_t state = __shfl_sync(0xFFFFFFFFu, lo32 & 31u, d);
                            uint32_t count = __shfl_sync(0xFFFFFFFFu, (uint32_t)(sw >> 32), d);
                            errbit = __shfl_sync(0xFFFFFFFFu, lo32 & LGF_ERR_BIT, d);
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
                    if (++spins > LGF_SPIN_LIMIT) { entry = LG_DEAD; errbit = LGF_ERR_BIT; break; }   // never expected
                    __nanosleep(spins < 16 ? 32 : 200);
                }
            }
            uint32_t exitv = LG_DEAD, total = base;
            if (entry != LG_DEAD) {
                const uint32_t v = sh_map[entry];
                exitv = v & 31u;
                total = base + (v >> 5);
            }
            if (lane == 0) {
                lgf_store_status(F.lg_status + tile, lgf_pack(total, epoch, LGF_ST_INCL, exitv | errbit));
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
            }
            // ---- 4. the bitmap for the true entry, then over to the decode warps
            const bool skip = entry == LG_DEAD || !fits || (unsigned long long)base >= need;     // nothing of the image starts here
            if (!skip && entry != 0) {
                const uint32_t m = sh_merge[entry];
                __syncwarp();
                if (m == LG_NO_MERGE) {
                    uint32_t t2;
                    lgf_chain(data, bitmap, 2u * entry, tile_rel, lane, t2);       // blocks of one constant width: walk it again from its entry
                } else {
                    for (uint32_t w = lane; w < (m >> 5); w += 32) bitmap[w] = 0;  // C0's marks before the merge point go
                    __syncwarp();
                    if (lane == 0) {
                        bitmap[m >> 5] &= ~((1u << (m & 31u)) - 1u);
                        uint32_t p = 2u * entry;
                        while (p < 2u * m) {
                            bitmap[p >> 6] |= 1u << ((p >> 1) & 31u);
                            p += leg_step(data[p]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                LgfStage s;
                s.frame = wk.frame; s.tile = tile; s.base = base; s.flags = skip ? LGF_STAGE_SKIP : 0u;
                stage[b] = s;
                mbar_arrive(bar_full);                        // release: bitmap, stage record (and the bulk copy observed above)
            }
        }
        // the last CTA to leave resets the ticket counters for the next launch
        if (lane == 0 && atomicAdd(&counters[3], 1u) == gridDim.x - 1u) { counters[2] = 0; counters[3] = 0; }
        return;
    }

    // ======================================== decode warps ========================================
    const uint32_t dt = (uint32_t)tid - 32u, dwarp = (uint32_t)warp - 1u;
    for (uint32_t k = 0;; k++) {
        const uint32_t b = k & 1u, par = (k >> 1) & 1u;
        const uint8_t* data = lg_smem + b * LGF_BUF;
        const uint32_t* bitmap = reinterpret_cast<const uint32_t*>(data + LGF_DATA);
        const uint32_t bar_full = bar0 + 16u + 8u * b, bar_empty = bar0 + 32u + 8u * b;
        mbar_wait(bar_full, par);
        const LgfStage st = stage[b];
        if (st.flags & LGF_STAGE_DONE) break;
        if (!(st.flags & LGF_STAGE_SKIP)) {
            const FrameDev& F = frames[st.frame];
            const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;
            const uint32_t need_pairs = ppr * (uint32_t)F.height;               // < 2^26: width * height <= 2^30 (prepare())
            // ordinal of the first block start in this thread's 128 bytes: prefix popcounts over the bitmap
            const uint32_t w0 = bitmap[2 * dt], w1 = bitmap[2 * dt + 1];
            const uint32_t c = __popc(w0) + __popc(w1);
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) warp_sums[dwarp] = incl;
            asm volatile("bar.sync 1, %0;\n" ::"n"(LGF_DEC_THREADS) : "memory");
            uint32_t before = 0;
#pragma unroll
            for (int w = 0; w < LGF_DEC_WARPS; w++)
                if ((uint32_t)w < dwarp) before += warp_sums[w];
            const uint32_t ord0 = st.base + before + incl - c;
            // block starts with an even ordinal lead a pair (even-column block, then odd-column block, :480-481); the
            // partner's start is the next mark -- in these words or the next thread's, where it has an odd ordinal and is dropped
            unsigned long long marks = (unsigned long long)w0 | ((unsigned long long)w1 << 32);
            if ((ord0 & 1u) && marks) marks &= marks - 1;
            uint32_t P = (ord0 >> 1) + (ord0 & 1u);                                  // ordinal of the first pair led here
            uint32_t y = P / ppr, xq = P - y * ppr;
            const int width = F.width;
            uint16_t* __restrict__ dst = F.dst;
            const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
            const uint32_t* d32 = reinterpret_cast<const uint32_t*>(data);
            while (marks && P < need_pairs) {
                const uint32_t bpos = (uint32_t)__ffsll((long long)marks) - 1u;
                marks &= marks - 1;                                                   // the leader ...
                marks &= marks - 1;                                                   // ... and its partner, if it starts in these words
                const uint32_t oE = 128u * dt + 2u * bpos;                               // two bitmap words = 128 bytes of the tile
                const uint32_t hE = leg_header(data, oE), bitsE = leg_hdr_bits(hE);
                const uint32_t oO = oE + 2u + leg_len(bitsE);
                const uint32_t hO = leg_header(data, oO), bitsO = leg_hdr_bits(hO);
                uint32_t px[16];
#pragma unroll
                for (int i = 0; i < 16; i++) px[i] = 0;
                lgf_block<0>(d32, oE, bitsE, px);
                lgf_block<16>(d32, oO, bitsO, px);
                const uint32_t refs = leg_hdr_ref(hE) | (leg_hdr_ref(hO) << 16);
#pragma unroll
                for (int i = 0; i < 16; i++) px[i] = __vadd2(px[i], refs);                    // :483-486, + reference mod 2^16
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
                P++;
                if (++xq == ppr) { xq = 0; y++; }
            }
        }
        // every decode thread is done with the buffer (and with warp_sums): give it back
        asm volatile("bar.sync 1, %0;\n" ::"n"(LGF_DEC_THREADS) : "memory");
        if (dt == 0) mbar_arrive(bar_empty);
    }
}

}  // namespace mcraw
