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

// ---------------------------------------------------------------------------------------------------------
// k_legacy_maps: grid = (max tiles, frames), block = LG_THREADS, dynamic smem = LG_TILE + 2 * 32 * 18
// ---------------------------------------------------------------------------------------------------------
constexpr int LG_MAPS_SMEM = LG_TILE + LG_TILE_SEGS * 18 * 2;

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
    uint16_t* maps = reinterpret_cast<uint16_t*>(lg_smem + LG_TILE);          // [32][18]
    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    lg_stage(data, F.src, len, tile_off, LG_TILE, tid);
    __syncthreads();

    const uint32_t segs_here = min((uint32_t)LG_TILE_SEGS, nseg - tile * LG_TILE_SEGS);
    for (uint32_t s = warp; s < segs_here; s += LG_THREADS / 32) {
        if (lane < LG_STATES) {
            const unsigned long long seg_abs = tile_off + (unsigned long long)s * LG_SEG;
            const uint8_t* seg = data + s * LG_SEG;
            uint32_t p = 2u * lane, cnt = 0;
            bool dead = false;
            while (p < (uint32_t)LG_SEG) {
                // RawData_Legacy.cpp:387,398: a block is decoded only if offset + 2 + payload < len
                const uint32_t L = leg_len(seg[p] >> 4);
                if (seg_abs + p + 2u + L >= len) { dead = true; break; }
                p += 2u + L;
                cnt++;
            }
            const uint32_t m = (dead ? LG_DEAD : (p - LG_SEG) >> 1) | (cnt << 5);
            maps[s * 18 + lane] = (uint16_t)m;
            F.lg_segmap[((size_t)tile * LG_TILE_SEGS + s) * LG_STATES + lane] = (uint16_t)m;
        }
    }
    __syncthreads();
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
// 16 samples of W bits each, MSB-first contiguous (RawData_Legacy.cpp:38-358), from big-endian words be[].
template <int W>
__device__ __forceinline__ void leg_unpack(const uint32_t (&be)[9], uint32_t (&v)[16]) {
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
    // bytes o .. o+35 as little-endian words aligned to the block start
    uint32_t a = w[i0], b = w[i0 + 1];
    const uint32_t h = __funnelshift_r(a, b, sh);
    const uint32_t bits = (h >> 4) & 15u;                                        // RawData_Legacy.cpp:372-375
    ref = ((h & 15u) << 8) | ((h >> 8) & 0xFFu);
    const uint32_t L = leg_len(bits);
    // payload words (big-endian view) starting at byte o + 2
    uint32_t be[9];
    const uint32_t pi = (o + 2u) >> 2, psh = ((o + 2u) & 2u) * 8u;
    const int nw = (int)((L + 3u) >> 2);
#pragma unroll
    for (int k = 0; k < 9; k++) be[k] = 0;
    {
        uint32_t lo = w[pi];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < nw) {
                const uint32_t hi = w[pi + k + 1];
                be[k] = __byte_perm(__funnelshift_r(lo, hi, psh), 0u, 0x0123);
                lo = hi;
            }
        }
    }
    switch (bits) {
    case 0:
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = 0;                                   // :402-404
        break;
    case 1: leg_unpack<1>(be, v); break;
    case 2: leg_unpack<2>(be, v); break;
    case 3: leg_unpack<3>(be, v); break;
    case 4: leg_unpack<4>(be, v); break;
    case 5: leg_unpack<5>(be, v); break;
    case 6: leg_unpack<6>(be, v); break;
    case 7: leg_unpack<7>(be, v); break;
    case 8: leg_unpack<8>(be, v); break;
    case 9: leg_unpack<9>(be, v); break;
    case 10: leg_unpack<10>(be, v); break;
    default:                                                                     // 11..15 -> 16-bit big-endian (:360-370,395)
#pragma unroll
        for (int k = 0; k < 8; k++) { v[2 * k] = be[k] >> 16; v[2 * k + 1] = be[k] & 0xFFFFu; }
        break;
    }
    return 2u + L;
}

constexpr int LG_DEC_DATA = LG_TILE + LG_OVERRUN;
constexpr int LG_DEC_SMEM = LG_DEC_DATA + LG_TILE_SEGS * 18 * 2 /*maps*/ + LG_TILE_SEGS * (LG_SLOTS / 8) /*bitmaps*/ +
                            (LG_THREADS / 32) * LG_SLOTS * 2 /*lists*/ + LG_TILE_SEGS * 8 /*entry, base*/;

__global__ void __launch_bounds__(LG_THREADS) k_legacy_decode(const FrameDev* __restrict__ frames, const FrameState* __restrict__ states) {
    extern __shared__ __align__(16) uint8_t lg_smem[];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_LEGACY || states[blockIdx.y].status[0]) return;
    const unsigned long long len = F.len;
    const uint32_t nseg = (uint32_t)((len + LG_SEG - 1) / LG_SEG);
    const uint32_t tile = blockIdx.x;
    if ((unsigned long long)tile * LG_TILE_SEGS >= nseg) return;
    const uint32_t tile_entry = F.lg_tilestate[2 * (size_t)tile];
    const uint32_t tile_base = F.lg_tilestate[2 * (size_t)tile + 1];
    const uint32_t ppr = ((uint32_t)F.width + 31u) / 32u;
    const unsigned long long need = 2ull * ppr * (unsigned long long)F.height;
    if (tile_entry == LG_DEAD || (unsigned long long)tile_base >= need) return;      // nothing of the image starts here

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* data = lg_smem;
    uint16_t* maps = reinterpret_cast<uint16_t*>(lg_smem + LG_DEC_DATA);                          // [32][18]
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(lg_smem + LG_DEC_DATA + LG_TILE_SEGS * 36);    // [32][16]
    uint16_t* lists = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(bitmap) + LG_TILE_SEGS * (LG_SLOTS / 8));
    uint32_t* seg_entry = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(lists) + (LG_THREADS / 32) * LG_SLOTS * 2);
    uint32_t* seg_base = seg_entry + LG_TILE_SEGS;

    const unsigned long long tile_off = (unsigned long long)tile * LG_TILE;
    const uint32_t segs_here = min((uint32_t)LG_TILE_SEGS, nseg - tile * LG_TILE_SEGS);
    lg_stage(data, F.src, len, tile_off, LG_DEC_DATA, tid);
    for (uint32_t i = tid; i < segs_here * LG_STATES; i += LG_THREADS) {
        const uint32_t s = i / LG_STATES, e = i - s * LG_STATES;
        maps[s * 18 + e] = F.lg_segmap[((size_t)tile * LG_TILE_SEGS + s) * LG_STATES + e];
    }
    for (uint32_t i = tid; i < LG_TILE_SEGS * (LG_SLOTS / 32); i += LG_THREADS) bitmap[i] = 0;
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
    }
    __syncthreads();
    // ---- one lane per segment: walk the (now known) chain and mark the block starts
    if (warp == 0) {
        const uint32_t s = lane;
        const uint32_t e = seg_entry[s];
        if (e != LG_DEAD) {
            const unsigned long long seg_abs = tile_off + (unsigned long long)s * LG_SEG;
            const uint8_t* seg = data + s * LG_SEG;
            uint32_t* bm = bitmap + s * (LG_SLOTS / 32);
            uint32_t p = 2u * e;
            uint32_t cur_word = p >> 6, acc = 0;
            while (p < (uint32_t)LG_SEG) {
                const uint32_t L = leg_len(seg[p] >> 4);
                if (seg_abs + p + 2u + L >= len) break;
                const uint32_t slot = p >> 1, wd = slot >> 5;
                if (wd != cur_word) { bm[cur_word] = acc; acc = 0; cur_word = wd; }
                acc |= 1u << (slot & 31u);
                p += 2u + L;
            }
            if (cur_word < (uint32_t)(LG_SLOTS / 32)) bm[cur_word] = acc;
        }
    }
    __syncthreads();
    // ---- decode: every warp takes segments warp, warp + 8, ...; a lane decodes one block pair at a time
    const int width = F.width;
    uint16_t* __restrict__ dst = F.dst;
    const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
    uint16_t* list = lists + warp * LG_SLOTS;
    for (uint32_t s = warp; s < segs_here; s += LG_THREADS / 32) {
        if (seg_entry[s] == LG_DEAD) break;
        const uint32_t wordv = lane < (uint32_t)(LG_SLOTS / 32) ? bitmap[s * (LG_SLOTS / 32) + lane] : 0u;
        const uint32_t c = __popc(wordv);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += o;
        }
        const uint32_t nblk = __shfl_sync(0xFFFFFFFFu, incl, 31);
        {
            uint32_t wv = wordv, k = incl - c;
            while (wv) {
                const uint32_t b = __ffs(wv) - 1;
                wv &= wv - 1;
                list[k++] = (uint16_t)(32u * lane + b);
            }
        }
        __syncwarp();
        const uint32_t base = seg_base[s];
        const uint32_t j0 = base & 1u;                       // an odd first block belongs to the pair led from the previous segment
        for (uint32_t j = j0 + 2u * lane; j < nblk; j += 64u) {
            const unsigned long long P = ((unsigned long long)base + j) >> 1;      // pair ordinal in the frame
            if (2ull * P >= need) break;
            const uint32_t o = s * LG_SEG + 2u * (uint32_t)list[j];
            uint32_t vE[16], vO[16], refE, refO;
            const uint32_t lenE = leg_block(data, o, vE, refE);
            leg_block(data, o + lenE, vO, refO);
            const uint32_t y = (uint32_t)(P / ppr), xq = (uint32_t)(P - (unsigned long long)y * ppr);
            const int x = (int)(32u * xq);
            uint32_t px[16];
#pragma unroll
            for (int k = 0; k < 16; k++)                                            // :483-486, u16 wrap
                px[k] = ((vE[k] + refE) & 0xFFFFu) | ((vO[k] + refO) << 16);
            uint16_t* orow = dst + (size_t)y * (size_t)width + x;
            if (vec && x + 32 <= width) {
                uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
                for (int k = 0; k < 4; k++) o4[k] = make_uint4(px[4 * k], px[4 * k + 1], px[4 * k + 2], px[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {                                      // crop at width (:490)
                    if (x + 2 * k < width) orow[2 * k] = (uint16_t)px[k];
                    if (x + 2 * k + 1 < width) orow[2 * k + 1] = (uint16_t)(px[k] >> 16);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace mcraw
