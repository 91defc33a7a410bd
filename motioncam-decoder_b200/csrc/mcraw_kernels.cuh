// mcraw_kernels.cuh -- sm_100a kernels of the MCRAW frame decoder.
//
// Current format (compressionType 7; reference: /root/reference/lib/RawData.cpp:528-612)
//   k_meta   one CTA per (frame, metadata stream): walks the inline-header chain of the stream
//            (RawData.cpp:463-498), unpacks the 64-value meta blocks, and -- for the "bits" stream -- turns
//            the running `offset +=` of the reference tile loop (RawData.cpp:562,576-579) into an exclusive
//            prefix sum.  Output: one 16-byte record per 64x4-pixel tile
//                { payload offset of the tile, bits[4], refs[0..1], refs[2..3] }.
//   k_tiles  one CTA per (frame, tile row): every lane decodes one 8-sample plane of an even/odd block pair
//            with 32-bit SWAR (table in mcraw_tables.h), interleaves the two Bayer phases with PRMT, adds the
//            per-block references with packed 16-bit adds (wraps mod 2^16 like the reference's uint16 stores,
//            RawData.cpp:582-592) and writes 2 x 16 bytes of one output row; columns >= width are cropped
//            (RawData.cpp:598-608).
//
// Legacy format (compressionType 6; reference: /root/reference/lib/RawData_Legacy.cpp:445-495)
//   k_legacy_index / k_legacy_decode  -- see below.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mcraw_b200.h"
#include "mcraw_tables.h"

namespace mcraw {

struct FrameDev {
    const uint8_t* src;            // device, 16-byte aligned
    unsigned long long len;
    uint16_t* dst;                 // device
    unsigned long long dst_cap;    // elements
    int width, height, type;
    unsigned tiles_x;              // expected encodedWidth/64   = ceil(width/64)
    unsigned tile_rows;            // expected ceil(encodedHeight/4) upper bound = ceil(height/4)
    unsigned flags;                // bit0: 16-byte vector stores allowed (width % 8 == 0, dst 16-byte aligned)
    uint4* tilemeta;               // scratch, tiles_x*tile_rows records               (type 7)
    uint32_t* aux;                 // scratch for the legacy index                    (type 6)
    unsigned long long aux_elems;
    // written on the device
    unsigned status;               // MCRAW_FRAME_* bits
    unsigned tile_rows_dev;        // ceil(encodedHeight/4) from the frame header
};

struct Result {
    unsigned long long written;    // uint16 elements, 0 = failed
    unsigned status;
    unsigned pad;
};

enum { FLAG_VEC_STORE = 1 };

__device__ __forceinline__ uint32_t ld_u32le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// payload bytes of a 64-sample block / 8, for header value b in 0..16 (RawData.cpp:27-45)
__device__ __forceinline__ uint32_t cur_len8(uint32_t b) {
    // nibbles for b = 0..10: 0,1,2,3,4,5,6,8,8,10,10 ; b >= 11 -> 16
    const unsigned long long lut = 0xAA886543210ull;
    uint32_t v = (uint32_t)(lut >> (4 * (b & 15))) & 15u;
    return b >= 11 ? 16u : v;
}

// --------------------------------------------------------------------------------------------------------
// k_meta
// --------------------------------------------------------------------------------------------------------
constexpr int K1_THREADS = 256;
constexpr int K1_CHUNK = 16384;   // bytes of the stream staged in shared memory per round
constexpr int K1_MB = 256;        // meta blocks per round (= threads, one per thread in the scan phase)

// One sample (index i = 8*j + l) of a block packed at header value b; scalar form of the SWAR recipe.
__device__ __forceinline__ uint32_t cur_sample_scalar(const uint8_t* p, uint32_t b, uint32_t i, const uint32_t* tab) {
    const uint32_t* e = tab + (b * 8 + (i >> 3)) * MCRAW_TAB_WORDS;
    const uint32_t l = i & 7;
    const uint32_t off = e[0], sh = e[1];
    if (off >> 24) return (uint32_t)p[2 * i] | ((uint32_t)p[2 * i + 1] << 8);
    uint32_t lo = 0, hi = 0;
    const uint32_t mA = e[2] & 0xFF, mBL = e[3] & 0xFF, mBH = e[4] & 0xFF, mC = e[5] & 0xFF;
    if (mA) lo |= ((uint32_t)p[(off & 0xFF) + l] >> (sh & 0xFF)) & mA;
    if (mBL | mBH) {
        uint32_t v = (uint32_t)p[((off >> 8) & 0xFF) + l] >> ((sh >> 8) & 0xFF);
        lo |= v & mBL;
        hi |= v & mBH;
    }
    if (mC) lo |= ((uint32_t)p[((off >> 16) & 0xFF) + l] >> ((sh >> 16) & 0xFF)) & mC;
    return lo | (hi << 8);
}

__global__ void __launch_bounds__(K1_THREADS) k_meta(FrameDev* __restrict__ frames, const uint32_t* __restrict__ tab_g) {
    __shared__ __align__(16) uint8_t stage[K1_CHUNK + 16];
    __shared__ __align__(16) uint8_t vals[K1_MB * 64];
    __shared__ uint16_t starts[K1_MB];
    __shared__ uint32_t tab[MCRAW_TAB_ENTRIES * MCRAW_TAB_WORDS];
    __shared__ uint32_t warp_sums[K1_THREADS / 32];
    __shared__ uint32_t sh_cnt, sh_err, sh_bad;
    __shared__ unsigned long long sh_nextpos;
    __shared__ uint32_t sh_hdr[4];

    const int f = blockIdx.x >> 1;
    const int stream = blockIdx.x & 1;   // 0 = bits, 1 = refs
    FrameDev& F = frames[f];
    if (F.type != MCRAW_COMPRESSION_CURRENT) return;
    const int tid = threadIdx.x;
    const uint8_t* __restrict__ src = F.src;
    const unsigned long long len = F.len;

    for (int i = tid; i < MCRAW_TAB_ENTRIES * MCRAW_TAB_WORDS; i += K1_THREADS) tab[i] = tab_g[i];

    if (tid == 0) {
        uint32_t err = 0;
        uint32_t ew = 0, eh = 0, boff = 0, roff = 0;
        if (len < 16) err = MCRAW_FRAME_BAD_HEADER;
        else {
            ew = ld_u32le(src); eh = ld_u32le(src + 4); boff = ld_u32le(src + 8); roff = ld_u32le(src + 12);
            if (boff > len || roff > len) err |= MCRAW_FRAME_BAD_HEADER;            // RawData.cpp:547
            if (ew % 64u) err |= MCRAW_FRAME_BAD_HEADER;                            // :550
            if (F.width <= 0 || ew < (uint32_t)F.width) err |= MCRAW_FRAME_BAD_HEADER;  // :553
            if (ew == 0 || eh == 0) err |= MCRAW_FRAME_BAD_HEADER;
            if (!err) {
                if (ew / 64u != F.tiles_x) err |= MCRAW_FRAME_GEOMETRY;
                if ((eh + 3u) / 4u > F.tile_rows) err |= MCRAW_FRAME_GEOMETRY;
            }
        }
        sh_hdr[0] = ew; sh_hdr[1] = eh; sh_hdr[2] = boff; sh_hdr[3] = roff;
        sh_err = err;
        sh_bad = 0;
        if (stream == 0) F.tile_rows_dev = err ? 0u : (eh + 3u) / 4u;
    }
    __syncthreads();
    if (sh_err) {
        if (tid == 0) atomicOr(&F.status, sh_err);
        return;
    }
    const uint32_t tiles_x = sh_hdr[0] / 64u;
    const uint32_t tile_rows = (sh_hdr[1] + 3u) / 4u;
    const uint32_t ntiles = tiles_x * tile_rows;
    const uint32_t nblocks = ntiles * 4u;
    const uint32_t need_mb = (nblocks + 63u) / 64u;
    unsigned long long pos = (unsigned long long)sh_hdr[2 + stream];

    if (tid == 0) {
        uint32_t err = 0;
        if (pos + 4 > len) err = MCRAW_FRAME_TRUNCATED;
        else if (ld_u32le(src + pos) < nblocks) err = MCRAW_FRAME_BAD_META_COUNT;  // RawData.cpp:470-476
        sh_err = err;
    }
    __syncthreads();
    if (sh_err) {
        if (tid == 0) atomicOr(&F.status, sh_err);
        return;
    }
    pos += 4;

    uint4* __restrict__ tilemeta = F.tilemeta;
    uint32_t done = 0;            // meta blocks finished
    uint32_t carry = 16;          // running payload offset, METADATA_OFFSET (RawData.cpp:25,562)

    while (done < need_mb) {
        // ---- stage [base, base + K1_CHUNK) of the frame buffer in shared memory (zero past len)
        const unsigned long long base = pos & ~15ull;
        for (int v = tid; v < K1_CHUNK / 16; v += K1_THREADS) {
            const unsigned long long o = base + (unsigned long long)v * 16;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (o + 16 <= len) q = __ldg(reinterpret_cast<const uint4*>(src + o));
            else if (o < len) {
                uint8_t tmp[16];
#pragma unroll
                for (int k = 0; k < 16; k++) tmp[k] = (o + k < len) ? src[o + k] : (uint8_t)0;
                q = *reinterpret_cast<uint4*>(tmp);
            }
            *reinterpret_cast<uint4*>(stage + v * 16) = q;
        }
        __syncthreads();
        // ---- serial chain walk over the inline 2-byte headers (RawData.cpp:485-489)
        if (tid == 0) {
            uint32_t p = (uint32_t)(pos - base), cnt = 0, err = 0;
            const uint32_t limit = min((uint32_t)K1_MB, need_mb - done);
            while (cnt < limit) {
                if (base + p + 2 > len) { err = MCRAW_FRAME_TRUNCATED; break; }
                if (p + 2 > K1_CHUNK) break;
                const uint32_t L = cur_len8(stage[p] >> 4) * 8u;
                if (base + p + 2 + L > len) { err = MCRAW_FRAME_TRUNCATED; break; }   // RawData.cpp:419
                if (p + 2 + L > K1_CHUNK) break;
                starts[cnt++] = (uint16_t)p;
                p += 2 + L;
            }
            sh_cnt = cnt;
            sh_err = err;
            sh_nextpos = base + p;
        }
        __syncthreads();
        const uint32_t cnt = sh_cnt;
        if (sh_err) break;
        // ---- unpack cnt meta blocks: value v -> meta block v>>6, sample v&63; + header reference (u16 wrap)
        for (uint32_t v = tid; v < cnt * 64u; v += K1_THREADS) {
            const uint32_t m = v >> 6, i = v & 63u;
            const uint8_t* h = stage + starts[m];
            const uint32_t b = h[0] >> 4;                                             // RawData.cpp:106-110
            const uint32_t ref = ((uint32_t)(h[0] & 0x0F) << 8) | h[1];
            const uint32_t val = (cur_sample_scalar(h + 2, b, i, tab) + ref) & 0xFFFFu; // :491-492
            const uint32_t k = (done + m) * 64u + i;                                  // block index
            if (stream == 0) {
                vals[v] = (uint8_t)min(val, 255u);
                if (k < nblocks && val > 16u) sh_bad = 1;                             // reference: OOB table read
            } else if (k < nblocks) {
                reinterpret_cast<uint16_t*>(tilemeta + (k >> 2))[4 + (k & 3u)] = (uint16_t)val;
            }
        }
        __syncthreads();
        if (stream == 0) {
            if (sh_bad) { if (tid == 0) sh_err = MCRAW_FRAME_BAD_BITS; __syncthreads(); break; }
            // ---- prefix sum of block lengths; thread t owns meta block t = 16 tiles
            uint32_t tsum[16];
            uint32_t bits4[16];
            uint32_t total = 0;
            if ((uint32_t)tid < cnt) {
                const uint4* vp = reinterpret_cast<const uint4*>(vals + tid * 64);
#pragma unroll
                for (int q4 = 0; q4 < 4; q4++) {
                    const uint4 q = vp[q4];
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t bb = w[t];
                        const uint32_t tile = (done + tid) * 16u + q4 * 4 + t;
                        uint32_t s = 0;
                        if (tile < ntiles)
                            s = 8u * (cur_len8(bb & 0xFF) + cur_len8((bb >> 8) & 0xFF) + cur_len8((bb >> 16) & 0xFF) + cur_len8(bb >> 24));
                        bits4[q4 * 4 + t] = bb;
                        tsum[q4 * 4 + t] = total;
                        total += s;
                    }
                }
            }
            // block-wide exclusive scan of `total`
            uint32_t incl = total;
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
            const uint32_t excl = carry + wbase + incl - total;
            if ((uint32_t)tid < cnt) {
#pragma unroll
                for (int t = 0; t < 16; t++) {
                    const uint32_t tile = (done + tid) * 16u + t;
                    if (tile < ntiles) {
                        uint2* dstp = reinterpret_cast<uint2*>(tilemeta + tile);
                        *dstp = make_uint2(excl + tsum[t], bits4[t]);
                    }
                }
            }
            carry += all;
            __syncthreads();
        }
        done += cnt;
        pos = sh_nextpos;
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t err = sh_err;
        if (!err && stream == 0 && (unsigned long long)carry > len) err = MCRAW_FRAME_TRUNCATED;  // RawData.cpp:419
        if (err) atomicOr(&F.status, err);
    }
}

// --------------------------------------------------------------------------------------------------------
// k_tiles
// --------------------------------------------------------------------------------------------------------
struct Plane8 {   // 8 samples (byte lanes 0..7) of one plane: low bytes and high bytes
    uint32_t lo0, lo1, hi0, hi1;
};

__device__ __forceinline__ uint2 ld_pay8(const uint8_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// Decode plane j of the block whose payload starts at `p` (8-byte aligned), header value b.
__device__ __forceinline__ Plane8 cur_decode_plane(const uint8_t* __restrict__ p, uint32_t b, uint32_t j, const uint32_t* tab) {
    const uint4 e0 = *reinterpret_cast<const uint4*>(tab + (b * 8 + j) * MCRAW_TAB_WORDS);       // off, sh, mA, mBL
    const uint2 e1 = *reinterpret_cast<const uint2*>(tab + (b * 8 + j) * MCRAW_TAB_WORDS + 4);   // mBH, mC
    const uint32_t off = e0.x, sh = e0.y;
    const uint32_t mA = e0.z, mBL = e0.w, mBH = e1.x, mC = e1.y;
    uint2 A = make_uint2(0, 0), B = make_uint2(0, 0), C = make_uint2(0, 0);
    if (mA) A = ld_pay8(p + (off & 0xFF));
    if (mBL | mBH) B = ld_pay8(p + ((off >> 8) & 0xFF));
    if (mC) C = ld_pay8(p + ((off >> 16) & 0xFF));
    Plane8 r;
    if (off >> 24) {   // 16-bit little-endian samples: de-interleave low/high bytes (RawData.cpp:376-408)
        r.lo0 = __byte_perm(A.x, A.y, 0x6420);
        r.hi0 = __byte_perm(A.x, A.y, 0x7531);
        r.lo1 = __byte_perm(B.x, B.y, 0x6420);
        r.hi1 = __byte_perm(B.x, B.y, 0x7531);
    } else {
        const uint32_t sA = sh & 31u, sB = (sh >> 8) & 31u, sC = (sh >> 16) & 31u;
        const uint32_t b0 = B.x >> sB, b1 = B.y >> sB;
        r.lo0 = ((A.x >> sA) & mA) | (b0 & mBL) | ((C.x >> sC) & mC);
        r.lo1 = ((A.y >> sA) & mA) | (b1 & mBL) | ((C.y >> sC) & mC);
        r.hi0 = b0 & mBH;
        r.hi1 = b1 & mBH;
    }
    return r;
}

__device__ __forceinline__ void st_vec16(uint16_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// grid = (max tile rows, frames); block = multiple of 32 chosen by the host to divide tiles_x*16 well.
__global__ void __launch_bounds__(256) k_tiles(FrameDev* __restrict__ frames, const uint32_t* __restrict__ tab_g,
                                               Result* __restrict__ results) {
    __shared__ __align__(16) uint32_t tab[MCRAW_TAB_ENTRIES * MCRAW_TAB_WORDS];
    const FrameDev& F = frames[blockIdx.y];
    if (F.type != MCRAW_COMPRESSION_CURRENT) return;
    const unsigned status = F.status;
    const uint32_t ty = blockIdx.x;
    const uint32_t tile_rows = F.tile_rows_dev;
    const int width = F.width;
    unsigned long long rows_fit = F.dst_cap / (unsigned long long)(width > 0 ? width : 1);
    if (rows_fit > 4ull * tile_rows) rows_fit = 4ull * tile_rows;
    if (ty == 0 && threadIdx.x == 0) {
        Result r;
        r.written = status ? 0ull : rows_fit * (unsigned long long)width;     // RawData.cpp:611
        r.status = status;
        r.pad = 0;
        results[blockIdx.y] = r;
    }
    if (status || ty >= tile_rows) return;
    for (int i = threadIdx.x; i < MCRAW_TAB_ENTRIES * MCRAW_TAB_WORDS; i += blockDim.x) tab[i] = tab_g[i];
    __syncthreads();

    const uint32_t tiles_x = F.tiles_x;
    const uint8_t* __restrict__ src = F.src;
    const uint4* __restrict__ tilemeta = F.tilemeta + (size_t)ty * tiles_x;
    uint16_t* __restrict__ dst = F.dst;
    const bool vec = (F.flags & FLAG_VEC_STORE) != 0;
    const uint32_t units = tiles_x * 16u;

    for (uint32_t u = threadIdx.x; u < units; u += blockDim.x) {
        const uint32_t tx = u >> 4, q = (u >> 3) & 1u, j = u & 7u;
        const uint4 tm = __ldg(tilemeta + tx);
        const uint32_t b0 = tm.y & 0xFF, b1 = (tm.y >> 8) & 0xFF, b2 = (tm.y >> 16) & 0xFF, b3 = tm.y >> 24;
        uint32_t offE = tm.x, bE, bO, refs;
        if (q == 0) { bE = b0; bO = b1; refs = tm.z; }
        else { bE = b2; bO = b3; refs = tm.w; offE += 8u * (cur_len8(b0) + cur_len8(b1)); }
        const uint32_t offO = offE + 8u * cur_len8(bE);

        const Plane8 E = cur_decode_plane(src + offE, bE, j, tab);
        const Plane8 O = cur_decode_plane(src + offO, bO, j, tab);

        // interleave even/odd Bayer columns and widen to u16 (RawData.cpp:581-593)
        const uint32_t x0 = __byte_perm(E.lo0, O.lo0, 0x5140), x1 = __byte_perm(E.lo0, O.lo0, 0x7362);
        const uint32_t x2 = __byte_perm(E.lo1, O.lo1, 0x5140), x3 = __byte_perm(E.lo1, O.lo1, 0x7362);
        const uint32_t y0 = __byte_perm(E.hi0, O.hi0, 0x5140), y1 = __byte_perm(E.hi0, O.hi0, 0x7362);
        const uint32_t y2 = __byte_perm(E.hi1, O.hi1, 0x5140), y3 = __byte_perm(E.hi1, O.hi1, 0x7362);
        uint32_t w[8];
        w[0] = __vadd2(__byte_perm(x0, y0, 0x5140), refs);
        w[1] = __vadd2(__byte_perm(x0, y0, 0x7362), refs);
        w[2] = __vadd2(__byte_perm(x1, y1, 0x5140), refs);
        w[3] = __vadd2(__byte_perm(x1, y1, 0x7362), refs);
        w[4] = __vadd2(__byte_perm(x2, y2, 0x5140), refs);
        w[5] = __vadd2(__byte_perm(x2, y2, 0x7362), refs);
        w[6] = __vadd2(__byte_perm(x3, y3, 0x5140), refs);
        w[7] = __vadd2(__byte_perm(x3, y3, 0x7362), refs);

        const unsigned long long row = 4ull * ty + q + 2u * (j >> 2);
        const int xpix = (int)(64u * tx + 16u * (j & 3u));
        if (row >= rows_fit || xpix >= width) continue;
        uint16_t* o = dst + row * (unsigned long long)width + xpix;
        if (vec) {
            st_vec16(o, w[0], w[1], w[2], w[3]);
            if (xpix + 8 < width) st_vec16(o + 8, w[4], w[5], w[6], w[7]);
        } else {
            const int n = min(16, width - xpix);
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (k < n) o[k] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
        }
    }
}

}  // namespace mcraw
