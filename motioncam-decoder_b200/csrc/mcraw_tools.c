/*
 * mcraw_tools.c -- CPU test-vector tools for the MCRAW frame codec (host only, plain C).
 *
 * What is here
 *   - mcraw_encode_current(): exact inverse of the current frame format (compressionType 7),
 *     i.e. it writes streams that /root/reference/lib/RawData.cpp:528-612 (raw::Decode) decodes
 *     back to the input image bit for bit.
 *   - mcraw_encode_legacy(): exact inverse of the legacy format (compressionType 6),
 *     /root/reference/lib/RawData_Legacy.cpp:445-495 (raw::DecodeLegacy).
 *   - deterministic synthetic Bayer generators ("photon", "flat+noise", uniform) built on
 *     splitmix64 so vectors reproduce on every host.
 *
 * The reference ships no encoder, no tests and no sample file; these tools are how every
 * parity vector in tests/ and every bench input is produced.  They never run on the decode
 * path of the product.
 *
 * Format facts used (cited against the reference decoder, which is the only specification):
 *   header            RawData.cpp:500-524   4 x u32 LE: encodedWidth, encodedHeight, bitsOffset, refsOffset
 *   payload lengths   RawData.cpp:27-45     {0,8,16,24,32,40,48,64,64,80,80,128...}
 *   block order       RawData.cpp:571-596   tile (64x4 px) row-major, 4 blocks per tile (Bayer phases)
 *   per-width layout  RawData.cpp:112-408   byte-lane planar groups of 8 bytes
 *   meta streams      RawData.cpp:463-498   u32 count, then [2-byte header][block] per 64 values
 *   legacy            RawData_Legacy.cpp:13-32,372-442,478-492
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* splitmix64 + helpers                                                                        */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t sm64_next(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Approximate N(0,1) * 2^16 as a fixed-point integer: sum of 12 uniform u16 minus 6*65536.
 * Integer only, so identical on every host. */
static inline int32_t sm64_gauss_q16(uint64_t* s) {
    uint64_t a = sm64_next(s), b = sm64_next(s), c = sm64_next(s);
    int32_t acc = 0;
    for (int i = 0; i < 4; i++) {
        acc += (int32_t)((a >> (16 * i)) & 0xFFFF);
        acc += (int32_t)((b >> (16 * i)) & 0xFFFF);
        acc += (int32_t)((c >> (16 * i)) & 0xFFFF);
    }
    return acc - 6 * 65536;
}

static inline uint32_t isqrt_u32(uint32_t v) {
    uint32_t r = 0, bit = 1u << 30;
    while (bit > v) bit >>= 2;
    while (bit) {
        if (v >= r + bit) { v -= r + bit; r = (r >> 1) + bit; }
        else r >>= 1;
        bit >>= 2;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* Synthetic images                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* "photon": smooth base in [64, ~0.94*maxval] + Gaussian noise sigma = 1 + 0.25*sqrt(base),
 * clipped to [0, maxval]; per-Bayer-phase gain so the four phases differ like R/G/G/B do. */
void mcraw_gen_photon(uint16_t* img, int width, int height, int maxval, uint64_t seed) {
    uint64_t s = seed * 0xD1342543DE82EF95ull + 0x1234567ull;
    const int gain_q8[4] = {150, 256, 256, 110}; /* R, G, G, B relative response (Q8) */
    const int span = (maxval * 15) / 16 - 64;
    /* a few low-frequency "blobs" make the base smooth but not a pure ramp */
    int cx[4], cy[4], amp[4];
    for (int k = 0; k < 4; k++) {
        cx[k] = (int)(sm64_next(&s) % (uint64_t)(width > 0 ? width : 1));
        cy[k] = (int)(sm64_next(&s) % (uint64_t)(height > 0 ? height : 1));
        amp[k] = (int)(sm64_next(&s) % 256);
    }
    const int64_t diag2 = (int64_t)width * width + (int64_t)height * height + 1;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            int64_t ramp = ((int64_t)x * 600) / (width > 1 ? width - 1 : 1) +
                           ((int64_t)y * 424) / (height > 1 ? height - 1 : 1); /* 0..1024 */
            int64_t blob = 0;
            for (int k = 0; k < 4; k++) {
                int64_t dx = x - cx[k], dy = y - cy[k];
                int64_t d2 = dx * dx + dy * dy;
                blob += (amp[k] * (diag2 - 4 * d2 > 0 ? diag2 - 4 * d2 : 0)) / diag2; /* 0..255 each */
            }
            int64_t t = (ramp * 3 + blob) % 4096;                 /* 0..4095 */
            if (((ramp * 3 + blob) / 4096) & 1) t = 4095 - t;     /* triangle fold, stays smooth */
            int phase = ((y & 1) << 1) | (x & 1);
            int64_t base = 64 + (t * span / 4096) * gain_q8[phase] / 256;
            uint32_t sig_q8 = 256 + 64 * isqrt_u32((uint32_t)base << 0) ; /* (1 + 0.25*sqrt(base)) in Q8 */
            int64_t n = ((int64_t)sm64_gauss_q16(&s) * (int64_t)sig_q8) >> 24; /* Q16*Q8 -> int */
            int64_t v = base + n;
            if (v < 0) v = 0;
            if (v > maxval) v = maxval;
            img[(size_t)y * width + x] = (uint16_t)v;
        }
    }
}

/* "flat+noise": cell x cell checkerboard of constant 64 (-> 0-bit blocks) versus
 * Gaussian(mean 511, sigma 170) clipped to [0,1023] (-> 10-bit blocks). */
void mcraw_gen_flatnoise(uint16_t* img, int width, int height, int cell, uint64_t seed) {
    uint64_t s = seed * 0xA0761D6478BD642Full + 0x7654321ull;
    if (cell <= 0) cell = 256;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            int flat = (((x / cell) + (y / cell)) & 1) == 0;
            int64_t v;
            if (flat) v = 64;
            else {
                v = 511 + (((int64_t)sm64_gauss_q16(&s) * 170) >> 16);
                if (v < 0) v = 0;
                if (v > 1023) v = 1023;
            }
            img[(size_t)y * width + x] = (uint16_t)v;
        }
    }
}

/* uniform in [lo, hi] */
void mcraw_gen_uniform(uint16_t* img, int width, int height, int lo, int hi, uint64_t seed) {
    uint64_t s = seed ^ 0xC0FFEE1234ull;
    uint64_t span = (uint64_t)(hi - lo + 1);
    for (size_t i = 0; i < (size_t)width * height; i++)
        img[i] = (uint16_t)(lo + (int)(sm64_next(&s) % span));
}

/* Per-tile forced widths: tile t (64x4 px, raster) gets values whose residual range needs exactly
 * widths[t % nwidths] bits in every one of its four blocks (0..16); base offsets vary so that
 * refs differ per block.  Used to hit every decode path incl. the 7/9/11..15 aliases. */
void mcraw_gen_forced_widths(uint16_t* img, int width, int height, const int* widths, int nwidths,
                             uint64_t seed) {
    uint64_t s = seed + 0x5151ull;
    int tiles_x = (width + 63) / 64;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            int t = (y / 4) * tiles_x + (x / 64);
            int w = widths[t % nwidths];
            int phase = ((y & 1) << 1) | (x & 1);
            uint32_t range = (w >= 16) ? 65535u : ((1u << w) - 1u);
            uint32_t basemax = 65535u - range;
            uint32_t base = basemax ? (uint32_t)((t * 37u + phase * 911u) % (basemax + 1u)) : 0u;
            if (base > 3000u && w < 16) base = base % 3001u;
            uint32_t r = (uint32_t)(sm64_next(&s) % ((uint64_t)range + 1u));
            /* make sure min (0) and max (range) both occur in each block: pin two samples */
            int j = (x % 64) / 2; /* sample index within the half block row */
            int half = ((y % 4) >> 1);
            if (j == 0 && half == 0) r = 0;
            if (j == 1 && half == 0) r = range;
            img[(size_t)y * width + x] = (uint16_t)(base + r);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Current format: block packers (inverse of RawData.cpp:112-408)                              */
/* ------------------------------------------------------------------------------------------ */
static const int CUR_LEN[17] = {0, 8, 16, 24, 32, 40, 48, 64, 64, 80, 80, 128, 128, 128, 128, 128, 128};

static inline int bitlen_u32(uint32_t v) {
    int n = 0;
    while (v) { n++; v >>= 1; }
    return n;
}

/* header value (0..16) that can carry residuals of `w` bits with the least bytes */
static inline int cur_min_header_for_width(int w) {
    if (w <= 6) return w;
    if (w <= 8) return (w == 7) ? 7 : 8;
    if (w <= 10) return (w == 9) ? 9 : 10;
    return w; /* 11..16 all mean "16-bit little-endian" */
}

/* v[64] residuals, hb = header value 0..16; writes CUR_LEN[hb] bytes; sample i = 8*j + l */
static void cur_pack_block(const uint16_t* v, int hb, uint8_t* out) {
    uint8_t* G = out;
#define VJ(j, l) ((uint32_t)v[8 * (j) + (l)])
    switch (hb) {
    case 0: break;
    case 1:
        for (int l = 0; l < 8; l++) {
            uint32_t g = 0;
            for (int j = 0; j < 8; j++) g |= (VJ(j, l) & 1u) << j;
            G[l] = (uint8_t)g;
        }
        break;
    case 2:
        for (int l = 0; l < 8; l++) {
            uint32_t g0 = 0, g1 = 0;
            for (int j = 0; j < 4; j++) { g0 |= (VJ(j, l) & 3u) << (2 * j); g1 |= (VJ(j + 4, l) & 3u) << (2 * j); }
            G[l] = (uint8_t)g0; G[8 + l] = (uint8_t)g1;
        }
        break;
    case 3:
        for (int l = 0; l < 8; l++) {
            uint32_t g0 = (VJ(0, l) & 7u) | ((VJ(1, l) & 7u) << 3) | ((VJ(2, l) & 3u) << 6);
            uint32_t g1 = (VJ(3, l) & 7u) | ((VJ(4, l) & 7u) << 3) | ((VJ(5, l) & 3u) << 6);
            uint32_t g2 = (VJ(6, l) & 7u) | ((VJ(7, l) & 7u) << 3) | (((VJ(2, l) >> 2) & 1u) << 6) | (((VJ(5, l) >> 2) & 1u) << 7);
            G[l] = (uint8_t)g0; G[8 + l] = (uint8_t)g1; G[16 + l] = (uint8_t)g2;
        }
        break;
    case 4:
        for (int l = 0; l < 8; l++)
            for (int m = 0; m < 4; m++)
                G[8 * m + l] = (uint8_t)((VJ(2 * m, l) & 15u) | ((VJ(2 * m + 1, l) & 15u) << 4));
        break;
    case 5:
        for (int l = 0; l < 8; l++) {
            uint32_t v5 = VJ(5, l), v6 = VJ(6, l), v7 = VJ(7, l);
            G[l]      = (uint8_t)((VJ(0, l) & 31u) | ((v5 & 7u) << 5));
            G[8 + l]  = (uint8_t)((VJ(1, l) & 31u) | ((v6 & 7u) << 5));
            G[16 + l] = (uint8_t)((VJ(2, l) & 31u) | ((v7 & 7u) << 5));
            G[24 + l] = (uint8_t)((VJ(3, l) & 31u) | (((v5 >> 3) & 3u) << 5) | (((v7 >> 3) & 1u) << 7));
            G[32 + l] = (uint8_t)((VJ(4, l) & 31u) | (((v6 >> 3) & 3u) << 5) | (((v7 >> 4) & 1u) << 7));
        }
        break;
    case 6:
        for (int l = 0; l < 8; l++) {
            uint32_t v6 = VJ(6, l), v7 = VJ(7, l);
            G[l]      = (uint8_t)((VJ(0, l) & 63u) | ((v6 & 3u) << 6));
            G[8 + l]  = (uint8_t)((VJ(1, l) & 63u) | (((v6 >> 2) & 3u) << 6));
            G[16 + l] = (uint8_t)((VJ(2, l) & 63u) | (((v6 >> 4) & 3u) << 6));
            G[24 + l] = (uint8_t)((VJ(3, l) & 63u) | ((v7 & 3u) << 6));
            G[32 + l] = (uint8_t)((VJ(4, l) & 63u) | (((v7 >> 2) & 3u) << 6));
            G[40 + l] = (uint8_t)((VJ(5, l) & 63u) | (((v7 >> 4) & 3u) << 6));
        }
        break;
    case 7:
    case 8:
        for (int j = 0; j < 8; j++)
            for (int l = 0; l < 8; l++) G[8 * j + l] = (uint8_t)VJ(j, l);
        break;
    case 9:
    case 10:
        for (int l = 0; l < 8; l++) {
            uint32_t h0 = 0, h1 = 0;
            for (int j = 0; j < 4; j++) {
                G[8 * j + l] = (uint8_t)(VJ(j, l) & 0xFFu);
                G[8 * (j + 5) + l] = (uint8_t)(VJ(j + 4, l) & 0xFFu);
                h0 |= ((VJ(j, l) >> 8) & 3u) << (2 * j);
                h1 |= ((VJ(j + 4, l) >> 8) & 3u) << (2 * j);
            }
            G[32 + l] = (uint8_t)h0;
            G[72 + l] = (uint8_t)h1;
        }
        break;
    default: /* 11..16: 64 x u16 little-endian */
        for (int i = 0; i < 64; i++) { G[2 * i] = (uint8_t)(v[i] & 0xFF); G[2 * i + 1] = (uint8_t)(v[i] >> 8); }
        break;
    }
#undef VJ
}

/* Encode one metadata stream (RawData.cpp:463-498): u32 LE count (padded to a multiple of 64),
 * then per 64 values [b0 b1][payload].  Header reference is 12 bits, header bits nibble 0..15.
 * alias_seed != 0 randomly picks wider-than-needed headers (7/8, 9/10, 11..15). */
static size_t cur_write_meta_stream(const uint16_t* vals, size_t n, uint8_t* out, uint64_t alias_seed) {
    size_t m = (n + 63) / 64 * 64;
    uint8_t* p = out;
    p[0] = (uint8_t)(m & 0xFF); p[1] = (uint8_t)((m >> 8) & 0xFF); p[2] = (uint8_t)((m >> 16) & 0xFF); p[3] = (uint8_t)((m >> 24) & 0xFF);
    p += 4;
    uint64_t s = alias_seed;
    for (size_t base = 0; base < m; base += 64) {
        uint16_t blk[64];
        uint32_t mn = 0xFFFFu, mx = 0;
        for (int i = 0; i < 64; i++) {
            /* padding past n repeats the last real value so it never widens the block */
            blk[i] = (base + i < n) ? vals[base + i] : (n ? vals[n - 1] : 0);
            if (blk[i] < mn) mn = blk[i];
            if (blk[i] > mx) mx = blk[i];
        }
        uint32_t ref = mn > 0xFFFu ? 0xFFFu : mn;
        int w = bitlen_u32(mx - ref);
        int hb = cur_min_header_for_width(w);
        if (hb > 15) hb = 15; /* nibble; 11..15 all take the 16-bit path */
        if (alias_seed) {
            uint64_t r = sm64_next(&s);
            if (hb == 7 && (r & 1)) hb = 8;
            else if (hb == 9 && (r & 1)) hb = 10;
            else if (hb >= 11) hb = 11 + (int)(r % 5);
        }
        for (int i = 0; i < 64; i++) blk[i] = (uint16_t)(blk[i] - ref);
        p[0] = (uint8_t)((hb << 4) | ((ref >> 8) & 0xF));
        p[1] = (uint8_t)(ref & 0xFF);
        p += 2;
        cur_pack_block(blk, hb, p);
        p += CUR_LEN[hb];
    }
    return (size_t)(p - out);
}

/* Worst-case encoded size for a width x height frame (current format). */
size_t mcraw_encode_current_bound(int width, int height) {
    size_t ew = ((size_t)width + 63) / 64 * 64, eh = ((size_t)height + 3) / 4 * 4;
    size_t nblocks = ew * eh / 64;
    size_t meta = 4 + ((nblocks + 63) / 64) * (2 + 128);
    return 16 + nblocks * 128 + 2 * meta + 64;
}

/*
 * Encode a frame in the current format.
 *   img            width x height u16, row-major
 *   policy         0 = minimal header per block
 *                  1 = random legal aliases (7<->8, 9<->10, 11..16) and occasional wider-than-needed headers
 *                  2 = every block uses header value `policy_arg` if it is wide enough, else minimal
 *   ref_wrap       when non-zero, some blocks store ref = min - delta (mod 2^16) and residual + delta so that
 *                  residual + ref wraps past 65535 (RawData.cpp:582-592 stores into uint16_t)
 * Returns bytes written, 0 on error (capacity too small / bad geometry).
 * encodedWidth = roundup(width,64); encodedHeight = roundup(height,4) (callers that want streams the
 * reference can decode safely pass height % 4 == 0).
 */
size_t mcraw_encode_current(const uint16_t* img, int width, int height, uint8_t* out, size_t cap,
                            int policy, int policy_arg, int ref_wrap, uint64_t seed) {
    if (width <= 0 || height <= 0) return 0;
    const int ew = (width + 63) / 64 * 64, eh = (height + 3) / 4 * 4;
    const size_t tiles_x = (size_t)ew / 64, tiles_y = (size_t)eh / 4;
    const size_t nblocks = tiles_x * tiles_y * 4;
    if (cap < mcraw_encode_current_bound(width, height)) return 0;
    uint16_t* bits = (uint16_t*)malloc(nblocks * sizeof(uint16_t));
    uint16_t* refs = (uint16_t*)malloc(nblocks * sizeof(uint16_t));
    if (!bits || !refs) { free(bits); free(refs); return 0; }
    uint64_t s = seed * 0x2545F4914F6CDD1Dull + 99;
    uint8_t* p = out + 16;
    size_t k = 0;
    for (size_t ty = 0; ty < tiles_y; ty++) {
        for (size_t tx = 0; tx < tiles_x; tx++) {
            for (int c = 0; c < 4; c++, k++) {
                uint16_t v[64];
                uint32_t mn = 0xFFFFu, mx = 0;
                for (int i = 0; i < 64; i++) {
                    int j = i & 31, half = i >> 5;
                    int y = (int)(4 * ty) + (c >> 1) + 2 * half;
                    int x = (int)(64 * tx) + 2 * j + (c & 1);
                    /* pad by replicating the nearest same-phase pixel */
                    while (x >= width) x -= 2;
                    while (y >= height) y -= 2;
                    if (x < 0) x = (c & 1) < width ? (c & 1) : 0;
                    if (y < 0) y = 0;
                    v[i] = img[(size_t)y * width + x];
                    if (v[i] < mn) mn = v[i];
                    if (v[i] > mx) mx = v[i];
                }
                int w = bitlen_u32(mx - mn);
                int hb = cur_min_header_for_width(w);
                uint32_t ref = mn;
                uint64_t r = (policy == 1 || ref_wrap) ? sm64_next(&s) : 0;
                if (policy == 1) {
                    if (hb == 7 && (r & 1)) hb = 8;
                    else if (hb == 9 && (r & 1)) hb = 10;
                    else if (hb >= 11) hb = 11 + (int)((r >> 1) % 6);
                    else if (((r >> 8) & 31) == 0 && hb < 16) hb = hb + 1 + (int)((r >> 16) % (uint64_t)(16 - hb)); /* wider than needed */
                } else if (policy == 2) {
                    int want = policy_arg;
                    int cap_w = (want >= 11) ? 16 : (want == 7 ? 8 : (want == 9 ? 10 : want));
                    if (want >= 0 && want <= 16 && cap_w >= w) hb = want;
                }
                uint32_t delta = 0;
                if (ref_wrap && ((r >> 40) & 3) == 0) {
                    /* room left in the chosen width */
                    int cap_w = (hb >= 11) ? 16 : (hb == 7 ? 8 : (hb == 9 ? 10 : hb));
                    uint32_t room = ((cap_w >= 16) ? 65535u : ((1u << cap_w) - 1u)) - (mx - mn);
                    if (room) delta = 1 + (uint32_t)((r >> 44) % room);
                }
                ref = (mn - delta) & 0xFFFFu;
                for (int i = 0; i < 64; i++) v[i] = (uint16_t)(v[i] - mn + delta);
                bits[k] = (uint16_t)hb;
                refs[k] = (uint16_t)ref;
                cur_pack_block(v, hb, p);
                p += CUR_LEN[hb];
            }
        }
    }
    uint32_t bits_off = (uint32_t)(p - out);
    p += cur_write_meta_stream(bits, nblocks, p, policy == 1 ? (seed | 1) : 0);
    uint32_t refs_off = (uint32_t)(p - out);
    p += cur_write_meta_stream(refs, nblocks, p, policy == 1 ? (seed * 3 | 1) : 0);
    uint32_t hdr[4] = {(uint32_t)ew, (uint32_t)eh, bits_off, refs_off};
    for (int i = 0; i < 4; i++) {
        out[4 * i + 0] = (uint8_t)(hdr[i] & 0xFF);
        out[4 * i + 1] = (uint8_t)((hdr[i] >> 8) & 0xFF);
        out[4 * i + 2] = (uint8_t)((hdr[i] >> 16) & 0xFF);
        out[4 * i + 3] = (uint8_t)((hdr[i] >> 24) & 0xFF);
    }
    free(bits);
    free(refs);
    return (size_t)(p - out);
}

/*
 * Assemble a current-format frame directly from caller-supplied bits[]/refs[] and random payload
 * bytes ("directly randomised well-formed stream").  bits[k] in 0..16.
 */
size_t mcraw_assemble_current(int enc_width, int enc_height, const uint16_t* bits, const uint16_t* refs,
                              uint8_t* out, size_t cap, uint64_t seed) {
    if (enc_width <= 0 || enc_height <= 0 || enc_width % 64 || enc_height % 4) return 0;
    size_t nblocks = (size_t)enc_width * enc_height / 64;
    if (cap < mcraw_encode_current_bound(enc_width, enc_height)) return 0;
    uint64_t s = seed ^ 0xABCDEF01ull;
    uint8_t* p = out + 16;
    for (size_t k = 0; k < nblocks; k++) {
        int n = CUR_LEN[bits[k] > 16 ? 16 : bits[k]];
        for (int i = 0; i < n; i += 8) {
            uint64_t r = sm64_next(&s);
            memcpy(p + i, &r, 8);
        }
        p += n;
    }
    uint32_t bits_off = (uint32_t)(p - out);
    p += cur_write_meta_stream(bits, nblocks, p, seed | 1);
    uint32_t refs_off = (uint32_t)(p - out);
    p += cur_write_meta_stream(refs, nblocks, p, (seed * 7) | 1);
    uint32_t hdr[4] = {(uint32_t)enc_width, (uint32_t)enc_height, bits_off, refs_off};
    for (int i = 0; i < 4; i++) {
        out[4 * i + 0] = (uint8_t)(hdr[i] & 0xFF);
        out[4 * i + 1] = (uint8_t)((hdr[i] >> 8) & 0xFF);
        out[4 * i + 2] = (uint8_t)((hdr[i] >> 16) & 0xFF);
        out[4 * i + 3] = (uint8_t)((hdr[i] >> 24) & 0xFF);
    }
    return (size_t)(p - out);
}

/* ------------------------------------------------------------------------------------------ */
/* Legacy format (inverse of RawData_Legacy.cpp:38-442)                                        */
/* ------------------------------------------------------------------------------------------ */
static const int LEG_LEN[17] = {0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 32, 32, 32, 32, 32, 32};

size_t mcraw_encode_legacy_bound(int width, int height) {
    size_t pw = ((size_t)width + 31) / 32 * 32;
    return (size_t)height * (pw / 16) * 34 + 16;
}

/* 16 residuals at w bits, MSB-first contiguous bitstream; w >= 11 -> 16-bit big-endian */
static size_t leg_pack_block(const uint16_t* v, int hb, uint8_t* out) {
    if (hb >= 11) {
        for (int i = 0; i < 16; i++) { out[2 * i] = (uint8_t)(v[i] >> 8); out[2 * i + 1] = (uint8_t)(v[i] & 0xFF); }
        return 32;
    }
    int nbytes = 2 * hb;
    memset(out, 0, (size_t)nbytes);
    int bitpos = 0;
    for (int i = 0; i < 16; i++) {
        for (int b = hb - 1; b >= 0; b--, bitpos++) {
            if ((v[i] >> b) & 1) out[bitpos >> 3] |= (uint8_t)(0x80u >> (bitpos & 7));
        }
    }
    return (size_t)nbytes;
}

/*
 * Encode a frame in the legacy format.  policy 0 = minimal header, 1 = random aliases for >=11 and
 * occasional wider-than-needed headers, 2 = force header nibble policy_arg when wide enough.
 * trailer_records > 0 appends that many [pos u32 BE][0xFF] restart records (RawData_Legacy.cpp:451-469,
 * parsed but ignored by the reference); otherwise a single 0x00 trailing byte is appended because the
 * reference's end checks use ">=" (RawData_Legacy.cpp:387,398).
 */
size_t mcraw_encode_legacy(const uint16_t* img, int width, int height, uint8_t* out, size_t cap,
                           int policy, int policy_arg, int trailer_records, uint64_t seed) {
    if (width <= 0 || height <= 0) return 0;
    const int pw = (width + 31) / 32 * 32;
    if (cap < mcraw_encode_legacy_bound(width, height) + (size_t)trailer_records * 5) return 0;
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 7;
    uint8_t* p = out;
    for (int y = 0; y < height; y++) {
        for (int x0 = 0; x0 < pw; x0 += 32) {
            for (int c = 0; c < 2; c++) {
                uint16_t v[16];
                uint32_t mn = 0xFFFFu, mx = 0;
                for (int j = 0; j < 16; j++) {
                    int x = x0 + 2 * j + c;
                    while (x >= width) x -= 2;
                    if (x < 0) x = 0;
                    v[j] = img[(size_t)y * width + x];
                    if (v[j] < mn) mn = v[j];
                    if (v[j] > mx) mx = v[j];
                }
                uint32_t ref = mn > 0xFFFu ? 0xFFFu : mn; /* 12-bit header reference */
                int w = bitlen_u32(mx - ref);
                int hb = (w <= 10) ? w : (w > 15 ? 15 : w);
                if (policy == 1) {
                    uint64_t r = sm64_next(&s);
                    if (hb >= 11) hb = 11 + (int)(r % 5);
                    else if (((r >> 8) & 31) == 0 && hb < 15) hb = hb + 1 + (int)((r >> 16) % (uint64_t)(15 - hb));
                } else if (policy == 2) {
                    int want = policy_arg;
                    int cap_w = want >= 11 ? 16 : want;
                    if (want >= 0 && want <= 15 && cap_w >= w) hb = want;
                }
                for (int j = 0; j < 16; j++) v[j] = (uint16_t)(v[j] - ref);
                p[0] = (uint8_t)((hb << 4) | ((ref >> 8) & 0xF));
                p[1] = (uint8_t)(ref & 0xFF);
                p += 2;
                p += leg_pack_block(v, hb, p);
            }
        }
    }
    if (trailer_records > 0) {
        /* a non-0xFF pad byte first so that the backwards scan stops on it, then the records */
        *p++ = 0x00;
        size_t payload = (size_t)(p - out);
        for (int i = 0; i < trailer_records; i++) {
            uint32_t pos = (uint32_t)((payload / (size_t)(trailer_records + 1)) * (size_t)(i + 1)) & ~1u;
            p[0] = (uint8_t)(pos >> 24); p[1] = (uint8_t)(pos >> 16); p[2] = (uint8_t)(pos >> 8); p[3] = (uint8_t)pos;
            p[4] = 0xFF;
            p += 5;
        }
    } else {
        *p++ = 0x00;
    }
    (void)LEG_LEN;
    return (size_t)(p - out);
}

/* Random well-formed legacy stream: random header nibbles/refs and random payload bytes. */
size_t mcraw_assemble_legacy(int width, int height, const uint8_t* nibbles, const uint16_t* refs12,
                             uint8_t* out, size_t cap, uint64_t seed) {
    const int pw = (width + 31) / 32 * 32;
    size_t nblocks = (size_t)height * (size_t)(pw / 16);
    if (cap < mcraw_encode_legacy_bound(width, height)) return 0;
    uint64_t s = seed ^ 0x13572468ull;
    uint8_t* p = out;
    for (size_t k = 0; k < nblocks; k++) {
        int hb = nibbles[k] & 15;
        p[0] = (uint8_t)((hb << 4) | ((refs12[k] >> 8) & 0xF));
        p[1] = (uint8_t)(refs12[k] & 0xFF);
        p += 2;
        int n = LEG_LEN[hb];
        for (int i = 0; i < n; i += 2) {
            uint64_t r = sm64_next(&s);
            p[i] = (uint8_t)r; p[i + 1] = (uint8_t)(r >> 8);
        }
        p += n;
    }
    *p++ = 0x00;
    return (size_t)(p - out);
}

/* 64-bit FNV-1a over a byte buffer -- the checksum used by the full-size property tests. */
uint64_t mcraw_fnv1a64(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xCBF29CE484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001B3ull; }
    return h;
}

#ifdef __cplusplus
}
#endif
