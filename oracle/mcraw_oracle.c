/*
 * mcraw_oracle.c -- CPU restatement of the reference MCRAW frame codec.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA decode path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * It restates, as plain scalar C (one sample at a time, no SIMD, no tables shared with the product):
 *   oracle_decode()         <- /root/reference/lib/RawData.cpp:528-612       (motioncam::raw::Decode)
 *   oracle_decode_legacy()  <- /root/reference/lib/RawData_Legacy.cpp:445-495 (motioncam::raw::DecodeLegacy)
 *
 * Pinning: the reference has no tests or golden vectors, so parity is pinned by
 * executing the reference itself: oracle/Makefile compiles the unmodified reference sources into
 * oracle/_ref/libmcraw_ref.so, tests/test_oracle_vs_ref.py checks this restatement against it on every
 * vector family, and tests/golden/ holds fixtures generated from that library (tests/golden/make_golden.py).
 *
 * Deliberate differences from the reference, all in corners where the reference has undefined behaviour
 *  -- none reachable by well-formed streams:
 *   - truncated payload/metadata: reference skips the block and leaves stale stack data
 *     (RawData.cpp:419-420); here the frame fails (returns 0).
 *   - header bits value > 16 in the bits stream: reference indexes past its table (RawData.cpp:419); here 0.
 *   - metadata count < number of blocks / not a multiple of 64: reference overflows a vector
 *     (RawData.cpp:476,485-495); here count >= blocks is required, any remainder is accepted.
 *   - encodedHeight rows are emitted like the reference (RawData.cpp:571,598-608) but never more than
 *     out_capacity_elems allows (the reference would overflow the caller's buffer).
 *   - len < 16 (current) or len == 0 (legacy): reference reads out of bounds; here 0.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RawData.cpp:27-45 */
static const int kLen[17] = {0, 8, 16, 24, 32, 40, 48, 64, 64, 80, 80, 128, 128, 128, 128, 128, 128};

static uint32_t rd_u32le(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* One sample of a block, index i = 8*j + l, from payload `in` packed at header value `bits`.
 * RawData.cpp:112-408 (Decode1..Decode16), dispatch RawData.cpp:424-458. */
static uint16_t cur_sample(const uint8_t* in, int bits, int i) {
    const int j = i >> 3, l = i & 7;
#define G(m) ((uint32_t)in[8 * (m) + l])
    switch (bits) {
    case 0: return 0;                                                     /* :425-427 */
    case 1: return (uint16_t)((G(0) >> j) & 1u);                          /* :112-136 */
    case 2: return (uint16_t)((G(j >> 2) >> (2 * (j & 3))) & 3u);         /* :138-162 */
    case 3:                                                               /* :164-199 */
        switch (j) {
        case 0: return (uint16_t)(G(0) & 7u);
        case 1: return (uint16_t)((G(0) >> 3) & 7u);
        case 2: return (uint16_t)(((G(0) >> 6) & 3u) | (((G(2) >> 6) & 1u) << 2));
        case 3: return (uint16_t)(G(1) & 7u);
        case 4: return (uint16_t)((G(1) >> 3) & 7u);
        case 5: return (uint16_t)(((G(1) >> 6) & 3u) | (((G(2) >> 7) & 1u) << 2));
        case 6: return (uint16_t)(G(2) & 7u);
        default: return (uint16_t)((G(2) >> 3) & 7u);
        }
    case 4: return (uint16_t)((G(j >> 1) >> (4 * (j & 1))) & 15u);        /* :201-223 */
    case 5:                                                               /* :225-262 */
        if (j <= 4) return (uint16_t)(G(j) & 31u);
        if (j == 5) return (uint16_t)(((G(0) >> 5) & 7u) | (((G(3) >> 5) & 3u) << 3));
        if (j == 6) return (uint16_t)(((G(1) >> 5) & 7u) | (((G(4) >> 5) & 3u) << 3));
        return (uint16_t)(((G(2) >> 5) & 7u) | (((G(3) >> 7) & 1u) << 3) | (((G(4) >> 7) & 1u) << 4));
    case 6:                                                               /* :264-304 */
        if (j <= 5) return (uint16_t)(G(j) & 63u);
        if (j == 6) return (uint16_t)(((G(0) >> 6) & 3u) | (((G(1) >> 6) & 3u) << 2) | (((G(2) >> 6) & 3u) << 4));
        return (uint16_t)(((G(3) >> 6) & 3u) | (((G(4) >> 6) & 3u) << 2) | (((G(5) >> 6) & 3u) << 4));
    case 7:
    case 8: return (uint16_t)G(j);                                        /* :306-326,446-449 */
    case 9:
    case 10:                                                              /* :328-374,450-453 */
        if (j < 4) return (uint16_t)(G(j) | (((G(4) >> (2 * j)) & 3u) << 8));
        return (uint16_t)(G(j + 1) | (((G(9) >> (2 * (j - 4))) & 3u) << 8));
    default:                                                              /* :376-408, little-endian host */
        return (uint16_t)((uint32_t)in[2 * i] | ((uint32_t)in[2 * i + 1] << 8));
    }
#undef G
}

/* RawData.cpp:463-498.  Decodes the first `need` values (rounded up to whole 64-value blocks) of the
 * stream at `offset`.  Returns 0 on success. */
static int cur_decode_meta(const uint8_t* in, size_t len, size_t offset, size_t need, uint16_t* out /* roundup(need,64) */) {
    if (offset + 4 > len) return -1;
    uint32_t count = rd_u32le(in + offset);                               /* :470-474 */
    if ((size_t)count < need) return -1;
    offset += 4;
    for (size_t i = 0; i < need; i += 64) {
        if (offset + 2 > len) return -1;
        int bits = (in[offset] >> 4) & 0x0F;                              /* :106-110 */
        uint32_t ref = ((uint32_t)(in[offset] & 0x0F) << 8) | in[offset + 1];
        offset += 2;
        if (offset + (size_t)kLen[bits] > len) return -1;                 /* :419 (reference: skip, stale) */
        for (int x = 0; x < 64; x++)
            out[i + x] = (uint16_t)(cur_sample(in + offset, bits, x) + ref); /* :491-492, u16 wrap */
        offset += (size_t)kLen[bits];
    }
    return 0;
}

size_t oracle_decode(uint16_t* output, size_t out_capacity_elems, int width, int height,
                     const uint8_t* input, size_t len) {
    (void)height;                                                         /* unused by the reference too (:531) */
    if (!input || !output || len < 16 || width <= 0) return 0;
    const uint32_t ew = rd_u32le(input), eh = rd_u32le(input + 4);        /* :500-524 */
    const uint32_t bits_off = rd_u32le(input + 8), refs_off = rd_u32le(input + 12);
    if (bits_off > len || refs_off > len) return 0;                       /* :547-548 */
    if (ew % 64) return 0;                                                /* :550-551 */
    if (ew < (uint32_t)width) return 0;                                   /* :553-554 */
    if (ew == 0 || eh == 0) return 0;
    const size_t tiles_x = ew / 64;
    const size_t tile_rows = ((size_t)eh + 3) / 4;                        /* loop :571 runs while y < encodedHeight */
    const size_t nblocks = tiles_x * tile_rows * 4;
    const size_t padded = (nblocks + 63) / 64 * 64;
    uint16_t* bits = (uint16_t*)malloc(padded * sizeof(uint16_t));
    uint16_t* refs = (uint16_t*)malloc(padded * sizeof(uint16_t));
    size_t written = 0;
    if (!bits || !refs) goto done;
    if (cur_decode_meta(input, len, bits_off, nblocks, bits)) goto done;  /* :557 */
    if (cur_decode_meta(input, len, refs_off, nblocks, refs)) goto done;  /* :560 */
    {
        size_t offset = 16;                                               /* :562 */
        size_t k = 0;
        const size_t rows_total = tile_rows * 4;                          /* reference emits 4 rows per iteration (:598-608) */
        size_t rows_fit = out_capacity_elems / (size_t)width;
        if (rows_fit > rows_total) rows_fit = rows_total;
        for (size_t ty = 0; ty < tile_rows; ty++) {
            for (size_t tx = 0; tx < tiles_x; tx++) {
                for (int c = 0; c < 4; c++, k++) {
                    const int b = bits[k];
                    if (b > 16) goto done;
                    if (offset + (size_t)kLen[b] > len) goto done;        /* :419 */
                    for (int i = 0; i < 64; i++) {
                        /* :581-593: p_c[i/2-th] -> rows 0,2 from p0/p1, rows 1,3 from p2/p3 */
                        const size_t y = 4 * ty + (size_t)(c >> 1) + 2 * (size_t)(i >> 5);
                        const size_t x = 64 * tx + 2 * (size_t)(i & 31) + (size_t)(c & 1);
                        if (x < (size_t)width && y < rows_fit)
                            output[y * (size_t)width + x] = (uint16_t)(cur_sample(input + offset, b, i) + refs[k]);
                    }
                    offset += (size_t)kLen[b];
                }
            }
        }
        written = rows_fit * (size_t)width;                               /* :611 */
    }
done:
    free(bits);
    free(refs);
    return written;
}

/* ------------------------------------------------------------------------------------------ */
/* Legacy                                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* RawData_Legacy.cpp:13-32 */
static const int kLegLen[17] = {0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 32, 32, 32, 32, 32, 32};

/* Sample k (0..15) of a w-bit MSB-first contiguous bitstream; RawData_Legacy.cpp:38-358 restated as
 * "sample k occupies stream bits [k*w, (k+1)*w)". */
static uint16_t leg_sample(const uint8_t* in, int w, int k) {
    uint32_t v = 0;
    int bit = k * w;
    for (int n = 0; n < w; n++, bit++)
        v = (v << 1) | ((in[bit >> 3] >> (7 - (bit & 7))) & 1u);
    return (uint16_t)v;
}

size_t oracle_decode_legacy(uint16_t* output, size_t out_capacity_elems, int width, int height,
                            const uint8_t* input, size_t len) {
    if (!input || !output || len == 0 || width <= 0 || height <= 0) return 0;
    if (out_capacity_elems < (size_t)width * (size_t)height) return 0;
    const int pw = 32 * ((width + 31) / 32);                              /* :34-36,449 */
    /* the trailer scan (:455-469) has no effect on the output and is not restated */
    size_t offset = 0;
    uint16_t p[32];
    memset(p, 0, sizeof p);
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < pw; x += 32) {
            uint16_t ref[2] = {0, 0};
            for (int c = 0; c < 2; c++) {
                /* RawData_Legacy.cpp:377-442 */
                if (offset + 2 >= len) return 0;                          /* :387 (reference: stale data) */
                int bits = (input[offset] >> 4) & 0x0F;                   /* :372-375 */
                ref[c] = (uint16_t)(((uint32_t)(input[offset] & 0x0F) << 8) | input[offset + 1]);
                if (bits > 16) bits = 16;                                 /* :395 */
                if (offset + 2 + (size_t)kLegLen[bits] >= len) return 0;  /* :398 */
                const uint8_t* q = input + offset + 2;
                for (int j = 0; j < 16; j++) {
                    if (bits == 0) p[16 * c + j] = 0;                     /* :402-404 */
                    else if (bits <= 10) p[16 * c + j] = leg_sample(q, bits, j);   /* :405-434 */
                    else p[16 * c + j] = (uint16_t)(((uint32_t)q[2 * j] << 8) | q[2 * j + 1]); /* :360-370 big-endian */
                }
                offset += 2 + (size_t)kLegLen[bits];                      /* :441 */
            }
            for (int i = 0; i < 32; i += 2) {                             /* :483-486 */
                if (x + i < width) output[(size_t)y * width + x + i] = (uint16_t)(p[i / 2] + ref[0]);
                if (x + i + 1 < width) output[(size_t)y * width + x + i + 1] = (uint16_t)(p[16 + i / 2] + ref[1]);
            }
        }
    }
    return (size_t)width * (size_t)height;                                /* :494 */
}

#ifdef __cplusplus
}
#endif
