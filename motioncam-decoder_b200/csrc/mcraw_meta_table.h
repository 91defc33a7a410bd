/* mcraw_meta_table.h -- table-driven extraction of single values from a 64-value block of the current format.
 *
 * Sample i = 8*j + b of a block sits in byte lane b of one to three 8-byte groups G_g (RawData.cpp:112-374, restated in
 * SURVEY.md appendix A):  value = OR over terms t of  ((G_g[b] >> s) & m) << d.   One table row per (header value 0..10,
 * plane j): three terms packed as  g | s << 4 | d << 8 | m << 16  (m = 0: unused).  Header values 11..15 are the 16-bit
 * path (two little-endian bytes per sample, RawData.cpp:376-408) and have no row.
 *
 * Used by k_units, where every lane needs just TWO values (its block pair) of each metadata block: a lane-uniform
 * formula with per-lane table rows instead of a per-plane switch.  Plain C so that tests/ can check it on the CPU
 * against the oracle for every (header, sample).
 */
#ifndef MCRAW_META_TABLE_H
#define MCRAW_META_TABLE_H
#include <stdint.h>

#define MC_T(g, s, m, d) ((uint32_t)(g) | ((uint32_t)(s) << 4) | ((uint32_t)(d) << 8) | ((uint32_t)(m) << 16))
#define MC_T0 0u

/* [header value 0..10][plane j 0..7][term 0..2] */
#define MCRAW_META_TERMS_INIT { \
  /* 0 */ {{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0},{MC_T0,MC_T0,MC_T0}}, \
  /* 1 */ {{MC_T(0,0,1,0),MC_T0,MC_T0},{MC_T(0,1,1,0),MC_T0,MC_T0},{MC_T(0,2,1,0),MC_T0,MC_T0},{MC_T(0,3,1,0),MC_T0,MC_T0},{MC_T(0,4,1,0),MC_T0,MC_T0},{MC_T(0,5,1,0),MC_T0,MC_T0},{MC_T(0,6,1,0),MC_T0,MC_T0},{MC_T(0,7,1,0),MC_T0,MC_T0}}, \
  /* 2 */ {{MC_T(0,0,3,0),MC_T0,MC_T0},{MC_T(0,2,3,0),MC_T0,MC_T0},{MC_T(0,4,3,0),MC_T0,MC_T0},{MC_T(0,6,3,0),MC_T0,MC_T0},{MC_T(1,0,3,0),MC_T0,MC_T0},{MC_T(1,2,3,0),MC_T0,MC_T0},{MC_T(1,4,3,0),MC_T0,MC_T0},{MC_T(1,6,3,0),MC_T0,MC_T0}}, \
  /* 3 */ {{MC_T(0,0,7,0),MC_T0,MC_T0},{MC_T(0,3,7,0),MC_T0,MC_T0},{MC_T(0,6,3,0),MC_T(2,6,1,2),MC_T0},{MC_T(1,0,7,0),MC_T0,MC_T0},{MC_T(1,3,7,0),MC_T0,MC_T0},{MC_T(1,6,3,0),MC_T(2,7,1,2),MC_T0},{MC_T(2,0,7,0),MC_T0,MC_T0},{MC_T(2,3,7,0),MC_T0,MC_T0}}, \
  /* 4 */ {{MC_T(0,0,15,0),MC_T0,MC_T0},{MC_T(0,4,15,0),MC_T0,MC_T0},{MC_T(1,0,15,0),MC_T0,MC_T0},{MC_T(1,4,15,0),MC_T0,MC_T0},{MC_T(2,0,15,0),MC_T0,MC_T0},{MC_T(2,4,15,0),MC_T0,MC_T0},{MC_T(3,0,15,0),MC_T0,MC_T0},{MC_T(3,4,15,0),MC_T0,MC_T0}}, \
  /* 5 */ {{MC_T(0,0,31,0),MC_T0,MC_T0},{MC_T(1,0,31,0),MC_T0,MC_T0},{MC_T(2,0,31,0),MC_T0,MC_T0},{MC_T(3,0,31,0),MC_T0,MC_T0},{MC_T(4,0,31,0),MC_T0,MC_T0},{MC_T(0,5,7,0),MC_T(3,5,3,3),MC_T0},{MC_T(1,5,7,0),MC_T(4,5,3,3),MC_T0},{MC_T(2,5,7,0),MC_T(3,7,1,3),MC_T(4,7,1,4)}}, \
  /* 6 */ {{MC_T(0,0,63,0),MC_T0,MC_T0},{MC_T(1,0,63,0),MC_T0,MC_T0},{MC_T(2,0,63,0),MC_T0,MC_T0},{MC_T(3,0,63,0),MC_T0,MC_T0},{MC_T(4,0,63,0),MC_T0,MC_T0},{MC_T(5,0,63,0),MC_T0,MC_T0},{MC_T(0,6,3,0),MC_T(1,6,3,2),MC_T(2,6,3,4)},{MC_T(3,6,3,0),MC_T(4,6,3,2),MC_T(5,6,3,4)}}, \
  /* 7 */ {{MC_T(0,0,255,0),MC_T0,MC_T0},{MC_T(1,0,255,0),MC_T0,MC_T0},{MC_T(2,0,255,0),MC_T0,MC_T0},{MC_T(3,0,255,0),MC_T0,MC_T0},{MC_T(4,0,255,0),MC_T0,MC_T0},{MC_T(5,0,255,0),MC_T0,MC_T0},{MC_T(6,0,255,0),MC_T0,MC_T0},{MC_T(7,0,255,0),MC_T0,MC_T0}}, \
  /* 8 */ {{MC_T(0,0,255,0),MC_T0,MC_T0},{MC_T(1,0,255,0),MC_T0,MC_T0},{MC_T(2,0,255,0),MC_T0,MC_T0},{MC_T(3,0,255,0),MC_T0,MC_T0},{MC_T(4,0,255,0),MC_T0,MC_T0},{MC_T(5,0,255,0),MC_T0,MC_T0},{MC_T(6,0,255,0),MC_T0,MC_T0},{MC_T(7,0,255,0),MC_T0,MC_T0}}, \
  /* 9 */ {{MC_T(0,0,255,0),MC_T(4,0,3,8),MC_T0},{MC_T(1,0,255,0),MC_T(4,2,3,8),MC_T0},{MC_T(2,0,255,0),MC_T(4,4,3,8),MC_T0},{MC_T(3,0,255,0),MC_T(4,6,3,8),MC_T0},{MC_T(5,0,255,0),MC_T(9,0,3,8),MC_T0},{MC_T(6,0,255,0),MC_T(9,2,3,8),MC_T0},{MC_T(7,0,255,0),MC_T(9,4,3,8),MC_T0},{MC_T(8,0,255,0),MC_T(9,6,3,8),MC_T0}}, \
  /*10 */ {{MC_T(0,0,255,0),MC_T(4,0,3,8),MC_T0},{MC_T(1,0,255,0),MC_T(4,2,3,8),MC_T0},{MC_T(2,0,255,0),MC_T(4,4,3,8),MC_T0},{MC_T(3,0,255,0),MC_T(4,6,3,8),MC_T0},{MC_T(5,0,255,0),MC_T(9,0,3,8),MC_T0},{MC_T(6,0,255,0),MC_T(9,2,3,8),MC_T0},{MC_T(7,0,255,0),MC_T(9,4,3,8),MC_T0},{MC_T(8,0,255,0),MC_T(9,6,3,8),MC_T0}} }

#define MCRAW_META_ROWS 11

#ifdef __CUDACC__
#define MC_HD __host__ __device__
#else
#define MC_HD
#endif

/* One term applied to the byte x found at G_g[b]. */
static inline MC_HD uint32_t mcraw_meta_term(uint32_t term, uint32_t x) {
    return ((x >> ((term >> 4) & 7u)) & (term >> 16)) << ((term >> 8) & 15u);
}
/* The same term applied to TWO adjacent byte lanes at once: w = G_g[b] | G_g[b + 1] << 8.  Result: the two contributions in
 * two 16-bit lanes (lane 0: byte lane b, lane 1: byte lane b + 1).  Every field of the table lies inside its byte
 * (bit length of m <= 8 - s), so the bits that (w >> s) drags from the high byte into the low one are masked away. */
static inline MC_HD uint32_t mcraw_meta_term_pair(uint32_t term, uint32_t w) {
    const uint32_t pair = (w >> ((term >> 4) & 7u)) & ((term >> 16) * 0x0101u);
    return ((pair & 0xFFu) | ((pair & 0xFF00u) << 8)) << ((term >> 8) & 15u);
}
/* Group index of a term (the byte to fetch is payload[8 * group + b]). */
static inline MC_HD uint32_t mcraw_meta_term_group(uint32_t term) { return term & 15u; }

/* Byte permute (PTX prmt, default mode): result byte i = byte (selector nibble i) of the 8 bytes {x, y}. */
static inline MC_HD uint32_t mcraw_prmt(uint32_t x, uint32_t y, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, y, s);
#else
    const uint64_t xy = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((xy >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}

/* Payload bytes / 8 of four blocks at once: v holds four header bits values (0..16, RawData.cpp:27-45), one per byte. */
static inline MC_HD uint32_t mcraw_len8x4(uint32_t v) {
    const uint32_t n = v & 0x07070707u;
    const uint32_t sel = mcraw_prmt(n | (n >> 4), 0u, 0x4420u);            /* nibble i = (value i) & 7 */
    const uint32_t lo = mcraw_prmt(0x03020100u, 0x08060504u, sel);          /* values 0..7  -> 0,1,2,3,4,5,6,8 */
    const uint32_t hi = mcraw_prmt(0x100A0A08u, 0x10101010u, sel);          /* values 8..15 -> 8,10,10,16,16,16,16,16 */
    const uint32_t m8 = ((v >> 3) & 0x01010101u) * 0xFFu;                   /* bytes whose value is 8..15 */
    return ((lo & ~m8) | (hi & m8)) + (v & 0x10101010u);                    /* value 16 -> 0 + 16 */
}

#endif /* MCRAW_META_TABLE_H */
