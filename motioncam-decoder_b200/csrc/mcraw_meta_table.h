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
/* Group index of a term (the byte to fetch is payload[8 * group + b]). */
static inline MC_HD uint32_t mcraw_meta_term_group(uint32_t term) { return term & 15u; }

#endif /* MCRAW_META_TABLE_H */
