// mcraw_tables.h -- the (header bits value, plane) -> extraction-recipe table of the current frame format.
//
// A 64-sample block is stored "byte-lane planar" (reference unpackers: /root/reference/lib/RawData.cpp:112-408):
// sample i = 8*j + l (plane j, byte lane l) is assembled from byte lane l of at most three 8-byte groups of the
// payload.  Because every group contributes to the SAME byte lane, four (or eight) samples of a plane can be
// extracted at once with 32-bit SWAR: for each contributing group   part = (word >> shift) & mask   and the
// parts are OR-ed.  The left shifts of the reference formulas are folded into a smaller net right shift plus a
// pre-shifted mask, e.g. ((G2 >> 6) & 1) << 2  ==  (G2 >> 4) & 0x04.   Bits 8..9 of the 10-bit layout go to a
// separate "high byte" word.  Rows for 11..16 mark the 16-bit little-endian layout (a byte de-interleave).
//
// One entry = 8 x u32:
//   [0] offA | offB<<8 | offC<<16 | flags<<24     byte offsets of the three source groups inside the block payload
//   [1] shA  | shB<<8  | shC<<16                  net right shifts (0..7)
//   [2] maskA  (low-byte part, replicated to 4 byte lanes)
//   [3] maskB  -> low byte      [4] maskB -> high byte (10-bit layout only)
//   [5] maskC  -> low byte
//   [6] payload length of the block in bytes (RawData.cpp:27-45)      [7] 0
// A zero mask means "term unused" (its load is skipped).  flags bit0 = 16-bit layout: A/B load bytes
// [16j,16j+8) and [16j+8,16j+16) and are de-interleaved instead of masked.
#pragma once
#include <stdint.h>

#define MCRAW_TAB_ENTRIES (17 * 8)
#define MCRAW_TAB_WORDS 8

static inline uint32_t mcraw_rep4(uint32_t m) { return m * 0x01010101u; }

static inline void mcraw_build_table(uint32_t* tab /* [17*8][8] */) {
    static const uint32_t len[17] = {0, 8, 16, 24, 32, 40, 48, 64, 64, 80, 80, 128, 128, 128, 128, 128, 128};
    for (int b = 0; b <= 16; b++) {
        for (int j = 0; j < 8; j++) {
            uint32_t gA = 0, sA = 0, mA = 0;
            uint32_t gB = 0, sB = 0, mBL = 0, mBH = 0;
            uint32_t gC = 0, sC = 0, mC = 0;
            uint32_t flags = 0, offA, offB, offC;
            switch (b) {
            case 0: break;
            case 1: gA = 0; sA = j; mA = 1; break;
            case 2: gA = j >> 2; sA = 2 * (j & 3); mA = 3; break;
            case 3:
                switch (j) {
                case 0: gA = 0; sA = 0; mA = 7; break;
                case 1: gA = 0; sA = 3; mA = 7; break;
                case 2: gA = 0; sA = 6; mA = 3; gB = 2; sB = 4; mBL = 0x04; break;
                case 3: gA = 1; sA = 0; mA = 7; break;
                case 4: gA = 1; sA = 3; mA = 7; break;
                case 5: gA = 1; sA = 6; mA = 3; gB = 2; sB = 5; mBL = 0x04; break;
                case 6: gA = 2; sA = 0; mA = 7; break;
                default: gA = 2; sA = 3; mA = 7; break;
                }
                break;
            case 4: gA = j >> 1; sA = 4 * (j & 1); mA = 15; break;
            case 5:
                if (j <= 4) { gA = j; sA = 0; mA = 31; }
                else if (j == 5) { gA = 0; sA = 5; mA = 7; gB = 3; sB = 2; mBL = 0x18; }
                else if (j == 6) { gA = 1; sA = 5; mA = 7; gB = 4; sB = 2; mBL = 0x18; }
                else { gA = 2; sA = 5; mA = 7; gB = 3; sB = 4; mBL = 0x08; gC = 4; sC = 3; mC = 0x10; }
                break;
            case 6:
                if (j <= 5) { gA = j; sA = 0; mA = 63; }
                else if (j == 6) { gA = 0; sA = 6; mA = 3; gB = 1; sB = 4; mBL = 0x0C; gC = 2; sC = 2; mC = 0x30; }
                else { gA = 3; sA = 6; mA = 3; gB = 4; sB = 4; mBL = 0x0C; gC = 5; sC = 2; mC = 0x30; }
                break;
            case 7:
            case 8: gA = j; sA = 0; mA = 0xFF; break;
            case 9:
            case 10:
                if (j < 4) { gA = j; mA = 0xFF; gB = 4; sB = 2 * j; mBH = 3; }
                else { gA = j + 1; mA = 0xFF; gB = 9; sB = 2 * (j - 4); mBH = 3; }
                break;
            default: flags = 1; mA = 0xFF; mBL = 0xFF; break;
            }
            if (flags & 1) { offA = 16 * j; offB = 16 * j + 8; offC = 0; }
            else { offA = 8 * gA; offB = 8 * gB; offC = 8 * gC; }
            uint32_t* e = tab + (b * 8 + j) * MCRAW_TAB_WORDS;
            e[0] = offA | (offB << 8) | (offC << 16) | (flags << 24);
            e[1] = sA | (sB << 8) | (sC << 16);
            e[2] = mcraw_rep4(mA);
            e[3] = mcraw_rep4(mBL);
            e[4] = mcraw_rep4(mBH);
            e[5] = mcraw_rep4(mC);
            e[6] = len[b];
            e[7] = 0;
        }
    }
}
