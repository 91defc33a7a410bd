/* Test helper (CPU): value i of a block through the table of csrc/mcraw_meta_table.h. */
#include "mcraw_meta_table.h"

static const uint32_t kTerms[MCRAW_META_ROWS][8][3] = MCRAW_META_TERMS_INIT;

unsigned meta_table_sample(const unsigned char* payload, int hb, int i) {
    const int j = i >> 3, b = i & 7;
    if (hb > 10) return (unsigned)payload[2 * i] | ((unsigned)payload[2 * i + 1] << 8);
    unsigned v = 0;
    for (int t = 0; t < 3; t++) {
        const uint32_t term = kTerms[hb][j][t];
        v |= mcraw_meta_term(term, payload[8 * mcraw_meta_term_group(term) + b]);
    }
    return v;
}

/* Samples i (even) and i + 1 through the two-lane form used on the device: value(i) | value(i + 1) << 16. */
unsigned meta_table_pair(const unsigned char* payload, int hb, int i) {
    const int j = i >> 3, b = i & 7;
    if (hb > 10) return ((unsigned)payload[2 * i] | ((unsigned)payload[2 * i + 1] << 8)) |
                        (((unsigned)payload[2 * i + 2] | ((unsigned)payload[2 * i + 3] << 8)) << 16);
    unsigned v = 0;
    for (int t = 0; t < 3; t++) {
        const uint32_t term = kTerms[hb][j][t];
        if (term >> 16) {
            const unsigned g = 8 * mcraw_meta_term_group(term) + b;
            v |= mcraw_meta_term_pair(term, (unsigned)payload[g] | ((unsigned)payload[g + 1] << 8));
        }
    }
    return v;
}

unsigned meta_len8x4(unsigned v) { return mcraw_len8x4(v); }
