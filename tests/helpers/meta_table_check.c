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
