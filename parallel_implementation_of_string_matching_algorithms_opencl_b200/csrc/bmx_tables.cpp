// bmx_tables.cpp -- per-pattern pre-processing (host, once per pattern).
//
// Replaces BoyreMoore/BoyreMoore/BoyreMoore.cpp:153-162 (bad-symbol table) and :165-190 with the
// helpers :16-60 (good-suffix table).  The reference derives the good-suffix shifts with nested
// rescans (O(m^3) worst case, m <= 99); here they come from the classical suffix-length array in
// O(m), which yields the same strong good-suffix values for every k = 1..m-1
// (tests/test_host_logic.py checks this against the reference's own code and against the oracle).
#include "bmx_internal.h"

#include <vector>

namespace bmx {

void build_bad_table(const unsigned char *pat, int32_t m, int32_t bad[256])
{
    // Every byte value shifts by m unless it occurs in P[0..m-2]; the last occurrence wins
    // (BoyreMoore.cpp:154-162, widened from 128 signed-char slots to 256 unsigned ones).
    for (int c = 0; c < 256; ++c) bad[c] = m;
    for (int32_t i = 0; i < m - 1; ++i) bad[pat[i]] = m - 1 - i;
}

void build_good_table(const unsigned char *pat, int32_t m, int32_t *good)
{
    if (m <= 0) return;
    good[0] = 0;  // never consulted: with k == 0 only the bad-symbol shift applies (kernel1.cl:30)
    if (m == 1) return;

    // suf[i] = length of the longest common suffix of P[0..i] and P.
    std::vector<int32_t> suf((size_t)m);
    suf[m - 1] = m;
    int32_t g = m - 1, f = m - 1;
    for (int32_t i = m - 2; i >= 0; --i) {
        if (i > g && suf[i + m - 1 - f] < i - g) {
            suf[i] = suf[i + m - 1 - f];
        } else {
            if (i < g) g = i;
            f = i;
            while (g >= 0 && pat[g] == pat[g + m - 1 - f]) --g;
            suf[i] = f - g;
        }
    }

    // shift_at[j] = shift when the mismatch is at pattern index j (suffix P[j+1..m) matched).
    std::vector<int32_t> shift_at((size_t)m, m);
    // Case 2 of the reference (:175-183): a border of P (prefix == suffix) shorter than the
    // matched suffix; longer borders (smaller shifts) are assigned first.
    int32_t j = 0;
    for (int32_t i = m - 1; i >= 0; --i) {
        if (suf[i] == i + 1) {
            for (; j < m - 1 - i; ++j)
                if (shift_at[j] == m) shift_at[j] = m - 1 - i;
        }
    }
    // Case 1 of the reference (:169-173): the matched suffix re-occurs further left, preceded
    // by a different byte (or by nothing); the right-most re-occurrence is visited last and so
    // leaves the smallest shift.
    for (int32_t i = 0; i <= m - 2; ++i) shift_at[m - 1 - suf[i]] = m - 1 - i;

    for (int32_t k = 1; k <= m - 1; ++k) good[k] = shift_at[m - 1 - k];
}

}  // namespace bmx
