// bmx_partition.cpp -- the reference's host-side word partitioner, kept for the mirror entry point
// bmx_search_partitions and the bmx_refmain demo.
//
// Replaces BoyreMoore/BoyreMoore/BoyreMoore.cpp:94-141.  The reference counts spaces (skipping
// offset 0, :94-98), measures one word per separator (:107-117) and hands each of the P
// work-items the same number of words as one inclusive byte range; the space between two ranges
// belongs to neither and left-over words are dropped (:119-141).  The new scan itself does not
// partition on words (tiles + halo instead); this exists so the reference's
// per-process counts can be reproduced through the same numbers it would have used.
#include <vector>

#include "bmx_internal.h"

extern "C" int bmx_partition_words(const char *text, int64_t n, int32_t nparts, int32_t *se)
{
    if (!se || nparts <= 0 || n < 0 || (!text && n > 0))
        return bmx::fail(BMX_E_BADARG, "bmx_partition_words: bad argument");
    if (n > INT32_MAX) return bmx::fail(BMX_E_BADARG, "bmx_partition_words: the reference's ranges are 32-bit");

    // The reference works on a NUL-terminated copy (:89-90): stop at the first NUL.
    int64_t len = 0;
    while (len < n && text[len] != '\0') ++len;

    // Token boundaries of split(' '): token t spans [tok_begin[t], tok_end[t]).
    std::vector<int32_t> tok_begin, tok_end;
    int32_t begin = 0;
    for (int64_t i = 0; i <= len; ++i) {
        if (i == len || text[i] == ' ') {
            tok_begin.push_back(begin);
            tok_end.push_back((int32_t)i);
            begin = (int32_t)i + 1;
        }
    }
    // Words as the reference counts them: one more than the spaces at offsets >= 1, so a leading
    // space costs the text its last token.
    int64_t words = (int64_t)tok_begin.size();
    if (len > 0 && text[0] == ' ') words -= 1;

    const int64_t per = words / nparts;
    int32_t next_start = 0;
    for (int32_t p = 0; p < nparts; ++p) {
        int32_t chars = 0;
        for (int64_t t = (int64_t)p * per; t < (int64_t)(p + 1) * per; ++t) chars += tok_end[(size_t)t] - tok_begin[(size_t)t];
        const int32_t stop = next_start + chars + (int32_t)per - 1;  // one past the range's last byte
        se[2 * p] = next_start;
        se[2 * p + 1] = stop - 1;
        next_start = stop + 1;
    }
    return BMX_OK;
}
