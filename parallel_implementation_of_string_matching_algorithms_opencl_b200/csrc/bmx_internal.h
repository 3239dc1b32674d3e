// bmx_internal.h -- declarations shared by the translation units of libbmx.so (not installed).
#pragma once

#include <cstddef>
#include <cstdint>

#include "../../include/bmx.h"

namespace bmx {

// ---- bmx_tables.cpp -----------------------------------------------------------------------
void build_bad_table(const unsigned char *pat, int32_t m, int32_t bad[256]);
void build_good_table(const unsigned char *pat, int32_t m, int32_t *good /* m ints */);

// ---- error plumbing (bmx_abi.cu) ----------------------------------------------------------
int fail(int code, const char *fmt, ...);

// ---- geometry of the scan kernels (bmx_scan.cu) -------------------------------------------
constexpr int kConsumerWarps = 8;                       // warps that filter/verify/emit
constexpr int kConsumerThreads = kConsumerWarps * 32;   // 256
constexpr int kThreads = kConsumerThreads + 32;         // + one producer warp (TMA issue)
constexpr int kMaxStages = 8;
constexpr int kPre = 16;            // bytes staged in front of a tile (q-gram owner offset -3)
constexpr int kPatSmemMax = 1024;   // patterns up to this length keep pattern+tables in smem
constexpr int kHaloSmemMax = 4096;  // longer patterns verify their tail from global memory

// Tile-status word of the decoupled look-back: [63:62] state, [61:0] value.
constexpr unsigned long long kStateAgg = 1ull << 62;    // value = hits inside this tile
constexpr unsigned long long kStateIncl = 2ull << 62;   // value = hits in tiles 0..this
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// Kernel arguments (passed by value: they live in the constant bank of the launch).
struct ScanArgs {
    const uint8_t *vtext;   // 16-byte aligned base A <= text; "V space" offset v addresses A[v]
    int64_t vlen;           // lead + n: bytes of V that may be read
    int64_t vmin;           // first valid start position in V (= lead)
    int64_t vmax;           // last valid start position in V (= lead + n - m)
    int64_t pos_bias;       // reported position = v + pos_bias (= pos_base - lead)
    int32_t m;
    uint32_t num_tiles;
    uint32_t stages;        // pipeline depth
    uint32_t stage_stride;  // bytes between stages in shared memory
    uint32_t halo;          // bytes staged behind a tile
    uint32_t verify_smem;   // 1: candidates are verified from the staged tile, 0: from global
    uint32_t pat_smem;      // 1: pattern + tables copied to shared memory
    // filter constants
    uint32_t f[4];          // QGRAM: hash of P[r..r+q) for r = 0..3; WINDOW: f[0] = target
    uint32_t mask2;         // QGRAM: mask of the second word (q-4 bytes)
    uint32_t mulc;          // WINDOW: 2^(32-8q), drops the bytes beyond q
    // per-pattern block in global memory
    const uint8_t *g_pat;
    const int32_t *g_bad;
    const int32_t *g_good;
    // output
    int64_t *pos_out;
    int64_t pos_cap;
    unsigned long long *tile_state;        // num_tiles words, zeroed before the launch
    uint32_t *tile_counter;                // zeroed before the launch
    const unsigned long long *carry_in;    // hits reported by earlier chained scans
    unsigned long long *carry_out;         // carry_in + hits of this scan (written by last tile)
    unsigned long long *count_acc;         // count-only mode: atomically accumulated
};

// Launch description produced by plan_scan() and consumed by launch_scan().
struct ScanLaunch {
    int variant;        // resolved bmx_variant
    int tile_bytes;
    int grid;
    size_t smem_bytes;
};

// Host entry points implemented in bmx_scan.cu.
int resolve_variant(int requested, int32_t m);
int plan_scan(int device, int variant, int32_t m, bool positions, ScanArgs *args, ScanLaunch *out);
int launch_scan(const ScanArgs &args, const ScanLaunch &launch, bool positions, void *stream);
int launch_synth_fill(void *d_text, int64_t offset, int64_t len, uint64_t seed,
                      const unsigned char *alphabet, int32_t sigma, void *stream);
int launch_partition_count(const int64_t *d_pos, const unsigned long long *d_count, int64_t pos_cap,
                           const int32_t *d_se, int32_t *d_ans, int32_t m, int32_t nparts, void *stream);

// Filter constants for a pattern (host).
void fill_filter_constants(int variant, const unsigned char *pat, int32_t m, ScanArgs *args);

constexpr uint32_t kHashMul = 0x9E3779B1u;

}  // namespace bmx
