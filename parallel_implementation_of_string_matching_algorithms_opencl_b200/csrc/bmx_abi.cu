// bmx_abi.cu -- the C ABI of libbmx.so (include/bmx.h): host glue over the scan kernels.
//
// Replaces the reference's OpenCL host layer, BoyreMoore/BoyreMoore/BoyreMoore.cpp:192-313
// (context/queue :213-231, buffers :233-244, blocking writes :246-252, kernel + arguments
// :261-270, NDRange launch :273-280, blocking read of the counts :286, releases :299-312).
// There is no CPU fallback anywhere in this file: every scanning entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "bmx_internal.h"

namespace bmx {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char tl_error[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof tl_error, fmt, ap);
    va_end(ap);
    return code;
}

#define BMX_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "%s: %s", #call,     \
                        cudaGetErrorString(e_));                                                         \
    } while (0)

static int check_device(int device)
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return fail(BMX_E_NODEVICE, "no CUDA device visible (%s); libbmx has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) return fail(BMX_E_BADARG, "device %d out of range (0..%d)", device, n - 1);
    return BMX_OK;
}

}  // namespace bmx

using namespace bmx;

// ---------------------------------------------------------------------------------------------
// scanner
// ---------------------------------------------------------------------------------------------
struct bmx_scanner {
    int device = 0;
    // per-pattern state
    int32_t m = 0;
    int variant = 0;
    std::vector<unsigned char> pat;
    void *d_block = nullptr;  // [bad 256 x i32][good m x i32][pattern m bytes]
    size_t d_block_cap = 0;
    ScanArgs proto{};         // filter constants + pattern pointers
    // result state
    unsigned long long *d_ctrl = nullptr;  // [0],[1] carry ping-pong, [2] count-only running total, [3] always 0
    unsigned long long *h_result = nullptr;  // pinned
    void *d_scratch = nullptr;  // ticket, block sums/bases, segment counts, hit masks (see bmx_scanner_scan)
    size_t d_scratch_cap = 0;
    size_t zero_cap = 0;        // bytes of each of the two zero-initialised halves at the front of d_scratch
    size_t dirty[2] = {0, 0};   // leading bytes of each half that enqueued work leaves non-zero
    int cur_half = 0;           // half the next scan uses
    int64_t *d_pos_out = nullptr;
    int64_t pos_cap = 0;
    bool positions = false;
    uint32_t scan_index = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around the scan kernel alone
    bool timing_open = false;
    int timing_level = 2;  // 0 none, 1 whole scan, 2 + scan kernel alone
    bmx_stats stats{};
};

// Device word holding the search's hit count: the carry slot the last scan wrote (positions mode), the
// running total (count-only mode), or the constant zero while no scan has been launched since begin().
static const unsigned long long *result_slot(const bmx_scanner *s)
{
    if (s->scan_index == 0) return s->d_ctrl + 3;
    return s->positions ? s->d_ctrl + (s->scan_index & 1u) : s->d_ctrl + 2;
}

extern "C" {

int bmx_version(void) { return BMX_VERSION; }

const char *bmx_last_error(void) { return tl_error; }

int bmx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int bmx_build_tables(const char *pat, int32_t m, int32_t bad[256], int32_t *good)
{
    if (!pat || m <= 0 || m > BMX_MAX_PATTERN || !bad || !good)
        return fail(BMX_E_BADARG, "bmx_build_tables: pat/bad/good must be non-NULL and 1 <= m <= %d", BMX_MAX_PATTERN);
    build_bad_table(reinterpret_cast<const unsigned char *>(pat), m, bad);
    build_good_table(reinterpret_cast<const unsigned char *>(pat), m, good);
    return BMX_OK;
}

int bmx_scanner_create(int device, bmx_scanner **out)
{
    if (!out) return fail(BMX_E_BADARG, "bmx_scanner_create: out is NULL");
    *out = nullptr;
    if (int rc = check_device(device)) return rc;
    BMX_CUDA(cudaSetDevice(device));
    bmx_scanner *s = new (std::nothrow) bmx_scanner();
    if (!s) return fail(BMX_E_NOMEM, "out of host memory");
    s->device = device;
    cudaError_t e = cudaMalloc(&s->d_ctrl, 64);
    if (e == cudaSuccess) e = cudaMemset(s->d_ctrl, 0, 64);
    if (e == cudaSuccess) e = cudaHostAlloc(&s->h_result, 64, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_stop);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_k0);
    if (e == cudaSuccess) e = cudaEventCreate(&s->ev_k1);
    if (e != cudaSuccess) {
        bmx_scanner_destroy(s);
        return fail(e == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "bmx_scanner_create: %s",
                    cudaGetErrorString(e));
    }
    *out = s;
    return BMX_OK;
}

void bmx_scanner_destroy(bmx_scanner *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->d_block) cudaFree(s->d_block);
    if (s->d_scratch) cudaFree(s->d_scratch);
    if (s->d_ctrl) cudaFree(s->d_ctrl);
    if (s->h_result) cudaFreeHost(s->h_result);
    if (s->ev_start) cudaEventDestroy(s->ev_start);
    if (s->ev_stop) cudaEventDestroy(s->ev_stop);
    if (s->ev_k0) cudaEventDestroy(s->ev_k0);
    if (s->ev_k1) cudaEventDestroy(s->ev_k1);
    delete s;
}

int bmx_scanner_set_pattern(bmx_scanner *s, const char *pat, int32_t m, int32_t variant, void *stream)
{
    if (!s || !pat) return fail(BMX_E_BADARG, "bmx_scanner_set_pattern: NULL argument");
    if (m <= 0 || m > BMX_MAX_PATTERN)
        return fail(BMX_E_BADARG, "pattern length %d outside 1..%d (an empty pattern is rejected)", m, BMX_MAX_PATTERN);
    if (variant < BMX_VARIANT_AUTO || variant > BMX_VARIANT_SHIFTAND)
        return fail(BMX_E_BADARG, "unknown variant %d", variant);
    BMX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    const unsigned char *p = reinterpret_cast<const unsigned char *>(pat);
    s->m = m;
    s->variant = resolve_variant(variant, m);
    s->pat.assign(p, p + m);

    // Host image of the device block: tables once per pattern (BoyreMoore.cpp:153-190).
    const size_t good_off = 256 * sizeof(int32_t);
    const size_t pat_off = good_off + (size_t)m * sizeof(int32_t);
    const size_t bytes = ((pat_off + (size_t)m + 15) & ~size_t(15)) + 16;  // verification reads aligned word pairs
    std::vector<unsigned char> image(bytes, 0);
    build_bad_table(p, m, reinterpret_cast<int32_t *>(image.data()));
    build_good_table(p, m, reinterpret_cast<int32_t *>(image.data() + good_off));
    memcpy(image.data() + pat_off, p, (size_t)m);

    if (bytes > s->d_block_cap) {
        // the old block may still be read by scans in flight on `stream`
        BMX_CUDA(cudaStreamSynchronize(st));
        if (s->d_block) cudaFree(s->d_block);
        s->d_block = nullptr;
        s->d_block_cap = 0;
        BMX_CUDA(cudaMalloc(&s->d_block, bytes));
        s->d_block_cap = bytes;
    }
    // pageable source: the runtime stages it before returning, so `image` may die afterwards
    BMX_CUDA(cudaMemcpyAsync(s->d_block, image.data(), bytes, cudaMemcpyHostToDevice, st));

    s->proto = ScanArgs{};
    s->proto.m = m;
    s->proto.g_bad = reinterpret_cast<const int32_t *>(s->d_block);
    s->proto.g_good = reinterpret_cast<const int32_t *>(static_cast<unsigned char *>(s->d_block) + good_off);
    s->proto.g_pat = static_cast<const uint8_t *>(s->d_block) + pat_off;
    fill_filter_constants(s->variant, p, m, &s->proto);
    return BMX_OK;
}

int bmx_scanner_begin(bmx_scanner *s, int64_t *d_pos_out, int64_t pos_cap, void *stream)
{
    if (!s) return fail(BMX_E_BADARG, "bmx_scanner_begin: NULL scanner");
    if (pos_cap < 0) return fail(BMX_E_BADARG, "pos_cap < 0");
    BMX_CUDA(cudaSetDevice(s->device));
    s->d_pos_out = d_pos_out;
    s->pos_cap = d_pos_out ? pos_cap : 0;
    s->positions = d_pos_out != nullptr;
    s->scan_index = 0;
    s->timing_open = false;
    s->stats = bmx_stats{};
    s->stats.variant = s->variant;
    // no device work here: the first scan of the search treats the carried state as zero (ScanArgs::first_scan)
    (void)stream;
    return BMX_OK;
}

int bmx_scanner_scan(bmx_scanner *s, const void *d_text, int64_t n, int64_t pos_base, void *stream)
{
    if (!s || s->m <= 0) return fail(BMX_E_BADARG, "bmx_scanner_scan: scanner has no pattern");
    if (n < 0 || (!d_text && n > 0)) return fail(BMX_E_BADARG, "bmx_scanner_scan: bad text (n=%lld)", (long long)n);
    if (n < s->m) return BMX_OK;  // kernel1.cl:15,19: the loop is never entered
    BMX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    ScanArgs a = s->proto;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(d_text);
    const int64_t lead = (int64_t)(addr & 15u);
    a.vtext = reinterpret_cast<const uint8_t *>(addr - (uintptr_t)lead);
    a.vlen = lead + n;
    a.vmin = lead;
    a.vmax = lead + n - s->m;
    a.pos_bias = pos_base - lead;
    a.pos_out = s->d_pos_out;
    a.pos_cap = s->pos_cap;

    ScanLaunch launch{};
    if (int rc = plan_scan(s->device, s->variant, s->m, s->positions, &a, &launch)) return rc;

    // scratch: two "zero halves" [tickets 16 B | scan count u64 + pad | block_sum u32 x blocks | seg_count u16 x segs |
    //          item_flag u8 x items], each zero whenever a scan starts on it, followed by
    //          [block_base u64 x blocks | dense_list | mask16 u16 x chunks]                <- written before read.
    // Scans alternate between the halves: the expand kernel of scan i clears the half scan i-1 used, so in
    // steady state no memset sits between the kernels; a count-only scan resets its 32-byte header itself.
    const size_t off_bsum = 32;
    const size_t off_segc = off_bsum + (((size_t)a.num_blocks * 4 + 15) & ~size_t(15));
    // seg_count is padded to whole blocks: the expand kernel reads a block's 1024 counts with vector loads
    const size_t off_flag = off_segc + (size_t)a.num_blocks * kBlockSegs * 2;
    const size_t zero_bytes = s->positions ? off_flag + (((size_t)a.num_blocks * kExpandSplit + 15) & ~size_t(15)) : off_bsum;
    const size_t off_bbase = 0;
    const size_t off_dense = off_bbase + (((size_t)a.num_blocks * 8 + 15) & ~size_t(15));  // mask16 is read with 16-byte loads
    const size_t off_mask = off_dense + (((size_t)a.num_blocks * 4 + 15) & ~size_t(15));
    const size_t rest_bytes = s->positions ? off_mask + (size_t)a.num_segs * kSegChunks * 2 : 0;
    if (zero_bytes > s->zero_cap || 2 * s->zero_cap + rest_bytes > s->d_scratch_cap) {
        BMX_CUDA(cudaStreamSynchronize(st));
        if (s->d_scratch) cudaFree(s->d_scratch);
        s->d_scratch = nullptr;
        s->d_scratch_cap = 0;
        const size_t zcap = std::max<size_t>(s->zero_cap, ((zero_bytes + zero_bytes / 8 + 255) & ~size_t(255)));
        const size_t want = std::max<size_t>(2 * zcap + rest_bytes + rest_bytes / 8, 1 << 20);
        BMX_CUDA(cudaMalloc(&s->d_scratch, want));
        s->d_scratch_cap = want;
        s->zero_cap = zcap;
        BMX_CUDA(cudaMemsetAsync(s->d_scratch, 0, 2 * zcap, st));
        s->dirty[0] = s->dirty[1] = 0;
        s->cur_half = 0;
    }
    const int half = s->cur_half;
    unsigned char *base = static_cast<unsigned char *>(s->d_scratch) + (size_t)half * s->zero_cap;
    unsigned char *rest = static_cast<unsigned char *>(s->d_scratch) + 2 * s->zero_cap;
    if (s->dirty[half]) {  // not reached by alternating scans; keeps any other call order correct
        BMX_CUDA(cudaMemsetAsync(base, 0, s->dirty[half], st));
        s->dirty[half] = 0;
    }
    a.tile_counter = reinterpret_cast<uint32_t *>(base);
    a.scan_count = reinterpret_cast<unsigned long long *>(base + 16);
    a.block_sum = reinterpret_cast<uint32_t *>(base + off_bsum);
    a.seg_count = reinterpret_cast<uint16_t *>(base + off_segc);
    a.item_flag = base + off_flag;
    a.block_base = reinterpret_cast<unsigned long long *>(rest + off_bbase);
    a.dense_list = reinterpret_cast<uint32_t *>(rest + off_dense);
    a.mask16 = reinterpret_cast<uint16_t *>(rest + off_mask);
    a.zero_ptr = static_cast<unsigned char *>(s->d_scratch) + (size_t)(1 - half) * s->zero_cap;
    a.zero_vec16 = s->positions ? (uint32_t)((s->dirty[1 - half] + 15) / 16) : 0u;
    a.carry_in = s->d_ctrl + (s->scan_index & 1u);
    a.carry_out = s->d_ctrl + ((s->scan_index + 1u) & 1u);
    a.count_acc = s->d_ctrl + 2;
    a.first_scan = s->scan_index == 0 ? 1u : 0u;

    if (s->timing_level >= 1 && !s->timing_open) {
        BMX_CUDA(cudaEventRecord(s->ev_start, st));
        s->timing_open = true;
    }
    if (s->timing_level >= 2) BMX_CUDA(cudaEventRecord(s->ev_k0, st));
    if (int rc = launch_scan(a, launch, s->positions, st)) return rc;
    if (s->timing_level >= 2) BMX_CUDA(cudaEventRecord(s->ev_k1, st));
    if (s->positions) {
        s->dirty[half] = zero_bytes;
        if (int rc = launch_emit(a, st)) return rc;
        s->dirty[1 - half] = 0;   // cleared by the expand kernel just enqueued
        s->cur_half = 1 - half;
    }
    if (s->timing_level >= 1) BMX_CUDA(cudaEventRecord(s->ev_stop, st));

    s->scan_index += 1;
    s->stats.kernel_launches += s->positions ? 2 : 1;
    s->stats.grid = launch.grid;
    s->stats.stages = (int32_t)a.stages;
    s->stats.tile_bytes = launch.tile_bytes;
    s->stats.smem_bytes = (int32_t)launch.smem_bytes;
    s->stats.tiles += a.num_tiles;
    return BMX_OK;
}

int bmx_scanner_set_timing(bmx_scanner *s, int level)
{
    if (!s || level < 0 || level > 2) return fail(BMX_E_BADARG, "bmx_scanner_set_timing: level must be 0, 1 or 2");
    s->timing_level = level;
    return BMX_OK;
}

int bmx_scanner_export_result(bmx_scanner *s, void *d_dst, int64_t head, void *stream)
{
    if (!s || !d_dst || head < 0) return fail(BMX_E_BADARG, "bmx_scanner_export_result: bad argument");
    BMX_CUDA(cudaSetDevice(s->device));
    const unsigned long long *src = result_slot(s);
    return launch_export_result(src, s->d_pos_out, s->pos_cap, d_dst, s->positions ? head : 0, stream);
}

int bmx_scanner_finish(bmx_scanner *s, uint64_t *count_out, bmx_stats *stats, void *stream)
{
    if (!s || !count_out) return fail(BMX_E_BADARG, "bmx_scanner_finish: NULL argument");
    BMX_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned long long *src = result_slot(s);
    BMX_CUDA(cudaMemcpyAsync(s->h_result, src, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    BMX_CUDA(cudaStreamSynchronize(st));
    *count_out = (uint64_t)s->h_result[0];
    if (s->timing_open) {
        float ms = 0.f;
        BMX_CUDA(cudaEventElapsedTime(&ms, s->ev_start, s->ev_stop));
        s->stats.device_ms = ms;
        if (s->timing_level >= 2 && s->stats.kernel_launches > 0 && cudaEventElapsedTime(&ms, s->ev_k0, s->ev_k1) == cudaSuccess)
            s->stats.scan_kernel_ms = ms;
    }
    if (stats) *stats = s->stats;
    return BMX_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// one cached scanner + streams per (host thread, device) for the convenience entry points
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxDevices = 64;
constexpr int kBounce = 3;

struct ThreadCtx {
    bmx_scanner *scanner = nullptr;
    cudaStream_t copy_stream = nullptr, scan_stream = nullptr;
    std::vector<cudaEvent_t> events;
    unsigned char *bounce[kBounce] = {nullptr, nullptr, nullptr};
    size_t bounce_bytes = 0;
    bool pool_tuned = false;
};
thread_local ThreadCtx tl_ctx[kMaxDevices];

int get_ctx(int device, ThreadCtx **out)
{
    if (int rc = check_device(device)) return rc;
    if (device >= kMaxDevices) return fail(BMX_E_BADARG, "device index %d too large", device);
    ThreadCtx &c = tl_ctx[device];
    BMX_CUDA(cudaSetDevice(device));
    if (!c.scanner) {
        if (int rc = bmx_scanner_create(device, &c.scanner)) return rc;
    }
    *out = &c;
    return BMX_OK;
}

int ensure_streams(ThreadCtx &c, int device, size_t nevents)
{
    if (!c.copy_stream) BMX_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    if (!c.scan_stream) BMX_CUDA(cudaStreamCreateWithFlags(&c.scan_stream, cudaStreamNonBlocking));
    while (c.events.size() < nevents) {
        cudaEvent_t e;
        BMX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.events.push_back(e);
    }
    if (!c.pool_tuned) {
        // keep freed blocks in the stream-ordered pool: the multi-GiB text buffer is reused
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        (void)cudaGetLastError();
        c.pool_tuned = true;
    }
    return BMX_OK;
}

// memcpy split over a few host threads: one core moves ~10 GB/s, a PCIe Gen5 x16 link takes ~55 GB/s
void parallel_copy(void *dst, const void *src, size_t bytes, int threads)
{
    if (threads <= 1 || bytes < (size_t(8) << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t slice = ((bytes / (size_t)threads) + 4095) & ~size_t(4095);
    std::vector<std::thread> helpers;
    for (int t = 1; t < threads; ++t) {
        const size_t lo = std::min(bytes, slice * (size_t)t), hi = std::min(bytes, lo + slice);
        if (hi > lo)
            helpers.emplace_back([=] { memcpy(static_cast<char *>(dst) + lo, static_cast<const char *>(src) + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, slice));
    for (auto &h : helpers) h.join();
}

// First chunk of a find-first search (doubles up to 1 GiB); BMX_FIND_CHUNK_KB is a test knob.
int64_t find_first_chunk_bytes()
{
    if (const char *e = getenv("BMX_FIND_CHUNK_KB")) {
        const long kb = atol(e);
        if (kb > 0) return (int64_t)kb << 10;
    }
    return (int64_t)16 << 20;
}

bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

}  // namespace

extern "C" {

int bmx_search_device_ex(const void *d_text, int64_t n, const char *pat, int32_t m, int64_t pos_base,
                         int64_t *d_pos_out, int64_t pos_cap, uint64_t *count_out, int32_t variant,
                         bmx_stats *stats, void *stream)
{
    if (!count_out || !pat) return fail(BMX_E_BADARG, "bmx_search_device: pat/count_out must be non-NULL");
    if (n < 0 || (!d_text && n > 0)) return fail(BMX_E_BADARG, "bmx_search_device: bad text (n=%lld)", (long long)n);
    if (pos_cap < 0) return fail(BMX_E_BADARG, "pos_cap < 0");
    int device = 0;
    if (int rc = check_device(0)) return rc;
    BMX_CUDA(cudaGetDevice(&device));
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    if (int rc = bmx_scanner_set_pattern(c->scanner, pat, m, variant, stream)) return rc;
    if (int rc = bmx_scanner_begin(c->scanner, d_pos_out, pos_cap, stream)) return rc;
    if (int rc = bmx_scanner_scan(c->scanner, d_text, n, pos_base, stream)) return rc;
    return bmx_scanner_finish(c->scanner, count_out, stats, stream);
}

int bmx_search_device(const void *d_text, int64_t n, const char *pat, int32_t m, int64_t *d_pos_out,
                      int64_t pos_cap, uint64_t *count_out, float *device_ms, void *stream)
{
    bmx_stats st{};
    const int rc = bmx_search_device_ex(d_text, n, pat, m, 0, d_pos_out, pos_cap, count_out, BMX_VARIANT_AUTO, &st, stream);
    if (rc == BMX_OK && device_ms) *device_ms = st.device_ms;
    return rc;
}


// First occurrence with early exit.  The text is scanned as chained scans of growing chunks (16 MiB doubling
// to 1 GiB; starts [s_k, e_k) need the bytes [s_k, e_k + m - 1)) with room for ONE position: chunks run in
// text order and a truncated list keeps the smallest positions, so that slot ends up holding the first
// occurrence.  After every chunk {count, first} is exported and copied to pinned memory; the host reads the
// result of chunk k-1 while chunk k is already running, so the GPU never waits for the host and at most one
// chunk is scanned in vain.
int bmx_find_first_device(const void *d_text, int64_t n, const char *pat, int32_t m, int64_t *first_out, void *stream)
{
    if (!first_out || !pat) return fail(BMX_E_BADARG, "bmx_find_first_device: pat/first_out must be non-NULL");
    if (n < 0 || (!d_text && n > 0)) return fail(BMX_E_BADARG, "bmx_find_first_device: bad text (n=%lld)", (long long)n);
    *first_out = -1;
    int device = 0;
    if (int rc = check_device(0)) return rc;
    BMX_CUDA(cudaGetDevice(&device));
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bmx_scanner *s = c->scanner;
    if (int rc = bmx_scanner_set_pattern(s, pat, m, BMX_VARIANT_AUTO, stream)) return rc;
    if (n < m) return BMX_OK;
    if (int rc = ensure_streams(*c, device, 2)) return rc;
    int64_t *d_slot = reinterpret_cast<int64_t *>(s->d_ctrl + 4);  // [4] first position, [5..7] exported {count, held, first}
    if (int rc = bmx_scanner_begin(s, d_slot, 1, stream)) return rc;

    int64_t chunk = find_first_chunk_bytes();
    const unsigned char *text = static_cast<const unsigned char *>(d_text);
    const int64_t last_start = n - m;
    int64_t k = 0;
    bool found = false;
    auto result_of = [&](int64_t kk) -> bool {   // waits for chunk kk's exported result
        if (cudaEventSynchronize(c->events[(size_t)(kk & 1)]) != cudaSuccess) return false;
        const unsigned long long *h = s->h_result + 3 * (kk & 1);
        if (h[0] == 0) return false;
        *first_out = (int64_t)h[2];
        return true;
    };
    for (int64_t s0 = 0; s0 <= last_start; ++k) {
        const int64_t e0 = std::min(last_start + 1, s0 + chunk);
        if (int rc = bmx_scanner_scan(s, text + s0, e0 - s0 + m - 1, s0, stream)) return rc;
        if (int rc = bmx_scanner_export_result(s, d_slot + 1, 1, stream)) return rc;
        BMX_CUDA(cudaMemcpyAsync(s->h_result + 3 * (k & 1), d_slot + 1, 24, cudaMemcpyDeviceToHost, st));
        BMX_CUDA(cudaEventRecord(c->events[(size_t)(k & 1)], st));
        if (k >= 1 && result_of(k - 1)) {
            found = true;
            break;
        }
        s0 = e0;
        chunk = std::min<int64_t>(chunk * 2, (int64_t)1 << 30);
    }
    BMX_CUDA(cudaStreamSynchronize(st));  // the scanner and its scratch are idle again when this returns
    if (!found && k >= 1) result_of(k - 1);
    return BMX_OK;
}

}  // extern "C"

namespace {

// Host text -> device (chunked, overlapped with scanning) -> hits in a device buffer.  On success *d_pos_out
// (when want_pos) is a stream-ordered allocation on c.scan_stream holding min(count, dev_cap) global
// positions (start + pos_base); the caller copies it out and frees it with cudaFreeAsync(.., c.scan_stream).
// With keep_text the device copy of the text (a stream-ordered allocation on c.scan_stream, n + 16 bytes) is
// handed to the caller instead of being freed, and the text is ingested even when it is shorter than the pattern.
// With first_out (find-first mode: want_pos, dev_cap = 1) the result of every chunk is read back one chunk
// behind the scans, and both the copies and the scans stop after the first chunk that holds a match;
// *first_out is that match's position (or stays -1) and *count_out is then only the count so far.
int ingest_and_scan(ThreadCtx &ctx, int device, const char *text, int64_t n, const char *pat, int32_t m, int64_t pos_base,
                    bool want_pos, int64_t dev_cap, int32_t variant, int64_t **d_pos_out, uint64_t *count_out, bmx_stats *stats,
                    int64_t *first_out = nullptr, unsigned char **keep_text = nullptr)
{
    ThreadCtx *c = &ctx;
    *count_out = 0;
    if (d_pos_out) *d_pos_out = nullptr;
    if (stats) *stats = bmx_stats{};
    if (keep_text) *keep_text = nullptr;
    if (n < m && !(keep_text && n > 0)) return BMX_OK;
    BMX_CUDA(cudaSetDevice(device));

    int64_t chunk = (int64_t)64 << 20;
    if (const char *e = getenv("BMX_H2D_CHUNK_MB")) {
        const long mb = atol(e);
        if (mb > 0) chunk = (int64_t)mb << 20;
    }
    chunk = std::max<int64_t>(chunk, (int64_t)m * 2);
    const int64_t nchunks = (n + chunk - 1) / chunk;
    if (int rc = ensure_streams(*c, device, (size_t)nchunks + kBounce + 3)) return rc;
    const size_t ev_first = (size_t)nchunks + kBounce + 1;  // two events: "result of scan q exported" (find-first mode)

    if (!want_pos) dev_cap = 0;
    unsigned char *d_text = nullptr;
    int64_t *d_pos = nullptr;
    int rc = BMX_OK;
    auto cleanup = [&]() {
        if (d_text) cudaFreeAsync(d_text, c->scan_stream);
        if (d_pos) cudaFreeAsync(d_pos, c->scan_stream);
        cudaStreamSynchronize(c->scan_stream);
    };
#define BMX_TRY(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            rc = fail(e_ == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "%s: %s", #call, \
                      cudaGetErrorString(e_));                                               \
            cleanup();                                                                       \
            return rc;                                                                       \
        }                                                                                    \
    } while (0)

    BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_text), (size_t)n + 16, c->scan_stream));
    if (dev_cap > 0) BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos), (size_t)dev_cap * 8, c->scan_stream));
    // the copy stream must not touch d_text before the allocation is ordered
    BMX_TRY(cudaEventRecord(c->events[(size_t)nchunks + kBounce], c->scan_stream));
    BMX_TRY(cudaStreamWaitEvent(c->copy_stream, c->events[(size_t)nchunks + kBounce], 0));

    if ((rc = bmx_scanner_set_pattern(c->scanner, pat, m, variant, c->scan_stream)) != BMX_OK ||
        (rc = bmx_scanner_begin(c->scanner, dev_cap > 0 ? d_pos : nullptr, dev_cap, c->scan_stream)) != BMX_OK) {
        cleanup();
        return rc;
    }

    const bool pinned = is_pinned_host(text);
    int staging_threads = 8;  // host threads that fill a pinned bounce buffer from pageable memory (profiles/e2e_host_memory.py)
    if (const char *e = getenv("BMX_STAGING_THREADS")) staging_threads = std::max(1, std::min(32, atoi(e)));
    staging_threads = (int)std::max(1u, std::min<unsigned>((unsigned)staging_threads, std::thread::hardware_concurrency()));
    if (!pinned && c->bounce_bytes < (size_t)std::min(chunk, n)) {
        for (int b = 0; b < kBounce; ++b) {
            if (c->bounce[b]) cudaFreeHost(c->bounce[b]);
            c->bounce[b] = nullptr;
        }
        c->bounce_bytes = 0;
        for (int b = 0; b < kBounce; ++b) BMX_TRY(cudaHostAlloc(reinterpret_cast<void **>(&c->bounce[b]), (size_t)std::min(chunk, n), cudaHostAllocDefault));
        c->bounce_bytes = (size_t)std::min(chunk, n);
    }

    int64_t scanned = 0;  // start positions < scanned are done
    int64_t q = 0;        // find-first mode: scans enqueued so far
    bool found = false;
    auto first_of = [&](int64_t qq) -> bool {   // waits for scan qq's exported {count, held, first position}
        if (cudaEventSynchronize(c->events[ev_first + (size_t)(qq & 1)]) != cudaSuccess) return false;
        const unsigned long long *h = c->scanner->h_result + 3 * (qq & 1);
        if (h[0] == 0) return false;
        *first_out = (int64_t)h[2];
        return true;
    };
    for (int64_t k = 0; k < nchunks && !found; ++k) {
        const int64_t off = k * chunk;
        const int64_t len = std::min(chunk, n - off);
        const char *src = text + off;
        if (!pinned) {
            const int b = (int)(k % kBounce);
            // the bounce buffer is free once the copy that used it kBounce chunks ago has finished
            if (k >= kBounce) BMX_TRY(cudaEventSynchronize(c->events[(size_t)nchunks + (size_t)b]));
            parallel_copy(c->bounce[b], src, (size_t)len, staging_threads);
            src = reinterpret_cast<const char *>(c->bounce[b]);
            BMX_TRY(cudaMemcpyAsync(d_text + off, src, (size_t)len, cudaMemcpyHostToDevice, c->copy_stream));
            BMX_TRY(cudaEventRecord(c->events[(size_t)nchunks + (size_t)b], c->copy_stream));
        } else {
            BMX_TRY(cudaMemcpyAsync(d_text + off, src, (size_t)len, cudaMemcpyHostToDevice, c->copy_stream));
        }
        BMX_TRY(cudaEventRecord(c->events[(size_t)k], c->copy_stream));
        BMX_TRY(cudaStreamWaitEvent(c->scan_stream, c->events[(size_t)k], 0));
        // every match lying fully inside the bytes copied so far, not yet reported
        const int64_t have = off + len;
        const int64_t span = have - scanned;
        if (span >= m) {
            if ((rc = bmx_scanner_scan(c->scanner, d_text + scanned, span, pos_base + scanned, c->scan_stream)) != BMX_OK) {
                cleanup();
                return rc;
            }
            scanned = have - m + 1;
            if (first_out) {
                int64_t *d_slot = reinterpret_cast<int64_t *>(c->scanner->d_ctrl + 5);
                if ((rc = bmx_scanner_export_result(c->scanner, d_slot, 1, c->scan_stream)) != BMX_OK) {
                    cleanup();
                    return rc;
                }
                BMX_TRY(cudaMemcpyAsync(c->scanner->h_result + 3 * (q & 1), d_slot, 24, cudaMemcpyDeviceToHost, c->scan_stream));
                BMX_TRY(cudaEventRecord(c->events[ev_first + (size_t)(q & 1)], c->scan_stream));
                if (q >= 1) found = first_of(q - 1);
                ++q;
            }
        }
    }
    uint64_t count = 0;
    bmx_stats st{};
    if ((rc = bmx_scanner_finish(c->scanner, &count, &st, c->scan_stream)) != BMX_OK) {
        cleanup();
        return rc;
    }
    *count_out = count;
    if (stats) *stats = st;
    if (first_out && !found && q >= 1) first_of(q - 1);
    if (keep_text) {
        *keep_text = d_text;
    } else {
        BMX_TRY(cudaFreeAsync(d_text, c->scan_stream));
    }
    d_text = nullptr;
    if (d_pos_out) {
        *d_pos_out = d_pos;
    } else if (d_pos) {
        cudaFreeAsync(d_pos, c->scan_stream);
    }
#undef BMX_TRY
    return BMX_OK;
}

}  // namespace

extern "C" {

int bmx_search_ex(int device, const char *text, int64_t n, const char *pat, int32_t m, int64_t *pos_out,
                  int64_t pos_cap, uint64_t *count_out, int32_t variant, bmx_stats *stats)
{
    if (!count_out || !pat) return fail(BMX_E_BADARG, "bmx_search: pat/count_out must be non-NULL");
    if (n < 0 || (!text && n > 0)) return fail(BMX_E_BADARG, "bmx_search: bad text (n=%lld)", (long long)n);
    if (pos_cap < 0) return fail(BMX_E_BADARG, "pos_cap < 0");
    if (m <= 0 || m > BMX_MAX_PATTERN)
        return fail(BMX_E_BADARG, "pattern length %d outside 1..%d (an empty pattern is rejected)", m, BMX_MAX_PATTERN);
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    *count_out = 0;
    if (stats) *stats = bmx_stats{};
    if (n < m) return BMX_OK;
    const int64_t dev_cap = pos_out ? std::min(pos_cap, n - m + 1) : 0;
    int64_t *d_pos = nullptr;
    if (int rc = ingest_and_scan(*c, device, text, n, pat, m, 0, dev_cap > 0, dev_cap, variant, &d_pos, count_out, stats)) return rc;
    int rc = BMX_OK;
    const int64_t ncopy = std::min<int64_t>((int64_t)*count_out, dev_cap);
    if (ncopy > 0) {
        cudaError_t e = cudaMemcpyAsync(pos_out, d_pos, (size_t)ncopy * 8, cudaMemcpyDeviceToHost, c->scan_stream);
        if (e != cudaSuccess) rc = fail(BMX_E_CUDA, "position read-back: %s", cudaGetErrorString(e));
    }
    if (d_pos) cudaFreeAsync(d_pos, c->scan_stream);
    const cudaError_t e = cudaStreamSynchronize(c->scan_stream);
    if (rc == BMX_OK && e != cudaSuccess) rc = fail(BMX_E_CUDA, "bmx_search: %s", cudaGetErrorString(e));
    return rc;
}

int bmx_search(const char *text, int64_t n, const char *pat, int32_t m, int64_t *pos_out, int64_t pos_cap,
               uint64_t *count_out)
{
    int device = 0;
    if (int rc = check_device(0)) return rc;
    BMX_CUDA(cudaGetDevice(&device));
    return bmx_search_ex(device, text, n, pat, m, pos_out, pos_cap, count_out, BMX_VARIANT_AUTO, nullptr);
}

// K patterns over one host text: the text crosses PCIe ONCE (chunked, overlapped with the scan for the first
// pattern); the other patterns are scans of the resident copy, microseconds per GiB next to the ingest.
int bmx_search_multi(int device, const char *text, int64_t n, int32_t npat, const char *const *pats, const int32_t *ms,
                     int64_t *const *pos_out, const int64_t *pos_cap, uint64_t *counts)
{
    if (npat < 0 || (npat > 0 && (!pats || !ms || !counts))) return fail(BMX_E_BADARG, "bmx_search_multi: NULL argument or npat < 0");
    if (n < 0 || (!text && n > 0)) return fail(BMX_E_BADARG, "bmx_search_multi: bad text (n=%lld)", (long long)n);
    for (int32_t k = 0; k < npat; ++k) {
        if (!pats[k] || ms[k] <= 0 || ms[k] > BMX_MAX_PATTERN)
            return fail(BMX_E_BADARG, "pattern %d: length %d outside 1..%d or NULL (an empty pattern is rejected)", k, ms[k], BMX_MAX_PATTERN);
        if (pos_out && pos_out[k] && (!pos_cap || pos_cap[k] < 0)) return fail(BMX_E_BADARG, "pattern %d: pos_cap < 0 or missing", k);
        counts[k] = 0;
    }
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    if (npat == 0 || n == 0) return BMX_OK;
    auto cap_of = [&](int32_t k) -> int64_t {
        if (!pos_out || !pos_out[k] || n < ms[k]) return 0;
        return std::min(pos_cap[k], n - ms[k] + 1);
    };
    unsigned char *d_text = nullptr;
    int rc = BMX_OK;
    for (int32_t k = 0; k < npat && rc == BMX_OK; ++k) {
        const int64_t dev_cap = cap_of(k);
        int64_t *d_pos = nullptr;
        if (k == 0) {
            rc = ingest_and_scan(*c, device, text, n, pats[0], ms[0], 0, dev_cap > 0, dev_cap, BMX_VARIANT_AUTO, &d_pos, &counts[0],
                                 nullptr, nullptr, &d_text);
        } else if (n >= ms[k]) {
            cudaError_t e = dev_cap > 0 ? cudaMallocAsync(reinterpret_cast<void **>(&d_pos), (size_t)dev_cap * 8, c->scan_stream) : cudaSuccess;
            if (e != cudaSuccess) {
                rc = fail(e == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "bmx_search_multi: %s", cudaGetErrorString(e));
                break;
            }
            if ((rc = bmx_scanner_set_pattern(c->scanner, pats[k], ms[k], BMX_VARIANT_AUTO, c->scan_stream)) == BMX_OK &&
                (rc = bmx_scanner_begin(c->scanner, d_pos, dev_cap, c->scan_stream)) == BMX_OK &&
                (rc = bmx_scanner_scan(c->scanner, d_text, n, 0, c->scan_stream)) == BMX_OK)
                rc = bmx_scanner_finish(c->scanner, &counts[k], nullptr, c->scan_stream);
        }
        const int64_t ncopy = rc == BMX_OK ? std::min<int64_t>((int64_t)counts[k], dev_cap) : 0;
        if (ncopy > 0 && cudaMemcpyAsync(pos_out[k], d_pos, (size_t)ncopy * 8, cudaMemcpyDeviceToHost, c->scan_stream) != cudaSuccess)
            rc = fail(BMX_E_CUDA, "bmx_search_multi: position read-back failed");
        if (d_pos) cudaFreeAsync(d_pos, c->scan_stream);
    }
    if (d_text) cudaFreeAsync(d_text, c->scan_stream);
    const cudaError_t e = cudaStreamSynchronize(c->scan_stream);
    if (rc == BMX_OK && e != cudaSuccess) rc = fail(BMX_E_CUDA, "bmx_search_multi: %s", cudaGetErrorString(e));
    return rc;
}

int bmx_find_first(const char *text, int64_t n, const char *pat, int32_t m, int64_t *first_out)
{
    if (!first_out || !pat) return fail(BMX_E_BADARG, "bmx_find_first: pat/first_out must be non-NULL");
    if (n < 0 || (!text && n > 0)) return fail(BMX_E_BADARG, "bmx_find_first: bad text (n=%lld)", (long long)n);
    if (m <= 0 || m > BMX_MAX_PATTERN)
        return fail(BMX_E_BADARG, "pattern length %d outside 1..%d (an empty pattern is rejected)", m, BMX_MAX_PATTERN);
    *first_out = -1;
    int device = 0;
    if (int rc = check_device(0)) return rc;
    BMX_CUDA(cudaGetDevice(&device));
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    if (n < m) return BMX_OK;
    int64_t *d_pos = nullptr;
    uint64_t count = 0;
    if (int rc = ingest_and_scan(*c, device, text, n, pat, m, 0, true, 1, BMX_VARIANT_AUTO, &d_pos, &count, nullptr, first_out)) return rc;
    if (d_pos) cudaFreeAsync(d_pos, c->scan_stream);
    // copies of chunks behind the match may still be in flight: the caller's buffer must be free on return
    cudaError_t e = cudaStreamSynchronize(c->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->scan_stream);
    if (e != cudaSuccess) return fail(BMX_E_CUDA, "bmx_find_first: %s", cudaGetErrorString(e));
    return BMX_OK;
}

int bmx_search_partitions(const char *text, const char *pat, const int32_t *se, int32_t *ans, const int32_t *gs,
                          const int32_t *bs, int32_t m, int32_t nparts)
{
    if (!pat || !se || !ans || nparts < 0 || (!text && nparts > 0))
        return fail(BMX_E_BADARG, "bmx_search_partitions: NULL argument or nparts < 0");
    if (m <= 0 || m > BMX_MAX_PATTERN)
        return fail(BMX_E_BADARG, "pattern length %d outside 1..%d (an empty pattern is rejected)", m, BMX_MAX_PATTERN);
    if (nparts == 0) return BMX_OK;

    // The reference hands its own tables to the kernel (BoyreMoore.cpp:268-269).  The device scan
    // keeps its own copy, so caller tables are only checked: a wrong table must not go unnoticed.
    if (gs || bs) {
        std::vector<int32_t> good((size_t)m);
        int32_t bad[256];
        build_bad_table(reinterpret_cast<const unsigned char *>(pat), m, bad);
        build_good_table(reinterpret_cast<const unsigned char *>(pat), m, good.data());
        if (gs)
            for (int32_t k = 1; k < m; ++k)
                if (gs[k] != good[(size_t)k]) return fail(BMX_E_TABLES, "gstable[%d] = %d, expected %d", k, gs[k], good[(size_t)k]);
        if (bs)
            for (int c = 0; c < 128; ++c)
                if (bs[c] != bad[c]) return fail(BMX_E_TABLES, "bstable[%d] = %d, expected %d", c, bs[c], bad[c]);
    }

    int64_t lo = INT64_MAX, hi = -1;
    for (int32_t id = 0; id < nparts; ++id) {
        if (se[2 * id] < 0) return fail(BMX_E_BADARG, "se[%d] = %d is negative", 2 * id, se[2 * id]);
        lo = std::min<int64_t>(lo, se[2 * id]);
        hi = std::max<int64_t>(hi, se[2 * id + 1]);
    }
    for (int32_t id = 0; id < nparts; ++id) ans[id] = 0;  // kernel1.cl:6
    const int64_t span = hi - lo + 1;
    if (span < m) return BMX_OK;

    int device = 0;
    if (int rc = check_device(0)) return rc;
    BMX_CUDA(cudaGetDevice(&device));
    ThreadCtx *c = nullptr;
    if (int rc = get_ctx(device, &c)) return rc;
    if (int rc = ensure_streams(*c, device, 1)) return rc;
    cudaStream_t st = c->scan_stream;

    unsigned char *d_text = nullptr;
    int64_t *d_pos = nullptr;
    int32_t *d_se = nullptr, *d_ans = nullptr;
    int rc = BMX_OK;
    auto cleanup = [&]() {
        if (d_text) cudaFreeAsync(d_text, st);
        if (d_pos) cudaFreeAsync(d_pos, st);
        if (d_se) cudaFreeAsync(d_se, st);
        if (d_ans) cudaFreeAsync(d_ans, st);
        cudaStreamSynchronize(st);
    };
#define BMX_TRY(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            rc = fail(e_ == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "%s: %s", #call, \
                      cudaGetErrorString(e_));                                               \
            cleanup();                                                                       \
            return rc;                                                                       \
        }                                                                                    \
    } while (0)
    const int64_t max_hits = span - m + 1;
    BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_text), (size_t)span + 16, st));
    BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_pos), (size_t)max_hits * 8, st));
    BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_se), (size_t)nparts * 8, st));
    BMX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_ans), (size_t)nparts * 4, st));
    BMX_TRY(cudaMemcpyAsync(d_text, text + lo, (size_t)span, cudaMemcpyHostToDevice, st));
    BMX_TRY(cudaMemcpyAsync(d_se, se, (size_t)nparts * 8, cudaMemcpyHostToDevice, st));
    if ((rc = bmx_scanner_set_pattern(c->scanner, pat, m, BMX_VARIANT_AUTO, st)) != BMX_OK ||
        (rc = bmx_scanner_begin(c->scanner, d_pos, max_hits, st)) != BMX_OK ||
        (rc = bmx_scanner_scan(c->scanner, d_text, span, lo, st)) != BMX_OK) {
        cleanup();
        return rc;
    }
    // the running count lives in the scanner's carry slot after one scan: slot 1
    if ((rc = launch_partition_count(d_pos, result_slot(c->scanner), max_hits, d_se, d_ans, m,
                                     nparts, st)) != BMX_OK) {
        cleanup();
        return rc;
    }
    BMX_TRY(cudaMemcpyAsync(ans, d_ans, (size_t)nparts * 4, cudaMemcpyDeviceToHost, st));
    BMX_TRY(cudaStreamSynchronize(st));
    cleanup();
#undef BMX_TRY
    return BMX_OK;
}

int bmx_synth_fill_device(void *d_text, int64_t offset, int64_t len, uint64_t seed, const unsigned char *alphabet,
                          int32_t sigma, void *stream)
{
    if (len < 0 || offset < 0 || (!d_text && len > 0) || !alphabet || sigma < 1 || sigma > 256)
        return fail(BMX_E_BADARG, "bmx_synth_fill_device: bad argument");
    if (int rc = check_device(0)) return rc;
    return launch_synth_fill(d_text, offset, len, seed, alphabet, sigma, stream);
}


// ---------------------------------------------------------------------------------------------
// single-process multi-GPU search over host text: contiguous shards + (m-1)-byte halo,
// one host thread per GPU for the ingest, positions gathered in shard order (= ascending)
// ---------------------------------------------------------------------------------------------
struct bmx_mg {
    std::vector<int> devices;
    std::vector<ThreadCtx *> ctx;
};

int bmx_mg_create(int ngpus, bmx_mg **out)
{
    if (!out) return fail(BMX_E_BADARG, "bmx_mg_create: out is NULL");
    *out = nullptr;
    if (int rc = check_device(0)) return rc;
    int have = 0;
    BMX_CUDA(cudaGetDeviceCount(&have));
    if (ngpus <= 0) ngpus = have;
    if (ngpus > have) return fail(BMX_E_BADARG, "bmx_mg_create: %d GPUs requested, %d visible", ngpus, have);
    int keep = 0;
    cudaGetDevice(&keep);
    bmx_mg *mg = new (std::nothrow) bmx_mg();
    if (!mg) return fail(BMX_E_NOMEM, "out of host memory");
    for (int d = 0; d < ngpus; ++d) {
        ThreadCtx *c = new (std::nothrow) ThreadCtx();
        int rc = c ? BMX_OK : fail(BMX_E_NOMEM, "out of host memory");
        if (rc == BMX_OK) rc = bmx_scanner_create(d, &c->scanner);
        if (rc != BMX_OK) {
            delete c;
            bmx_mg_destroy(mg);
            cudaSetDevice(keep);
            return rc;
        }
        mg->devices.push_back(d);
        mg->ctx.push_back(c);
    }
    cudaSetDevice(keep);
    *out = mg;
    return BMX_OK;
}

void bmx_mg_destroy(bmx_mg *mg)
{
    if (!mg) return;
    int keep = 0;
    cudaGetDevice(&keep);
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        ThreadCtx *c = mg->ctx[i];
        cudaSetDevice(mg->devices[i]);
        if (c->scanner) bmx_scanner_destroy(c->scanner);
        if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
        if (c->scan_stream) cudaStreamDestroy(c->scan_stream);
        for (cudaEvent_t e : c->events) cudaEventDestroy(e);
        for (int b = 0; b < kBounce; ++b)
            if (c->bounce[b]) cudaFreeHost(c->bounce[b]);
        delete c;
    }
    cudaSetDevice(keep);
    delete mg;
}

int bmx_mg_device_count(const bmx_mg *mg) { return mg ? (int)mg->devices.size() : 0; }

int bmx_mg_search(bmx_mg *mg, const char *text, int64_t n, const char *pat, int32_t m, int64_t *pos_out, int64_t pos_cap,
                  uint64_t *count_out, uint64_t *shard_counts)
{
    if (!mg || !count_out || !pat) return fail(BMX_E_BADARG, "bmx_mg_search: NULL argument");
    if (n < 0 || (!text && n > 0)) return fail(BMX_E_BADARG, "bmx_mg_search: bad text (n=%lld)", (long long)n);
    if (pos_cap < 0) return fail(BMX_E_BADARG, "pos_cap < 0");
    if (m <= 0 || m > BMX_MAX_PATTERN)
        return fail(BMX_E_BADARG, "pattern length %d outside 1..%d (an empty pattern is rejected)", m, BMX_MAX_PATTERN);
    const int R = (int)mg->devices.size();
    *count_out = 0;
    for (int r = 0; r < R && shard_counts; ++r) shard_counts[r] = 0;
    if (n < m) return BMX_OK;

    // rank r owns the START positions [lo_r, hi_r) and reads (m-1) bytes of halo behind hi_r
    int64_t per = (n + R - 1) / R;
    per = (per + 15) & ~int64_t(15);
    struct Shard {
        int64_t lo = 0, hi = 0, end = 0, cap = 0;
        int64_t *d_pos = nullptr;
        uint64_t count = 0;
        int rc = BMX_OK;
        std::string err;
    };
    std::vector<Shard> sh((size_t)R);
    std::vector<std::thread> workers;
    for (int r = 0; r < R; ++r) {
        Shard &s = sh[(size_t)r];
        s.lo = std::min<int64_t>(n, (int64_t)r * per);
        s.hi = std::min<int64_t>(n, s.lo + per);
        s.end = std::min<int64_t>(n, s.hi + m - 1);
        s.cap = pos_out ? std::min<int64_t>(pos_cap, s.hi - s.lo) : 0;
        if (s.end - s.lo < m) continue;
        workers.emplace_back([&, r]() {
            Shard &w = sh[(size_t)r];
            w.rc = ingest_and_scan(*mg->ctx[(size_t)r], mg->devices[(size_t)r], text + w.lo, w.end - w.lo, pat, m, w.lo, w.cap > 0,
                                   w.cap, BMX_VARIANT_AUTO, &w.d_pos, &w.count, nullptr);
            if (w.rc != BMX_OK) w.err = bmx_last_error();
        });
    }
    for (auto &t : workers) t.join();

    int rc = BMX_OK;
    std::string err;
    uint64_t total = 0;
    for (int r = 0; r < R; ++r) {
        if (sh[(size_t)r].rc != BMX_OK && rc == BMX_OK) {
            rc = sh[(size_t)r].rc;
            err = sh[(size_t)r].err;
        }
        total += sh[(size_t)r].count;
        if (shard_counts) shard_counts[r] = sh[(size_t)r].count;
    }
    // gather: shard lists are ascending and shards are ordered, so concatenation is the sorted result
    int keep = 0;
    cudaGetDevice(&keep);
    int64_t off = 0;
    for (int r = 0; r < R; ++r) {
        Shard &s = sh[(size_t)r];
        cudaSetDevice(mg->devices[(size_t)r]);
        cudaStream_t st = mg->ctx[(size_t)r]->scan_stream;
        if (rc == BMX_OK && s.d_pos) {
            const int64_t have = std::min<int64_t>((int64_t)s.count, s.cap);
            const int64_t ncopy = std::max<int64_t>(0, std::min<int64_t>(have, pos_cap - off));
            if (ncopy > 0 && cudaMemcpyAsync(pos_out + off, s.d_pos, (size_t)ncopy * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
                rc = BMX_E_CUDA;
                err = "position gather failed";
            }
        }
        off += (int64_t)s.count;
        if (s.d_pos) cudaFreeAsync(s.d_pos, st);
    }
    for (int r = 0; r < R; ++r) {
        cudaSetDevice(mg->devices[(size_t)r]);
        if (mg->ctx[(size_t)r]->scan_stream && cudaStreamSynchronize(mg->ctx[(size_t)r]->scan_stream) != cudaSuccess && rc == BMX_OK) {
            rc = BMX_E_CUDA;
            err = "stream synchronisation failed";
        }
    }
    cudaSetDevice(keep);
    if (rc != BMX_OK) return fail(rc, "bmx_mg_search: %s", err.c_str());
    *count_out = total;
    return BMX_OK;
}

}  // extern "C"
