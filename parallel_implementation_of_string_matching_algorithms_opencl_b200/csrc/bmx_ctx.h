// bmx_ctx.h -- per (host thread, device) context behind the convenience entry points: one cached scanner, two
// streams, events, pinned bounce buffers and two cached device buffers (text, positions).  Not installed.
//
// Device memory policy: the library never touches the attributes of the device's default memory pool and never
// sizes a buffer by the worst case (8 bytes per text byte).  The cached buffers are plain cudaMalloc blocks that
// grow on demand, are reused by later calls of the same thread and are given back by bmx_release_memory() or
// when the thread exits.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "bmx_internal.h"
#include "bmx_scanner.h"

#define BMX_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            (void)cudaGetLastError();                                                                    \
            return bmx::fail(e_ == cudaErrorMemoryAllocation ? BMX_E_NOMEM : BMX_E_CUDA, "%s: %s", #call, \
                             cudaGetErrorString(e_));                                                    \
        }                                                                                                \
    } while (0)

namespace bmx {

constexpr int kMaxDevices = 64;
constexpr int kBounce = 3;
constexpr int64_t kSmallSpec = 8192;   // positions read back together with the count (64 KiB)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ThreadCtx {
    int device = -1;
    bmx_scanner *scanner = nullptr;
    cudaStream_t copy_stream = nullptr, scan_stream = nullptr;
    std::vector<cudaEvent_t> events;
    unsigned char *bounce[kBounce] = {nullptr, nullptr, nullptr};
    size_t bounce_bytes = 0;
    DevBuf text, pos, misc, aux;
    int64_t *h_small = nullptr;   // pinned landing area for the speculative result read of small host texts (kSmallSpec entries)
    ThreadCtx() = default;
    ThreadCtx(const ThreadCtx &) = delete;
    ThreadCtx &operator=(const ThreadCtx &) = delete;
    ~ThreadCtx();            // a host thread that exits gives everything back
    void release_buffers();  // text / positions / bounce buffers only (bmx_release_memory)
};

int check_device(int device);
int get_ctx(int device, ThreadCtx **out);                        // the calling thread's context for `device`
int ensure_streams(ThreadCtx &c, int device, size_t nevents);
// Grows b to at least `bytes` (never shrinks).  The old block may still be in use by work enqueued on the
// context's streams, so they are drained before it is freed.  NOMEM leaves b empty.
int ensure_buf(ThreadCtx &c, DevBuf &b, size_t bytes);
int scanner_begin_find(bmx_scanner *s, void *stream);   // find-first search: count-only kernels + early stop (bmx_abi.cu)
// K patterns in one pass over device-resident text (bmx_multipat.cu)
bool multi_set_eligible(int32_t npat, const int32_t *ms);
int multi_scan_resident(ThreadCtx &c, const unsigned char *d_text, int64_t n, int32_t npat, const char *const *pats,
                        const int32_t *ms, int64_t *const *d_out, const int64_t *caps, uint64_t *counts, cudaStream_t st);

}  // namespace bmx
