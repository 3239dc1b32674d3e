// read_patterns.cu -- how fast can a B200 stream 4 GiB from HBM with (a) plain 16-byte loads and
// (b) the scan kernel's TMA ring with almost no compute?  The ceiling the scan kernel is measured against.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o read_patterns read_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a) grid-stride LDG.128, xor-reduce
__global__ void ldg_read(const uint4 *in, size_t n4, unsigned *out)
{
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(in + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *out = acc;
}
// (a2) ticket-ordered 32 KiB units per CTA, LDG.128
__global__ void ldg_ticket(const uint4 *in, size_t n4, unsigned *out, unsigned *ticket)
{
    __shared__ unsigned s_t;
    uint32_t acc = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        const size_t lo = (size_t)s_t * 2048;
        if (lo >= n4) break;
#pragma unroll 8
        for (size_t i = lo + threadIdx.x; i < lo + 2048; i += blockDim.x) {
            const uint4 v = __ldg(in + i);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *out = acc;
}
// (b) TMA ring: producer lane + 8 consumer warps, 32 KiB tiles, S stages, ticket order
template <int S>
__global__ void __launch_bounds__(288, 2) tma_read(const uint8_t *in, uint32_t ntiles, unsigned *out, unsigned *ticket)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = (uint64_t *)smem, *empty = full + 8;
    int *slot = (int *)(empty + 8);
    uint8_t *stages = smem + 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [](uint64_t *bar, uint32_t parity) {
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
    };
    if (warp == 8) {
        if (lane) return;
        uint32_t next = atomicAdd(ticket, 1u);
        for (uint32_t it = 0;; ++it) {
            const uint32_t s = it % S, t = next;
            const bool live = t < ntiles;
            if (live) next = atomicAdd(ticket, 1u);
            wait(&empty[s], ((it / S) & 1u) ^ 1u);
            slot[s] = live ? (int)t : -1;
            if (!live) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory"); break; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(32768) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(stages + s * 32768)), "l"(in + (size_t)t * 32768), "r"(32768), "r"(smem_u32(&full[s])) : "memory");
        }
        return;
    }
    uint32_t acc = 0;
    for (uint32_t it = 0;; ++it) {
        const uint32_t s = it % S;
        wait(&full[s], (it / S) & 1u);
        if (slot[s] < 0) break;
        const uint4 *p = (const uint4 *)(stages + s * 32768) + warp * 256 + lane;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const uint4 v = p[k * 32]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 0x12345678u) *out = acc;
}

template <typename L>
static void timeit(const char *name, size_t bytes, L launch)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-46s %8.1f us  %8.1f GB/s  (%s)\n", name, ms / 10 * 1e3, bytes / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t n = (size_t)4 << 30;
    uint8_t *in; cudaMalloc(&in, n); cudaMemset(in, 1, n);
    unsigned *out, *ticket; cudaMalloc(&out, 4); cudaMalloc(&ticket, 4);
    timeit("LDG.128 grid-stride 148*8 x 256", n, [&] { ldg_read<<<148 * 8, 256>>>((const uint4 *)in, n / 16, out); });
    timeit("LDG.128 grid-stride 148*16 x 128", n, [&] { ldg_read<<<148 * 16, 128>>>((const uint4 *)in, n / 16, out); });
    timeit("LDG.128 ticket per 32 KiB, 148*8 x 256", n, [&] { cudaMemsetAsync(ticket, 0, 4); ldg_ticket<<<148 * 8, 256>>>((const uint4 *)in, n / 16, out, ticket); });
    cudaFuncSetAttribute(tma_read<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 + 3 * 32768);
    cudaFuncSetAttribute(tma_read<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 + 2 * 32768);
    timeit("TMA ring 3 x 32 KiB, 296 CTAs, xor only", n, [&] { cudaMemsetAsync(ticket, 0, 4); tma_read<3><<<296, 288, 256 + 3 * 32768>>>(in, (uint32_t)(n / 32768), out, ticket); });
    timeit("TMA ring 2 x 32 KiB, 296 CTAs, xor only", n, [&] { cudaMemsetAsync(ticket, 0, 4); tma_read<2><<<296, 288, 256 + 2 * 32768>>>(in, (uint32_t)(n / 32768), out, ticket); });
    return 0;
}
