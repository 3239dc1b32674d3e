// write_patterns.cu -- which store pattern reaches the HBM write peak on B200?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_patterns write_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// P1: grid-stride, 16 B per thread per iteration (what a fill kernel does)
__global__ void p1(longlong2 *out, size_t n2, long long base)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
}
// P2: each warp owns `unit` consecutive elements at a time (round robin over units), 512 B per warp instruction
__global__ void p2(longlong2 *out, size_t n2, long long base, size_t unit2)
{
    const int lane = threadIdx.x & 31;
    const size_t gw = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t u = gw; u * unit2 < n2; u += nw) {
        const size_t lo = u * unit2, hi = lo + unit2 < n2 ? lo + unit2 : n2;
#pragma unroll 4
        for (size_t i = lo + lane; i < hi; i += 32) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
    }
}
// P3: like P2 but each warp walks `chain` consecutive units before jumping (= the expand kernel's items of 16 segments)
__global__ void p3(longlong2 *out, size_t n2, long long base, size_t unit2, int chain)
{
    const int lane = threadIdx.x & 31;
    const size_t gw = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t item2 = unit2 * chain;
    for (size_t it = gw; it * item2 < n2; it += nw)
        for (int c = 0; c < chain; ++c) {
            const size_t lo = it * item2 + c * unit2, hi = lo + unit2 < n2 ? lo + unit2 : n2;
#pragma unroll 4
            for (size_t i = lo + lane; i < hi; i += 32) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
        }
}


// P4: non-persistent, one CTA per `per_cta2` pairs; warps take consecutive 1024-pair (16 KiB) units
__global__ void p4(longlong2 *out, size_t n2, long long base, size_t per_cta2)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const size_t cta_lo = blockIdx.x * per_cta2;
    for (size_t u = warp; u * 1024 < per_cta2; u += nwarp) {
        const size_t lo = cta_lo + u * 1024, hi = lo + 1024 < n2 ? lo + 1024 : n2;
#pragma unroll 4
        for (size_t i = lo + lane; i < hi; i += 32) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
    }
}
// P5: non-persistent, one CTA per `per_cta2` pairs, threads interleaved over the whole CTA range
__global__ void p5(longlong2 *out, size_t n2, long long base, size_t per_cta2)
{
    const size_t lo = blockIdx.x * per_cta2, hi = lo + per_cta2 < n2 ? lo + per_cta2 : n2;
#pragma unroll 4
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
}


// P6: persistent grid, CTAs take 16 KiB * upc units in order from a global ticket counter
__global__ void p6(longlong2 *out, size_t n2, long long base, size_t per_cta2, unsigned *ticket)
{
    __shared__ unsigned s_t;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_t = atomicAdd(ticket, 1u);
        __syncthreads();
        const size_t lo = (size_t)s_t * per_cta2;
        if (lo >= n2) return;
        const size_t hi = lo + per_cta2 < n2 ? lo + per_cta2 : n2;
#pragma unroll 4
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
    }
}
// P7: persistent grid, each WARP takes 16 KiB units in order from a global ticket counter
__global__ void p7(longlong2 *out, size_t n2, long long base, unsigned *ticket)
{
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        const size_t lo = (size_t)t * 1024;
        if (lo >= n2) return;
#pragma unroll 4
        for (size_t i = lo + lane; i < lo + 1024; i += 32) out[i] = make_longlong2(base + 2 * i, base + 2 * i + 1);
    }
}

template <typename L>
static void timeit(const char *name, size_t bytes, L launch)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-44s %8.1f us  %8.1f GB/s  (%s)\n", name, ms / 10 * 1e3, bytes / (ms / 10 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t n = (size_t)1 << 30;        // int64 elements = 8 GiB
    longlong2 *out; cudaMalloc(&out, n * 8);
    const size_t n2 = n / 2;
    timeit("P1 grid-stride 148*8 x 256", n * 8, [&] { p1<<<148 * 8, 256>>>(out, n2, 5); });
    timeit("P1 grid-stride 740 x 256", n * 8, [&] { p1<<<740, 256>>>(out, n2, 5); });
    timeit("P1 one block per 16 KiB (huge grid)", n * 8, [&] { p1<<<(unsigned)(n2 / 1024), 256>>>(out, n2, 5); });
    for (size_t unit : {2048ul, 32768ul, 262144ul})
        for (int grid : {740, 148 * 8}) {
            char nm[96]; snprintf(nm, sizeof nm, "P2 warp per %zu-element unit, grid %d", unit, grid);
            timeit(nm, n * 8, [&] { p2<<<grid, 256>>>(out, n2, 5, unit / 2); });
        }
    timeit("P3 warp walks 16 x 2048-element units, 740", n * 8, [&] { p3<<<740, 256>>>(out, n2, 5, 1024, 16); });
    timeit("P3 warp walks 16 x 2048-element units, 1184", n * 8, [&] { p3<<<1184, 256>>>(out, n2, 5, 1024, 16); });
    for (size_t per : {1024ul, 16384ul, 131072ul})
        for (int threads : {32, 128, 256}) {
            char nm[96];
            snprintf(nm, sizeof nm, "P4 CTA(%d thr) per %zu pairs, warp per 16 KiB", threads, per);
            timeit(nm, n * 8, [&] { p4<<<(unsigned)(n2 / per), threads>>>(out, n2, 5, per); });
            snprintf(nm, sizeof nm, "P5 CTA(%d thr) per %zu pairs, interleaved", threads, per);
            timeit(nm, n * 8, [&] { p5<<<(unsigned)(n2 / per), threads>>>(out, n2, 5, per); });
        }
    unsigned *ticket; cudaMalloc(&ticket, 4);
    for (size_t per : {1024ul, 16384ul})
        for (int grid : {740, 1184}) {
            char nm[96];
            snprintf(nm, sizeof nm, "P6 persistent %d CTAs, ticket per %zu pairs", grid, per);
            timeit(nm, n * 8, [&] { cudaMemsetAsync(ticket, 0, 4); p6<<<grid, 256>>>(out, n2, 5, per, ticket); });
        }
    timeit("P7 persistent 740 CTAs, warp ticket per 16 KiB", n * 8, [&] { cudaMemsetAsync(ticket, 0, 4); p7<<<740, 256>>>(out, n2, 5, ticket); });
    timeit("P7 persistent 1184 CTAs, warp ticket per 16 KiB", n * 8, [&] { cudaMemsetAsync(ticket, 0, 4); p7<<<1184, 256>>>(out, n2, 5, ticket); });
    return 0;
}
