// stream_floor.cu -- what one stream costs per call on this box, without any library code: the floor under the
// small-host-text path (H2D of the text, two dependent kernels, D2H of the positions, one synchronisation).
//   nvcc -arch=sm_100a -O2 -o stream_floor stream_floor.cu && ./stream_floor [text bytes] [result bytes]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void tiny(int *p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
__global__ void busy(int *p, long long cycles)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1;
}

template <typename F>
static double us(F f, int reps = 2000)
{
    for (int i = 0; i < 100; ++i) f();
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) f();
    cudaDeviceSynchronize();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main(int argc, char **argv)
{
    const size_t n = argc > 1 ? atol(argv[1]) : 500007, r = argc > 2 ? atol(argv[2]) : 65536;
    char *h_text, *d_text, *h_res, *d_res;
    int *d_flag;
    cudaHostAlloc(&h_text, n, cudaHostAllocDefault);
    cudaHostAlloc(&h_res, r, cudaHostAllocDefault);
    cudaMalloc(&d_text, n + 16);
    cudaMalloc(&d_res, r);
    cudaMalloc(&d_flag, 4);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    const long long c12 = 12 * 1965;   // ~12 us of kernel time in two launches, like scan + expand
    printf("sync only                                   %6.1f us\n", us([&] { cudaStreamSynchronize(st); }));
    printf("kernel + sync                               %6.1f us\n", us([&] { tiny<<<1, 32, 0, st>>>(d_flag); cudaStreamSynchronize(st); }));
    printf("2 kernels (8 + 4 us busy) + sync            %6.1f us\n", us([&] { busy<<<16, 288, 0, st>>>(d_flag, c12 * 2 / 3); busy<<<64, 256, 0, st>>>(d_flag, c12 / 3); cudaStreamSynchronize(st); }));
    printf("H2D %7zu B + sync                        %6.1f us\n", n, us([&] { cudaMemcpyAsync(d_text, h_text, n, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); }));
    printf("D2H %7zu B + sync                        %6.1f us\n", r, us([&] { cudaMemcpyAsync(h_res, d_res, r, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }));
    printf("H2D + 2 kernels + sync                      %6.1f us\n", us([&] { cudaMemcpyAsync(d_text, h_text, n, cudaMemcpyHostToDevice, st); busy<<<16, 288, 0, st>>>(d_flag, c12 * 2 / 3); busy<<<64, 256, 0, st>>>(d_flag, c12 / 3); cudaStreamSynchronize(st); }));
    printf("H2D + 2 kernels + D2H + sync                %6.1f us\n", us([&] { cudaMemcpyAsync(d_text, h_text, n, cudaMemcpyHostToDevice, st); busy<<<16, 288, 0, st>>>(d_flag, c12 * 2 / 3); busy<<<64, 256, 0, st>>>(d_flag, c12 / 3); cudaMemcpyAsync(h_res, d_res, r, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }));
    printf("H2D + 2 kernels + sync + D2H + sync         %6.1f us\n", us([&] { cudaMemcpyAsync(d_text, h_text, n, cudaMemcpyHostToDevice, st); busy<<<16, 288, 0, st>>>(d_flag, c12 * 2 / 3); busy<<<64, 256, 0, st>>>(d_flag, c12 / 3); cudaStreamSynchronize(st); cudaMemcpyAsync(h_res, d_res, r, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }));
    // zero-copy flavour: the kernels would read the pinned text and write the pinned result themselves
    printf("2 kernels + sync (zero-copy would add PCIe reads inside the kernels)  see line 3\n");
    return 0;
}
