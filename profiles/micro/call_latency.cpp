// call_latency.cpp -- wall time per call of the C ABI on a small text, with no Python in the loop.
// The reference's own inputs are 60 B .. 560 KB (BoyreMoore/**/input*.txt) and its timed window is kernel creation +
// launch + an 8-byte read (BoyreMoore.cpp:258-290): this is the like-for-like number for that window.
//   call_latency <text-file> <pattern> [reps]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "bmx.h"

template <typename F>
static double us_per_call(F f, int reps)
{
    for (int i = 0; i < 50; ++i) f();
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) f();
    cudaDeviceSynchronize();
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: call_latency <text-file> <pattern> [reps]\n");
        return 2;
    }
    std::ifstream ifs(argv[1], std::ios::binary);
    const std::string text((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    const std::string pat = argv[2];
    const int reps = argc > 3 ? atoi(argv[3]) : 2000;
    const int64_t n = (int64_t)text.size();
    const int32_t m = (int32_t)pat.size();
    void *d_text = nullptr;
    int64_t *d_pos = nullptr;
    const int64_t cap = 1 << 16;
    cudaMalloc(&d_text, (size_t)n + 16);
    cudaMalloc(reinterpret_cast<void **>(&d_pos), (size_t)cap * 8);
    cudaMemcpy(d_text, text.data(), (size_t)n, cudaMemcpyHostToDevice);
    std::vector<int64_t> h_pos((size_t)cap);
    uint64_t count = 0;
    int64_t first = -1;
    int rc = bmx_search_device(d_text, n, pat.data(), m, d_pos, cap, &count, nullptr, nullptr);
    if (rc != BMX_OK) {
        fprintf(stderr, "bmx_search_device: %d %s\n", rc, bmx_last_error());
        return 1;
    }
    printf("%s (%lld bytes), pattern '%s': %llu hits\n", argv[1], (long long)n, pat.c_str(), (unsigned long long)count);
    printf("bmx_search_device, positions   (text resident)   %8.1f us per call\n",
           us_per_call([&] { bmx_search_device(d_text, n, pat.data(), m, d_pos, cap, &count, nullptr, nullptr); }, reps));
    printf("bmx_search_device, count only  (text resident)   %8.1f us per call\n",
           us_per_call([&] { bmx_search_device(d_text, n, pat.data(), m, nullptr, 0, &count, nullptr, nullptr); }, reps));
    float ms = 0.f;
    printf("bmx_search_device, positions + device_ms events  %8.1f us per call\n",
           us_per_call([&] { bmx_search_device(d_text, n, pat.data(), m, d_pos, cap, &count, &ms, nullptr); }, reps));
    printf("bmx_find_first_device          (text resident)   %8.1f us per call\n",
           us_per_call([&] { bmx_find_first_device(d_text, n, pat.data(), m, &first, nullptr); }, reps));
    printf("bmx_search, positions          (pageable host)   %8.1f us per call\n",
           us_per_call([&] { bmx_search(text.data(), n, pat.data(), m, h_pos.data(), cap, &count); }, reps / 4 + 1));
    char *pinned = nullptr;
    cudaHostAlloc(reinterpret_cast<void **>(&pinned), (size_t)n, cudaHostAllocDefault);
    memcpy(pinned, text.data(), (size_t)n);
    printf("bmx_search, positions          (pinned host)     %8.1f us per call\n",
           us_per_call([&] { bmx_search(pinned, n, pat.data(), m, h_pos.data(), cap, &count); }, reps / 4 + 1));
    {
        // the same steps by hand through the streaming scanner API on one stream: the floor of the host path
        bmx_scanner *sc = nullptr;
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        int64_t *h_land = nullptr;
        cudaHostAlloc(reinterpret_cast<void **>(&h_land), 8192 * 8, cudaHostAllocDefault);
        if (bmx_scanner_create(0, &sc) == BMX_OK) {
            bmx_scanner_set_pattern(sc, pat.data(), m, BMX_VARIANT_AUTO, st);
            printf("by hand: H2D + scanner + D2H, one stream (pinned)  %6.1f us per call\n",
                   us_per_call([&] {
                       cudaMemcpyAsync(d_text, pinned, (size_t)n, cudaMemcpyHostToDevice, st);
                       bmx_scanner_set_pattern(sc, pat.data(), m, BMX_VARIANT_AUTO, st);
                       bmx_scanner_begin(sc, d_pos, cap, st);
                       bmx_scanner_scan(sc, d_text, n, 0, st);
                       cudaMemcpyAsync(h_land, d_pos, 8192 * 8, cudaMemcpyDeviceToHost, st);
                       bmx_scanner_finish(sc, &count, nullptr, st);
                   }, reps / 4 + 1));
            printf("by hand: scanner only (resident), own stream       %6.1f us per call\n",
                   us_per_call([&] {
                       bmx_scanner_begin(sc, d_pos, cap, st);
                       bmx_scanner_scan(sc, d_text, n, 0, st);
                       bmx_scanner_finish(sc, &count, nullptr, st);
                   }, reps / 4 + 1));
            printf("by hand: H2D + scanner, one stream (pinned)        %6.1f us per call\n",
                   us_per_call([&] {
                       cudaMemcpyAsync(d_text, pinned, (size_t)n, cudaMemcpyHostToDevice, st);
                       bmx_scanner_begin(sc, d_pos, cap, st);
                       bmx_scanner_scan(sc, d_text, n, 0, st);
                       bmx_scanner_finish(sc, &count, nullptr, st);
                   }, reps / 4 + 1));
            bmx_scanner_destroy(sc);
        }
    }
    printf("first match at %lld, last count %llu\n", (long long)first, (unsigned long long)count);
    return 0;
}
