// Development probe: what do driver allocations cost on this box? (explains set-up outliers when the stream-ordered pool must grow)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); }
int main() {
    cudaFree(0);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long th = ~0ULL; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &th);
    for (size_t gb : {1, 1, 4, 4}) {
        void* p; auto t = clk::now();
        cudaMalloc(&p, gb << 30); double a = ms(t); t = clk::now(); cudaFree(p);
        printf("cudaMalloc %zu GB: %.2f ms, cudaFree %.2f ms\n", gb, a, ms(t));
    }
    for (size_t gb : {1, 1, 2, 1, 4, 4}) {
        void* p; auto t = clk::now();
        cudaMallocAsync(&p, gb << 30, s); cudaStreamSynchronize(s); double a = ms(t); t = clk::now();
        cudaFreeAsync(p, s); cudaStreamSynchronize(s);
        printf("cudaMallocAsync %zu GB: %.2f ms, free %.2f ms\n", gb, a, ms(t));
    }
    // fragmentation: many 64 MB blocks, free every other one, then ask for 1 GB
    void* q[64];
    for (int i = 0; i < 64; i++) cudaMallocAsync(&q[i], 64 << 20, s);
    for (int i = 0; i < 64; i += 2) cudaFreeAsync(q[i], s);
    cudaStreamSynchronize(s);
    { void* p; auto t = clk::now(); cudaMallocAsync(&p, 3ull << 30, s); cudaStreamSynchronize(s); printf("after fragmentation, 3 GB async: %.2f ms\n", ms(t)); cudaFreeAsync(p, s); }
    for (int i = 0; i < 3; i++) { void* h; auto t = clk::now(); cudaMallocHost(&h, 64); double a = ms(t); t = clk::now(); cudaFreeHost(h); printf("cudaMallocHost 64 B: %.3f ms, free %.3f ms\n", a, ms(t)); }
    return 0;
}
