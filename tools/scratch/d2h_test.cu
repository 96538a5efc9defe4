// Development probe: D2H bandwidth of a pitched (2-D) copy vs a contiguous copy, pinned host memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <chrono>
int main()
{
    const size_t rows = 1029ull * 1029, dense = 1029 * 4, pitch = 1056 * 4;
    void *d, *h;
    cudaMalloc(&d, rows * pitch);
    cudaMallocHost(&h, rows * pitch);
    cudaMemset(d, 1, rows * pitch);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int rep = 0; rep < 2; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        cudaMemcpy2DAsync(h, dense, d, pitch, dense, rows, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        cudaMemcpyAsync(h, d, rows * dense, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        auto t2 = std::chrono::steady_clock::now();
        double a = std::chrono::duration<double>(t1 - t0).count(), b = std::chrono::duration<double>(t2 - t1).count();
        printf("2D pitched: %.2f GB/s   1D contiguous: %.2f GB/s\n", rows * dense / a / 1e9, rows * dense / b / 1e9);
    }
    // two streams, two halves (both copy engines?)
    cudaStream_t s2; cudaStreamCreate(&s2);
    auto t0 = std::chrono::steady_clock::now();
    cudaMemcpyAsync(h, d, rows * dense / 2, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync((char *)h + rows * dense / 2, (char *)d + rows * dense / 2, rows * dense / 2, cudaMemcpyDeviceToHost, s2);
    cudaStreamSynchronize(st); cudaStreamSynchronize(s2);
    auto t1 = std::chrono::steady_clock::now();
    printf("1D on two streams: %.2f GB/s\n", rows * dense / std::chrono::duration<double>(t1 - t0).count() / 1e9);
    return 0;
}
