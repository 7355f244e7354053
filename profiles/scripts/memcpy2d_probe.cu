// probe: D2H bandwidth of cudaMemcpy2DAsync with short rows vs 1-D copies (pinned host memory)
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const size_t rows = 1 << 20, pitch = 359 * 8, width = 108 * 8;
    char *d, *h;
    cudaMalloc(&d, rows * pitch);
    cudaHostAlloc(&h, rows * pitch, cudaHostAllocDefault);
    cudaMemset(d, 1, rows * pitch);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        cudaMemcpy2DAsync(h, pitch, d, pitch, width, rows, cudaMemcpyDeviceToHost, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("2D  D2H %zu rows x %zu B (pitch %zu): %.2f ms  %.1f GB/s\n", rows, width, pitch, ms, rows * width / ms * 1e-6);
        cudaEventRecord(a);
        cudaMemcpyAsync(h, d, rows * width, cudaMemcpyDeviceToHost, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("1D  D2H same bytes: %.2f ms  %.1f GB/s\n", ms, rows * width / ms * 1e-6);
        cudaEventRecord(a);
        cudaMemcpyAsync(h, d, rows * pitch, cudaMemcpyDeviceToHost, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("1D  D2H full rows:  %.2f ms  %.1f GB/s\n", ms, rows * pitch / ms * 1e-6);
        cudaEventRecord(a);
        cudaMemcpyAsync(d, h, rows * pitch, cudaMemcpyHostToDevice, 0);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("1D  H2D full rows:  %.2f ms  %.1f GB/s\n", ms, rows * pitch / ms * 1e-6);
    }
    return 0;
}
