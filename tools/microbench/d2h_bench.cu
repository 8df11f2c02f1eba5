// D2H rate of one 4K fp32 plane: contiguous cudaMemcpyAsync against cudaMemcpy2DAsync out of the row-interleaved
// u|v buffer (row of 15360 B every 30720 B), pinned host memory.   nvcc -O2 -o d2h_bench d2h_bench.cu && ./d2h_bench
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const int W = 3840, H = 2160, P = 16, reps = 6;
    const size_t row = (size_t)W * 4, plane = row * H;
    char *d, *h;
    cudaMalloc(&d, 2 * plane * P);
    cudaMallocHost(&h, 2 * plane * P);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int r = 0; r < reps; ++r) {
            cudaEventRecord(a, s);
            for (int p = 0; p < P; ++p) {
                if (mode == 0) cudaMemcpyAsync(h + 2 * plane * p, d + 2 * plane * p, 2 * plane, cudaMemcpyDeviceToHost, s);
                else if (mode == 1) {
                    cudaMemcpy2DAsync(h + 2 * plane * p, row, d + 2 * plane * p, 2 * row, row, H, cudaMemcpyDeviceToHost, s);
                    cudaMemcpy2DAsync(h + 2 * plane * p + plane, row, d + 2 * plane * p + row, 2 * row, row, H, cudaMemcpyDeviceToHost, s);
                } else {
                    cudaMemcpyAsync(h + 2 * plane * p, d + 2 * plane * p, plane, cudaMemcpyDeviceToHost, s);
                    cudaMemcpyAsync(h + 2 * plane * p + plane, d + 2 * plane * p + plane, plane, cudaMemcpyDeviceToHost, s);
                }
            }
            cudaEventRecord(b, s); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        printf("%s: %.2f GB/s\n", mode == 0 ? "1-D, one copy per pair (u|v together)" : mode == 1 ? "2-D, de-interleaving u and v rows" : "1-D, one copy per plane",
               2.0 * plane * P / best / 1e6);
    }
    return 0;
}
