// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) issue and pipe throughput on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>   // 0: 16 scalar FFMA per trip; 1: 8 FFMA2 per trip (same flops); 2: 8 FFMA2 + 8 scalar FFMA interleaved
__global__ void k(float* out, int iters, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                p[i] = fma2(p[i], ss, ss);
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, double flops_per_trip, int warps_per_sm) {
    int sms = 148;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 2048);
    const int iters = 200000;
    dim3 grid(sms), block(warps_per_sm * 32);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(out, 1000, 0.999f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double threads = (double)sms * warps_per_sm * 32;
    double tf = threads * iters * flops_per_trip / (ms * 1e-3) / 1e12;
    printf("%-28s warps/SM %2d  %8.3f ms  %7.2f TFLOP/s  (%.2f FMA-lanes per clk per SM at 1.9 GHz)\n", name, warps_per_sm, ms, tf,
           tf * 1e12 / 2 / sms / 1.9e9);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("16 x FFMA", 32, w);
        run<1>("8 x FFMA2", 32, w);
        run<2>("8 x FFMA2 + 8 x FFMA", 48, w);
    }
    return 0;
}
