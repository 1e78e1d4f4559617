// Microbenchmark: issue throughput of FFMA, FFMA2 (fma.rn.f32x2) and FMNMX / FMNMX3 on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float m = s, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {                       // 16 independent FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        } else if (MODE == 1) {                // 8 independent FFMA2 = 16 FMAs
#pragma unroll
            for (int i = 0; i < 16; i += 2) { float2 r = ffma2(make_float2(a[i], a[i + 1]), make_float2(m, m), make_float2(c, c)); a[i] = r.x; a[i + 1] = r.y; }
        } else if (MODE == 2) {                // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fminf(a[i], m + i);
        } else if (MODE == 3) {                // 8 FMNMX3
#pragma unroll
            for (int i = 0; i < 16; i += 2) a[i] = fminf(fminf(a[i], a[i + 1]), m + i);
        } else {                               // mix: 12 FFMA2 (24 FMAs) + 8 FMNMX3: like the scan
#pragma unroll
            for (int i = 0; i < 16; i += 2) { float2 r = ffma2(make_float2(a[i], a[i + 1]), make_float2(m, m), make_float2(c, c)); a[i] = r.x; a[i + 1] = r.y; }
#pragma unroll
            for (int i = 0; i < 8; i += 2) { float2 r = ffma2(make_float2(a[i], a[i + 1]), make_float2(c, m), make_float2(c, c)); a[i] = r.x; a[i + 1] = r.y; }
#pragma unroll
            for (int i = 0; i < 16; i += 2) a[i] = fminf(fminf(a[i], a[i + 1]), m + i);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double ops_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_inst = (double)148 * 8 * 8 * iters * ops_per_iter;
    printf("%-28s %8.3f ms  %7.2f warp-instr/clk/SM (at 1.965 GHz)  \n", name, ms, warp_inst / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("FFMA x16", 16); run<1>("FFMA2 x8 (16 FMAs)", 8); run<2>("FMNMX x16", 16); run<3>("FMNMX3 x8", 8);
    run<4>("12 FFMA2 + 8 FMNMX3", 20);
    return 0;
}
