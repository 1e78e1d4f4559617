// Microbenchmark: throughput of the legacy tensor path (mma.sync) on sm_100a for the shapes a split-precision distance
// GEMM would use: m16n8k8 tf32 and m16n8k16 bf16, fp32 accumulate.  Reported per SM and per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float d[8][4];
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u}, b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { if (MODE == 0) mma_tf32(d[i], a, b); else mma_bf16(d[i], a, b); }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double flop_per_mma) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)148 * 8 * 8 * iters * 8;                 // warp-level MMAs
    printf("%-26s %8.3f ms  %6.2f MMA/clk/SM  (%5.2f per scheduler)  %7.1f TFLOP/s  | 16x8 scores per MMA: %6.2f scores/clk/scheduler\n", name, ms,
           mmas / (ms * 1e-3) / 148 / 1.965e9, mmas / (ms * 1e-3) / 148 / 1.965e9 / 4, mmas * flop_per_mma / (ms * 1e-3) / 1e12,
           128.0 * mmas / (ms * 1e-3) / 148 / 1.965e9 / 4);
    cudaFree(out);
}
int main() {
    run<0>("mma.sync m16n8k8 tf32", 2.0 * 16 * 8 * 8);
    run<1>("mma.sync m16n8k16 bf16", 2.0 * 16 * 8 * 16);
    printf("for comparison: the CUDA-core scan delivers 32 scores per ~5.6 clk per scheduler = 5.7 scores/clk/scheduler\n");
    return 0;
}
