// Microbenchmark: do FFMA (fma pipe) and FMNMX / FMNMX3 / integer ALU ops overlap on sm_100a, or share dispatch?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s, int si) {
    float a[16], b[8];
    int c[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { b[i] = threadIdx.x * 2e-3f + i; c[i] = threadIdx.x + i; }
    const float m = s, cc = 1e-7f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) {       // 16 FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, cc);
        }
        if (MODE == 1 || MODE == 2) {                                  // 8 FMNMX (2-input), values keep changing
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = fminf(b[i], -b[(i + 1) & 7]);
        }
        if (MODE == 3 || MODE == 5) {                                  // 8 FMNMX3
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = fminf(fminf(b[i], -b[(i + 1) & 7]), -b[(i + 3) & 7]);
        }
        if (MODE == 4 || MODE == 6) {                                  // 8 integer LOP3/IADD3
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = (c[i] ^ c[(i + 1) & 7]) + si;
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += b[i] + c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, 0.999f, 3);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.999f, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0;
    printf("%-34s %8.3f ms   %6.1f cycles per loop iteration per warp (16 warps / scheduler)\n", name, ms, cyc);
    cudaFree(out);
}
int main() {
    run<0>("16 FFMA"); run<1>("8 FMNMX"); run<2>("16 FFMA + 8 FMNMX"); run<5>("8 FMNMX3"); run<3>("16 FFMA + 8 FMNMX3");
    run<6>("8 LOP3+IADD3"); run<4>("16 FFMA + 8 (LOP3+IADD3)");
    return 0;
}
