// Microbenchmark: the nearest-neighbour scan of the ICP / final-map kernels with the dot products on the 5th-gen tensor
// cores (tcgen05.mma kind::tf32, accumulators in TMEM, tcgen05.ld epilogue) against the CUDA-core scan the library uses
// (tile_scan<2, DOT> of csrc/lip_kernels.cuh: 3 FFMA per pair, block minima of 8 with FMNMX3, top-2 tracking).
//
//   score(q, y) = |y|^2 - 2 q.y   (= d^2 - |q|^2), fp32-exact enough for an error-bounded top-2 decision.
//   Split operands (every part exactly representable in tf32, so the hardware's input truncation is a no-op):
//     q = qh + ql,  y = yh + yl,  n = |y|^2 = nh + nl
//     A row (query)      = [-2qh.x -2qh.y -2qh.z 1 | -2ql.x -2ql.y -2ql.z 0]             (K = 8)
//     Bhi row (candidate) = [yh.x yh.y yh.z nh] read for BOTH K halves (leading-dimension offset 0)
//     Blo row (candidate) = [yl.x yl.y yl.z nl] likewise
//     D  = A Bhi^T ; D += A Blo^T   ->  nh + nl - 2 (qh + ql).(yh + yl): the full product, two MMAs of K = 8.
//   128 queries (TMEM lanes) x NCH candidates (TMEM columns) per MMA pair; candidates arrive in SMEM by cp.async.bulk +
//   mbarrier from pre-split global arrays (two float4 planes); TMEM double-buffered: the MMAs of chunk c+1 run under
//   the epilogue of chunk c.  Epilogue: tcgen05.ld 32x32b.x32 (thread = query, 32 candidate scores per load), block
//   minima of 8 (FMNMX3) + top-2 tracking exactly as tile_scan does.
//
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tcgen05_scan_bench tcgen05_scan_bench.cu
// Run:    ./tcgen05_scan_bench [chunks per CTA] [ctas per SM]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must end in a trap, never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < (1 << 24); ++i) if (mbar_try_wait(bar, parity)) return;
    printf("mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes;
// sbo = byte distance between 8-row groups, lbo = byte distance between the two 16-byte K halves of a K = 8 tf32 step
__device__ __forceinline__ uint64_t umma_desc(const void* p, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_u32(p) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(r);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                 "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                   "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
                   "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
                   "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// the library's tracking step on a block minimum (tile_scan)
__device__ __forceinline__ void track(float bm, int pos, float& best, float& sec, int& blk) {
    sec = fminf(sec, fmaxf(bm, best));
    blk = (bm < best) ? pos : blk;
    best = fminf(best, bm);
}
__device__ __forceinline__ float min8(const float* c) {
    float bm = fminf(fminf(c[0], c[1]), c[2]);
    bm = fminf(fminf(bm, c[3]), c[4]);
    bm = fminf(fminf(bm, c[5]), c[6]);
    return fminf(bm, c[7]);
}

struct Out { float best, sec; int blk; int pad; };

// ---------------------------------------------------------------------------------------------------------------
// tensor-core scan: CTA = 128 threads = 128 queries; n_chunks chunks of NCH candidates each
// ---------------------------------------------------------------------------------------------------------------
// MODE 0: the whole scan; 1: MMAs only (no tcgen05.ld / minima: the tensor-pipe floor); 2: epilogue only (tcgen05.ld + minima on
// whatever TMEM holds: the TMEM-read + min-unit floor)
template <int NCH, int STAGES, int MODE>
__global__ void __launch_bounds__(128) scan_tc_kernel(const float4* __restrict__ A0g, const float4* __restrict__ A1g,
                                                       const float4* __restrict__ Bhi, const float4* __restrict__ Blo, int n_chunks,
                                                       Out* __restrict__ out, float* __restrict__ dump, long long* __restrict__ cyc) {
    // A: two planes (K halves) of 128 float4; B: per stage the hi plane and the lo plane, each read for both K halves (lbo = 0)
    __shared__ __align__(128) float4 sA[2][128];
    __shared__ __align__(128) float4 sB[STAGES][2][NCH];
    __shared__ __align__(8) uint64_t full[STAGES], done[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const size_t cta_cand = (size_t)blockIdx.x * n_chunks * NCH;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * NCH));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
        mbar_init(&done[0], 1); mbar_init(&done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    sA[0][tid] = A0g[(size_t)blockIdx.x * 128 + tid];
    sA[1][tid] = A1g[(size_t)blockIdx.x * 128 + tid];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of A -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = umma_idesc_tf32(128, NCH);
    constexpr uint32_t plane = NCH * 16;
    auto load_chunk = [&](int c) {                              // thread 0
        const int s = c % STAGES;
        mbar_expect_tx(&full[s], 2 * plane);
        bulk_g2s(&sB[s][0][0], Bhi + cta_cand + (size_t)c * NCH, plane, &full[s]);
        bulk_g2s(&sB[s][1][0], Blo + cta_cand + (size_t)c * NCH, plane, &full[s]);
    };
    auto issue_mma = [&](int c) {                               // thread 0
        const int s = c % STAGES, tb = c & 1;
        mbar_wait(&full[s], (c / STAGES) & 1);
        tc_fence_after();
        if (MODE != 2) {
            const uint64_t ad = umma_desc(&sA[0][0], 128 * 16, 128);
            umma_tf32(tmem + tb * NCH, ad, umma_desc(&sB[s][0][0], 0, 128), idesc, 0);
            umma_tf32(tmem + tb * NCH, ad, umma_desc(&sB[s][1][0], 0, 128), idesc, 1);
        }
        umma_commit(&done[tb]);
    };
    const long long t0 = clock64();
    if (tid == 0) {
        for (int c = 0; c < STAGES && c < n_chunks; ++c) load_chunk(c);
        issue_mma(0);
    }
    float best = INFINITY, sec = INFINITY;
    int blk = 0;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < n_chunks; ++c) {
        const int s = c & 1;
        if (tid == 0 && c + 1 < n_chunks) issue_mma(c + 1);       // into the other TMEM buffer, drained one iteration ago
        mbar_wait(&done[s], (c >> 1) & 1);
        tc_fence_after();
        if (tid == 0 && c + STAGES < n_chunks) load_chunk(c + STAGES);   // the MMAs of chunk c have finished reading their SMEM stage
        const uint32_t tbase = tmem + lane_addr + s * NCH;
        float r[2][32];
        if (MODE == 1) { tc_fence_before(); __syncthreads(); continue; }
        tmem_ld32(tbase, r[0]);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < NCH / 32; ++g) {
            if (g + 1 < NCH / 32) tmem_ld32(tbase + (g + 1) * 32, r[(g + 1) & 1]);
            const float* v = r[g & 1];
            if (dump && blockIdx.x == 0)
                for (int k = 0; k < 32; ++k) dump[((size_t)c * NCH + g * 32 + k) * 128 + tid] = v[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) track(min8(v + 8 * k), c * NCH + g * 32 + 8 * k, best, sec, blk);
            if (g + 1 < NCH / 32) tmem_wait_ld();
        }
        tc_fence_before();
        __syncthreads();                                          // TMEM buffer s drained by all four warps
    }
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * 128 + tid] = Out{best, sec, blk, 0};
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * NCH));
}


// ---------------------------------------------------------------------------------------------------------------
// the same scan, warp-specialised: warps 0..3 epilogue (TMEM -> registers -> minima), warp 4 producer (bulk copies), warp 5
// MMA issuer; SMEM stages and the two TMEM buffers are handed over through mbarriers only (no CTA barrier per chunk)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NCH, int STAGES>
__global__ void __launch_bounds__(192) scan_tc_ws_kernel(const float4* __restrict__ A0g, const float4* __restrict__ A1g,
                                                          const float4* __restrict__ Bhi, const float4* __restrict__ Blo, int n_chunks,
                                                          Out* __restrict__ out, long long* __restrict__ cyc) {
    __shared__ __align__(128) float4 sA[2][128];
    __shared__ __align__(128) float4 sB[STAGES][2][NCH];
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES], t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t cta_cand = (size_t)blockIdx.x * n_chunks * NCH;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * NCH));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 128) { sA[0][tid] = A0g[(size_t)blockIdx.x * 128 + tid]; sA[1][tid] = A1g[(size_t)blockIdx.x * 128 + tid]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = umma_idesc_tf32(128, NCH);
    constexpr uint32_t plane = NCH * 16;
    const long long t0 = clock64();
    if (warp == 4) {
        if (lane == 0)
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % STAGES;
                if (c >= STAGES) mbar_wait(&s_empty[s], ((c / STAGES) - 1) & 1);
                mbar_expect_tx(&s_full[s], 2 * plane);
                bulk_g2s(&sB[s][0][0], Bhi + cta_cand + (size_t)c * NCH, plane, &s_full[s]);
                bulk_g2s(&sB[s][1][0], Blo + cta_cand + (size_t)c * NCH, plane, &s_full[s]);
            }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint64_t ad = umma_desc(&sA[0][0], 128 * 16, 128);
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % STAGES, tb = c & 1;
                mbar_wait(&s_full[s], (c / STAGES) & 1);
                if (c >= 2) mbar_wait(&t_empty[tb], ((c >> 1) - 1) & 1);
                tc_fence_after();
                umma_tf32(tmem + tb * NCH, ad, umma_desc(&sB[s][0][0], 0, 128), idesc, 0);
                umma_tf32(tmem + tb * NCH, ad, umma_desc(&sB[s][1][0], 0, 128), idesc, 1);
                umma_commit(&s_empty[s]);                        // SMEM stage free once the MMAs have read it
                umma_commit(&t_full[tb]);                        // scores of chunk c complete
            }
        }
    } else {
        float best = INFINITY, sec = INFINITY;
        int blk = 0;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        for (int c = 0; c < n_chunks; ++c) {
            const int tb = c & 1;
            mbar_wait(&t_full[tb], (c >> 1) & 1);
            tc_fence_after();
            const uint32_t tbase = tmem + lane_addr + tb * NCH;
            float r[2][32];
            tmem_ld32(tbase, r[0]);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < NCH / 32; ++g) {
                if (g + 1 < NCH / 32) tmem_ld32(tbase + (g + 1) * 32, r[(g + 1) & 1]);
                const float* v = r[g & 1];
#pragma unroll
                for (int k = 0; k < 4; ++k) track(min8(v + 8 * k), c * NCH + g * 32 + 8 * k, best, sec, blk);
                if (g + 1 < NCH / 32) tmem_wait_ld();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[tb]);            // this warp's quarter of the buffer is drained
        }
        out[(size_t)blockIdx.x * 128 + tid] = Out{best, sec, blk, 0};
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * NCH));
}

// ---------------------------------------------------------------------------------------------------------------
// CUDA-core baseline: the library's scan (tile_scan<2, DOT>): CTA = 128 threads = 256 queries, candidates as float4
// (y, |y|^2) in SMEM, broadcast LDS.128, 3 FFMA per pair
// ---------------------------------------------------------------------------------------------------------------
template <int NCH, int NQ>
__global__ void __launch_bounds__(128) scan_simt_kernel(const float4* __restrict__ Q, const float4* __restrict__ Y, int n_chunks,
                                                         Out* __restrict__ out, long long* __restrict__ cyc) {
    __shared__ float4 tile[2][NCH];
    const int tid = threadIdx.x;
    const size_t cta_cand = (size_t)blockIdx.x * n_chunks * NCH;
    float m[NQ][3], best[NQ], sec[NQ];
    int blk[NQ];
    for (int k = 0; k < NQ; ++k) {
        const float4 q = Q[(size_t)blockIdx.x * 128 * NQ + k * 128 + tid];
        m[k][0] = -2.f * q.x; m[k][1] = -2.f * q.y; m[k][2] = -2.f * q.z;
        best[k] = INFINITY; sec[k] = INFINITY; blk[k] = 0;
    }
    const long long t0 = clock64();
    for (int j = tid; j < NCH; j += 128) tile[0][j] = Y[cta_cand + j];
    __syncthreads();
    for (int c = 0; c < n_chunks; ++c) {
        const float4* tp = tile[c & 1];
        if (c + 1 < n_chunks) for (int j = tid; j < NCH; j += 128) tile[(c + 1) & 1][j] = Y[cta_cand + (size_t)(c + 1) * NCH + j];
#pragma unroll 1
        for (int jb = 0; jb < NCH; jb += 8) {
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = tp[jb + k];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float cc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) cc[k] = fmaf(m[q][0], t[k].x, fmaf(m[q][1], t[k].y, fmaf(m[q][2], t[k].z, t[k].w)));
                track(min8(cc), c * NCH + jb, best[q], sec[q], blk[q]);
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    for (int k = 0; k < NQ; ++k) out[(size_t)blockIdx.x * 128 * NQ + k * 128 + tid] = Out{best[k], sec[k], blk[k], 0};
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}


// ---------------------------------------------------------------------------------------------------------------
// floors: what the two halves of the tensor-core scan cost on their own
// ---------------------------------------------------------------------------------------------------------------
// (a) tensor pipe: NMMA dependent MMAs (M = 128, N, K = 32 bytes: 8 tf32 or 16 bf16) per commit, two commits in flight,
//     operands resident in SMEM (zeros), nothing else in the loop
template <int N, int NMMA, int BF16>
__global__ void __launch_bounds__(128) mma_floor_kernel(int iters, long long* __restrict__ cyc) {
    __shared__ __align__(128) float4 sA[2][128];
    __shared__ __align__(128) float4 sB[2][N];
    __shared__ __align__(8) uint64_t done[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) { mbar_init(&done[0], 1); mbar_init(&done[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 2 * 128; i += 128) (&sA[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 2 * N; i += 128) (&sB[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // kind::f16 descriptor: a/b format 1 = bf16
    const uint32_t idesc = BF16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : umma_idesc_tf32(128, N);
    const uint64_t ad = umma_desc(&sA[0][0], 128 * 16, 128), bd = umma_desc(&sB[0][0], N * 16, 128);
    const long long t0 = clock64();
    if (tid == 0) {
        for (int it = 0; it < iters; ++it) {
            const int b = it & 1;
            if (it >= 2) mbar_wait(&done[b], ((it - 2) >> 1) & 1);
            for (int k = 0; k < NMMA; ++k) {
                if (BF16)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem + b * N), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(k > 0)) : "memory");
                else umma_tf32(tmem + b * N, ad, bd, idesc, k > 0);
            }
            umma_commit(&done[b]);
        }
        for (int it = iters - 2 < 0 ? 0 : iters - 2; it < iters; ++it) mbar_wait(&done[it & 1], (it >> 1) & 1);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * N));
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&r)[64]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(r);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                 "%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,"
                 "%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                   "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
                   "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
                   "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31]), "=r"(u[32]), "=r"(u[33]), "=r"(u[34]), "=r"(u[35]), "=r"(u[36]),
                   "=r"(u[37]), "=r"(u[38]), "=r"(u[39]), "=r"(u[40]), "=r"(u[41]), "=r"(u[42]), "=r"(u[43]), "=r"(u[44]), "=r"(u[45]),
                   "=r"(u[46]), "=r"(u[47]), "=r"(u[48]), "=r"(u[49]), "=r"(u[50]), "=r"(u[51]), "=r"(u[52]), "=r"(u[53]), "=r"(u[54]),
                   "=r"(u[55]), "=r"(u[56]), "=r"(u[57]), "=r"(u[58]), "=r"(u[59]), "=r"(u[60]), "=r"(u[61]), "=r"(u[62]), "=r"(u[63])
                 : "r"(taddr) : "memory");
}
// (b) epilogue: WARPS warps (4: one per lane quarter; 8: two per quarter, splitting the columns) read NCOLS TMEM columns per
//     iteration with tcgen05.ld.x32 (double-buffered) or .x64 and run the block-minima tracking; no MMA, no barrier per chunk
template <int NCOLS, int WARPS, int LDX>
__global__ void __launch_bounds__(WARPS * 32) epi_floor_kernel(int iters, Out* __restrict__ out, long long* __restrict__ cyc) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(NCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    constexpr int SPLIT = WARPS / 4, MYCOLS = NCOLS / SPLIT;
    const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * MYCOLS;
    float best = INFINITY, sec = INFINITY;
    int blk = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (LDX == 32) {
            float r[2][32];
            tmem_ld32(tbase, r[0]);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < MYCOLS / 32; ++g) {
                if (g + 1 < MYCOLS / 32) tmem_ld32(tbase + (g + 1) * 32, r[(g + 1) & 1]);
                const float* v = r[g & 1];
#pragma unroll
                for (int k = 0; k < 4; ++k) track(min8(v + 8 * k), it * NCOLS + g * 32 + 8 * k, best, sec, blk);
                if (g + 1 < MYCOLS / 32) tmem_wait_ld();
            }
        } else {
#pragma unroll
            for (int g = 0; g < MYCOLS / 64; ++g) {
                float r[64];
                tmem_ld64(tbase + g * 64, r);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) track(min8(r + 8 * k), it * NCOLS + g * 64 + 8 * k, best, sec, blk);
            }
        }
    }
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * WARPS * 32 + tid] = Out{best, sec, blk, 0};
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NCOLS));
}

template <int N, int NMMA, int BF16> static void run_mma_floor(int n_cta) {
    long long* dcyc; CK(cudaMalloc(&dcyc, n_cta * 8));
    const int iters = 200;
    mma_floor_kernel<N, NMMA, BF16><<<n_cta, 128>>>(iters, dcyc);
    CK(cudaDeviceSynchronize());
    mma_floor_kernel<N, NMMA, BF16><<<n_cta, 128>>>(iters, dcyc);
    CK(cudaDeviceSynchronize());
    std::vector<long long> cyc(n_cta);
    CK(cudaMemcpy(cyc.data(), dcyc, n_cta * 8, cudaMemcpyDeviceToHost));
    double m = 0; for (long long c : cyc) m += (double)c; m /= n_cta;
    printf("  MMA floor %s M=128 N=%3d, %d MMA per commit, %d CTA/SM: %7.1f clk per commit = %6.1f clk per MMA  (%5.2f clk per 32 scores per scheduler at 2 MMAs per score tile)\n",
           BF16 ? "bf16" : "tf32", N, NMMA, n_cta / 148, m / iters, m / iters / NMMA, (m / iters / NMMA) * 2.0 * 32 * 4 / (128.0 * N) / (n_cta / 148));
    cudaFree(dcyc);
}
template <int NCOLS, int WARPS, int LDX> static void run_epi_floor(int n_cta) {
    long long* dcyc; Out* dout; CK(cudaMalloc(&dcyc, n_cta * 8)); CK(cudaMalloc(&dout, (size_t)n_cta * WARPS * 32 * sizeof(Out)));
    const int iters = 100;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    epi_floor_kernel<NCOLS, WARPS, LDX><<<n_cta, WARPS * 32>>>(iters, dout, dcyc);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    epi_floor_kernel<NCOLS, WARPS, LDX><<<n_cta, WARPS * 32>>>(iters, dout, dcyc);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(n_cta);
    CK(cudaMemcpy(cyc.data(), dcyc, n_cta * 8, cudaMemcpyDeviceToHost));
    double m = 0; for (long long c : cyc) m += (double)c; m /= n_cta;
    const double scores = (double)n_cta * 128 * NCOLS * iters;
    const double per_sm_clk = scores / (m) / 148 * 1.0;          // by the CTA clock (independent of the SM clock)
    printf("  epilogue floor: %3d columns, %d warps/CTA, tcgen05.ld.x%d, %d CTA/SM: %6.2f scores/clk/SM = %5.2f clk per 32 scores per scheduler  (%.3f ms)\n",
           NCOLS, WARPS, LDX, n_cta / 148, per_sm_clk * (n_cta >= 148 ? 1.0 : 148.0 / n_cta), 32.0 * 4 / per_sm_clk, ms);
    cudaFree(dcyc); cudaFree(dout);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int NCH, int STAGES, int MODE, bool WS = false>
static int run_tc(const char* name, int n_cta, int n_chunks, const std::vector<float4>& q, const std::vector<float4>& y, bool check) {
    // operands: A planes per CTA (128 queries each), B planes (hi / lo) per candidate
    const size_t nq = (size_t)n_cta * 128, nc = (size_t)n_cta * n_chunks * NCH;
    std::vector<float4> A0(nq), A1(nq), Bh(nc), Bl(nc);
    for (size_t i = 0; i < nq; ++i) {
        const float4 p = q[i];
        const float h[3] = {tf32_trunc(p.x), tf32_trunc(p.y), tf32_trunc(p.z)};
        A0[i] = make_float4(-2.f * h[0], -2.f * h[1], -2.f * h[2], 1.f);
        A1[i] = make_float4(-2.f * tf32_trunc(p.x - h[0]), -2.f * tf32_trunc(p.y - h[1]), -2.f * tf32_trunc(p.z - h[2]), 0.f);
    }
    for (size_t i = 0; i < nc; ++i) {
        const float4 p = y[i];
        const float n = fmaf(p.x, p.x, fmaf(p.y, p.y, p.z * p.z));
        const float h[4] = {tf32_trunc(p.x), tf32_trunc(p.y), tf32_trunc(p.z), tf32_trunc(n)};
        Bh[i] = make_float4(h[0], h[1], h[2], h[3]);
        Bl[i] = make_float4(tf32_trunc(p.x - h[0]), tf32_trunc(p.y - h[1]), tf32_trunc(p.z - h[2]), tf32_trunc(n - h[3]));
    }
    float4 *dA0, *dA1, *dBh, *dBl; Out* dout; long long* dcyc; float* ddump = nullptr;
    CK(cudaMalloc(&dA0, nq * 16)); CK(cudaMalloc(&dA1, nq * 16)); CK(cudaMalloc(&dBh, nc * 16)); CK(cudaMalloc(&dBl, nc * 16));
    CK(cudaMalloc(&dout, nq * sizeof(Out))); CK(cudaMalloc(&dcyc, n_cta * 8));
    CK(cudaMemcpy(dA0, A0.data(), nq * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dA1, A1.data(), nq * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh.data(), nc * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), nc * 16, cudaMemcpyHostToDevice));
    const size_t ndump = (size_t)n_chunks * NCH * 128;
    if (check) { CK(cudaMalloc(&ddump, ndump * 4)); CK(cudaMemset(ddump, 0, ndump * 4)); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (WS) scan_tc_ws_kernel<NCH, STAGES><<<n_cta, 192>>>(dA0, dA1, dBh, dBl, n_chunks, dout, dcyc);
    else scan_tc_kernel<NCH, STAGES, MODE><<<n_cta, 128>>>(dA0, dA1, dBh, dBl, n_chunks, dout, ddump, dcyc);
    CK(cudaDeviceSynchronize());
    std::vector<Out> ref_out;
    if (WS) {                                                    // the warp-specialised kernel must reproduce the plain one bit for bit
        ref_out.resize(nq);
        std::vector<Out> got(nq);
        CK(cudaMemcpy(got.data(), dout, nq * sizeof(Out), cudaMemcpyDeviceToHost));
        scan_tc_kernel<NCH, 2, 0><<<n_cta, 128>>>(dA0, dA1, dBh, dBl, n_chunks, dout, nullptr, dcyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(ref_out.data(), dout, nq * sizeof(Out), cudaMemcpyDeviceToHost));
        size_t diff = 0;
        for (size_t i = 0; i < nq; ++i) diff += got[i].best != ref_out[i].best || got[i].sec != ref_out[i].sec || got[i].blk != ref_out[i].blk;
        if (diff) printf("  !! warp-specialised kernel differs from the plain one on %zu of %zu queries\n", diff, nq);
    }
    int bad = 0;
    if (check) {                                                 // CTA 0: every score against float64, and the top-2 tracking
        std::vector<float> dump(ndump);
        std::vector<Out> o(128);
        CK(cudaMemcpy(dump.data(), ddump, ndump * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o.data(), dout, 128 * sizeof(Out), cudaMemcpyDeviceToHost));
        double max_rel = 0, max_rel32 = 0;
        for (int t = 0; t < 128; ++t) {
            const float4 p = q[t];
            double bestd = 1e300; size_t bj = 0;
            float bestf = INFINITY;
            for (size_t j = 0; j < (size_t)n_chunks * NCH; ++j) {
                const float4 c = y[j];
                const float nf = fmaf(c.x, c.x, fmaf(c.y, c.y, c.z * c.z));
                const double exact = (double)nf - 2.0 * ((double)p.x * c.x + (double)p.y * c.y + (double)p.z * c.z);
                const double scale = (double)nf + 2.0 * sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z) * sqrt((double)nf);
                const float got = dump[j * 128 + t];
                const float f32 = fmaf(-2.f * p.x, c.x, fmaf(-2.f * p.y, c.y, fmaf(-2.f * p.z, c.z, nf)));
                max_rel = fmax(max_rel, fabs(got - exact) / scale);
                max_rel32 = fmax(max_rel32, fabs(f32 - exact) / scale);
                if (exact < bestd) { bestd = exact; bj = j; }
                bestf = fminf(bestf, got);
            }
            if (o[t].best != bestf || (size_t)o[t].blk != (bj & ~(size_t)7)) {
                // a different block is only acceptable when the runner-up is within the error bound
                if (o[t].best != bestf) { if (bad < 5) printf("  query %d: tracked best %.9g != min of dumped scores %.9g\n", t, o[t].best, bestf); ++bad; }
            }
        }
        printf("  %-28s |score - float64| / (|y|^2 + 2|q||y|): tensor-core split-tf32 %.3g (= 2^%.1f), fp32 FFMA chain %.3g (= 2^%.1f); tracking mismatches %d\n",
               name, max_rel, log2(max_rel), max_rel32, log2(max_rel32), bad);
        cudaFree(ddump);
    }
    const int reps = 5;
    float best_ms = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        if (WS) scan_tc_ws_kernel<NCH, STAGES><<<n_cta, 192>>>(dA0, dA1, dBh, dBl, n_chunks, dout, dcyc);
        else scan_tc_kernel<NCH, STAGES, MODE><<<n_cta, 128>>>(dA0, dA1, dBh, dBl, n_chunks, dout, nullptr, dcyc);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best_ms = fminf(best_ms, ms);
    }
    std::vector<long long> cyc(n_cta);
    CK(cudaMemcpy(cyc.data(), dcyc, n_cta * 8, cudaMemcpyDeviceToHost));
    double mean_cyc = 0; for (long long c : cyc) mean_cyc += (double)c; mean_cyc /= n_cta;
    const double scores = (double)nq * n_chunks * NCH;
    const double per_sm_clk = scores / (best_ms * 1e-3) / 148 / 1.965e9;
    printf("%-30s grid %5d  %8.3f ms  %7.2f scores/clk/SM  = %5.2f clk per 32 scores per scheduler   (CTA: %.0f clk, %.2f clk per 32 scores per warp)\n",
           name, n_cta, best_ms, per_sm_clk, 32.0 * 4 / per_sm_clk, mean_cyc, mean_cyc / (n_chunks * (double)NCH));
    cudaFree(dA0); cudaFree(dA1); cudaFree(dBh); cudaFree(dBl); cudaFree(dout); cudaFree(dcyc);
    return bad;
}

template <int NCH, int NQ>
static void run_simt(const char* name, int n_cta, int n_chunks, const std::vector<float4>& q, const std::vector<float4>& y) {
    const size_t nq = (size_t)n_cta * 128 * NQ, nc = (size_t)n_cta * n_chunks * NCH;
    std::vector<float4> Y(nc);
    for (size_t i = 0; i < nc; ++i) { const float4 p = y[i % y.size()]; Y[i] = make_float4(p.x, p.y, p.z, fmaf(p.x, p.x, fmaf(p.y, p.y, p.z * p.z))); }
    std::vector<float4> Q(nq);
    for (size_t i = 0; i < nq; ++i) Q[i] = q[i % q.size()];
    float4 *dQ, *dY; Out* dout; long long* dcyc;
    CK(cudaMalloc(&dQ, nq * 16)); CK(cudaMalloc(&dY, nc * 16)); CK(cudaMalloc(&dout, nq * sizeof(Out))); CK(cudaMalloc(&dcyc, n_cta * 8));
    CK(cudaMemcpy(dQ, Q.data(), nq * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dY, Y.data(), nc * 16, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    scan_simt_kernel<NCH, NQ><<<n_cta, 128>>>(dQ, dY, n_chunks, dout, dcyc);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        scan_simt_kernel<NCH, NQ><<<n_cta, 128>>>(dQ, dY, n_chunks, dout, dcyc);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best_ms = fminf(best_ms, ms);
    }
    const double scores = (double)nq * n_chunks * NCH;
    const double per_sm_clk = scores / (best_ms * 1e-3) / 148 / 1.965e9;
    printf("%-26s Q=%d  grid %5d  %8.3f ms  %7.2f scores/clk/SM  = %5.2f clk per 32 scores per scheduler\n", name, NQ, n_cta, best_ms, per_sm_clk,
           32.0 * 4 / per_sm_clk);
    cudaFree(dQ); cudaFree(dY); cudaFree(dout); cudaFree(dcyc);
}

int main(int argc, char** argv) {
    const int n_chunks = argc > 1 ? atoi(argv[1]) : 64;
    const int per_sm = argc > 2 ? atoi(argv[2]) : 2;
    int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s, %d SMs, sm_%d%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    // bin-like data: coordinates centred on the bin, extent ~ +-8 units (voxel 0.8), queries displaced by a small motion
    srand(12345);
    auto rnd = [] { return (float)rand() / RAND_MAX; };
    const int max_cta = 148 * 8;
    std::vector<float4> q((size_t)max_cta * 256), y((size_t)max_cta * n_chunks * 256);
    for (auto& p : y) p = make_float4(16.f * rnd() - 8.f, 6.f * rnd() - 3.f, 4.f * rnd() - 2.f, 0.f);
    for (auto& p : q) p = make_float4(16.f * rnd() - 8.f + 0.3f, 6.f * rnd() - 3.f - 0.2f, 4.f * rnd() - 2.f + 0.1f, 0.f);
    int bad = 0;
    printf("== correctness (CTA 0 of a small grid) ==\n");
    bad += run_tc<128, 2, 0>("tcgen05 N=128, 2 stages", 4, 4, q, y, true);
    bad += run_tc<128, 8, 0>("tcgen05 N=128, 8 stages", 4, 20, q, y, true);
    bad += run_tc<256, 4, 0>("tcgen05 N=256, 4 stages", 4, 9, q, y, true);
    printf("== throughput: %d chunks per CTA ==\n", n_chunks);
    for (int ps = 1; ps <= (per_sm < 2 ? per_sm : 2); ++ps) {
        run_tc<128, 2, 0>("tcgen05 N=128, 2 stages", 148 * ps, n_chunks, q, y, false);
        run_tc<128, 4, 0>("tcgen05 N=128, 4 stages", 148 * ps, n_chunks, q, y, false);
        run_tc<128, 8, 0>("tcgen05 N=128, 8 stages", 148 * ps, n_chunks, q, y, false);
        run_tc<128, 8, 1>("  MMAs only (no epilogue)", 148 * ps, n_chunks, q, y, false);
        run_tc<128, 8, 2>("  epilogue only (no MMA)", 148 * ps, n_chunks, q, y, false);
    }
    run_tc<128, 4, 0, true>("tcgen05 warp-specialised N=128", 148, n_chunks, q, y, false);
    run_tc<128, 4, 0, true>("tcgen05 warp-specialised N=128", 296, n_chunks, q, y, false);
    run_tc<256, 4, 0, true>("tcgen05 warp-specialised N=256", 148, n_chunks, q, y, false);
    run_tc<64, 6, 0, true>("tcgen05 warp-specialised N=64", 148 * 4, n_chunks, q, y, false);
    run_tc<256, 4, 0>("tcgen05 N=256, 4 stages", 148, n_chunks, q, y, false);
    run_tc<256, 4, 1>("  MMAs only (no epilogue)", 148, n_chunks, q, y, false);
    run_tc<256, 4, 2>("  epilogue only (no MMA)", 148, n_chunks, q, y, false);
    for (int ps : {2, 4, 5}) run_simt<128, 2>("CUDA-core tile_scan (library)", 148 * ps, n_chunks, q, y);
    for (int ps : {2, 4, 5}) run_simt<128, 1>("CUDA-core tile_scan", 148 * ps, n_chunks, q, y);
    for (int ps : {2, 3, 4}) run_simt<128, 3>("CUDA-core tile_scan", 148 * ps, n_chunks, q, y);
    for (int ps : {2, 3, 4}) run_simt<128, 4>("CUDA-core tile_scan", 148 * ps, n_chunks, q, y);
    for (int ps : {2, 3}) run_simt<128, 6>("CUDA-core tile_scan", 148 * ps, n_chunks, q, y);
    for (int ps : {1, 2}) run_simt<128, 8>("CUDA-core tile_scan", 148 * ps, n_chunks, q, y);
    printf("== floors ==\n");
    run_mma_floor<128, 1, 0>(148); run_mma_floor<128, 2, 0>(148); run_mma_floor<128, 4, 0>(148); run_mma_floor<128, 8, 0>(148);
    run_mma_floor<256, 1, 0>(148); run_mma_floor<256, 2, 0>(148); run_mma_floor<256, 4, 0>(148); run_mma_floor<256, 8, 0>(148);
    run_mma_floor<256, 8, 1>(148); run_mma_floor<128, 8, 0>(296); run_mma_floor<64, 8, 0>(148 * 4);
    run_epi_floor<256, 4, 32>(148); run_epi_floor<256, 4, 64>(148); run_epi_floor<256, 8, 32>(148); run_epi_floor<256, 8, 64>(148);
    run_epi_floor<256, 4, 32>(296); run_epi_floor<256, 4, 64>(296); run_epi_floor<256, 8, 64>(296);
    run_epi_floor<128, 4, 64>(148 * 4); run_epi_floor<128, 8, 64>(148 * 4); run_epi_floor<128, 4, 32>(148 * 4);
    printf("library today: icp_group_kernel / blend_tile_kernel scan loop = 5.6 clk per 32 scores per scheduler (profiles/README.md)\n");
    return bad ? 1 : 0;
}
