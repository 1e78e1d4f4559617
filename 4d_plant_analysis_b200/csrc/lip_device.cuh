// Device-side building blocks shared by the local-ICP kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#ifndef LIP_MAP_SUB
#define LIP_MAP_SUB 4         // ranges per landmark bin with a bounding box of their own (the unit the final map culls and stages)
#endif

#define LIP_FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// Batch layout in HBM.  Meshes are numbered side-major: m in [0,S) = mesh1 of pair m (frame 1),
// m in [S,2S) = mesh2 of pair m-S.  "Unified" vertex index u = position in [xyz1 | xyz2];
// "bin" b = unified landmark index in [lm1 | lm2].  Bin b of mesh m holds the vertices of m whose
// nearest landmark is b; the ICP of bin b has src = bin b and tgt = the same landmark of the
// partner mesh, bin (b + sumK) mod 2 sumK.
// ---------------------------------------------------------------------------------------------
struct GridHdr {
    double c[3];     // centre of the bounding box: fp32 coordinates are taken relative to it
    double sum[3];   // sum of the vertices (mean-shift fallback of the fill step)
    float gmin[3];   // grid origin (bbox min, centred)
    float inv_h, h;  // cell size
    int dim[3];      // cells per axis
    int mask;        // hash table size - 1 (power of two)
    float mabs;      // max |centred coordinate| over the mesh
    int n;
};

// one tile of the final map: up to LIP_MAP_TILE queries of source bin b (written by icp_schedule_kernel)
struct MapTile {
    int b, first, nq, m, bb, K, bbt, so, ub, ts, te;   // source bin, first slot, queries, mesh, its first bin / bin count, first bin of
    float home_mabs;                                    // the target mesh, slot bases, target bin of the same landmark [ts, te)
    double C[3];                                        // tile origin: the target's landmark
};

// frame of one partition ICP: centre = target centroid, tile shift = target landmark - centre, bound on |tile coordinate|,
// source / target ranges in bin order (written per landmark bin by icp_schedule_kernel, read once per pass)
struct __align__(16) IcpFrame { double c0, c1, c2; float sh0, sh1, sh2, mabs; int so, ns, to, nt; int axis; float wb; };   // axis, wb: sort axis of the target bin and the width of one bucket of its counting sort

struct LipArgs {
    int S, M, B, sumK, n1_total, n_total, dir_mask, n_tiles;
    const double *xyz1, *xyz2, *lm1, *lm2;
    const int* mesh_row;     // nullptr, or [M]: first row of each mesh in ONE coordinate array (xyz1); meshes may be shared between
                             // segment pairs (time series: a frame is mesh2 of one frame pair and mesh1 of the next)
    const int* mesh_ubase;   // [M+1] unified vertex base of each mesh
    const int* mesh_bin;     // [M+1] first bin (= unified landmark index) of each mesh
    const int* mesh_tab;     // [M+1] first entry of each mesh's hash table in grid_start
    const int* bin_mesh;     // [B]   mesh of each bin
    const int* tile_mesh;    // [n_tiles] blend/query tiles: mesh and first local vertex
    const int* tile_first;   // [n_tiles]
    const int* ptile_mesh;   // [n_ptiles] tiles of the partition kernels (LIP_PART_CHUNK vertices of one mesh, every mesh): mesh,
    const int* ptile_first;  // [n_ptiles] first local vertex,
    const int* ptile_cnt;    // [n_ptiles + 1] offset of the tile's per-landmark counts in tile_cnt;
    const int* mesh_ptile;   // [M+1] first tile of each mesh
    int* tile_cnt;           // per (tile, landmark of its mesh): members in this tile
    int* tile_base;          // ... and in the earlier tiles of the mesh (partition_scan_kernel)
    float4* bin_tmp;         // [n_total] fp32 offset from the bin's landmark + local vertex id, bins in member (vertex) order
    double* tile_sum;        // [n_ptiles * 4] coordinate sums of the tile's vertices
    int *nn0, *nn1;          // [n_total] nearest / second nearest landmark (local index)
    double *d0, *d1;         // [n_total] their distances
    int* bin_off;            // [B+1] unified start of each bin in part_vid
    int* part_vid;           // [n_total] local vertex ids grouped by bin, ascending inside a bin
    int* nn_idx;             // [n_total] ICP scratch: current nearest-neighbour index of every source point (same layout)
    float4* bin_pts;         // [n_total] vertices in part_vid order as fp32 offsets from their OWN landmark, .w = local vertex id
    double* bin_xyz;         // [3 * n_total] the same vertices in float64, planes x | y | z in part_vid order
    double* bin_q;           // [3 * n_total] blended positions of the source vertices (blend_kernel), same layout
    int *pos_bin, *pos_nn1;  // [n_total] bin and second-nearest landmark of every slot (bin order)
    double* pos_d;           // [2 * n_total] distances to the nearest / second-nearest landmark, planes, bin order
    float4* sub_box;         // [B*2] extents along the bin's sort axis of the LIP_MAP_SUB consecutive ranges of its FINAL order (ceil(n / LIP_MAP_SUB) members each): lo[4], hi[4] in the frame of bin_box; empty range: lo > hi
    float* bin_box;          // [B*8] bounding box of every bin in the same frame: lo xyz, hi xyz, max |coordinate|, sort axis (-1: not sorted)
    MapTile* map_tiles;      // [n_total / LIP_MAP_TILE + B] final-map tiles; count in sched[20]
    double* bin_sum;         // [B*4] float64 mean of the vertices of every bin (x, y, z) and their number
    double* mesh_sum;        // [M*4] sum of the vertices of every mesh (x, y, z) and the vertex count
    double *icp_R, *icp_t;   // [B*9], [B*3]
    int *icp_iters, *icp_status;   // [B]
    double *fill_R, *fill_t;       // [B*9], [B*3] after the fill step
    IcpFrame* icp_frames;          // [B]
    double* bin_desc;              // [B*16] per landmark bin, for blend_kernel: filled R (9), t (3), landmark (3), {first bin of the mesh, K} as two ints (K = 0: mesh not a source)
    int* sched;              // [32] [0] partitions of icp_group_kernel [1] n_big [2] group task cursor [3] n_far
                             //      [6] far cursor [7] n_retry [12..15] big class counts [24..31] group class counts
                             //      [16] big cursor [20] number of final-map tiles [21] their cursor
    int *warp_list, *big_list;     // [8B] (eight work classes), [4B] (four)
    int* far_list;           // [n_total] unified ids of queries whose ring search could not certify
    double* far_q;           // [3 * n_total] their blended positions
    int* retry_list;         // [2 * n_total] (unified id, target mesh) of queries the cooperative pass left open
    double* retry_q;         // [3 * n_total]
    float4* grid_pts;        // [n_total] cell-sorted centred fp32 points, .w = local vertex id
    int* grid_start;         // hash tables, mesh m at mesh_tab[m], size mask+3
    GridHdr* grid_hdr;       // [M]
    long long *T12, *T21;    // outputs (unified order: T12 over xyz1, T21 over xyz2)
    double* q_dist;          // raw-query mode: distance output over xyz2 (or nullptr)
    unsigned long long* counters;  // [8]
    int max_iter, profile;   // profile != 0: count evaluated candidate pairs (counters[7])
    int icp_tasks_per_warp, icp_class_shift;   // group sizes of icp_group_kernel (tuning knobs)
    int cluster_coop_parts;                    // ... whose CTAs split one partition between them only in batches of at most this many partitions
    int big_cluster_cls;                       // CTA work classes >= this one run on thread-block clusters (icp_cluster_kernel); 4: none
    long long warp_work;     // partitions with n_src * n_tgt <= this go to icp_group_kernel, heavier ones to icp_big_kernel
    int map_mode;            // final map: 0 = landmark-bin tiles (blend_tile_kernel), 1 = hashed grid (blend_nn_kernel)
    int map32;               // entry type of T12 / T21: 0 int64, 1 int32, 2 uint16 (lip_set_map_dtype)
    double rise_tol, flat_tol, obtuse_cos, cos_den_eps, reject_dist;
};

// dense-map entry of unified vertex u (mesh1 vertices fill T12, mesh2 vertices T21), int64 like numpy's intp or int32
__device__ __forceinline__ void store_map(const LipArgs& a, int u, int idx) {
    const bool side1 = u < a.n1_total;
    const int i = side1 ? u : u - a.n1_total;
    long long* t = side1 ? a.T12 : a.T21;
    if (a.map32 == 1) reinterpret_cast<int*>(t)[i] = idx;
    else if (a.map32 == 2) reinterpret_cast<unsigned short*>(t)[i] = (unsigned short)idx;
    else t[i] = idx;
}
__device__ __forceinline__ const double* mesh_xyz(const LipArgs& a, int m) {
    if (a.mesh_row) return a.xyz1 + 3 * (size_t)a.mesh_row[m];
    int ub = a.mesh_ubase[m];
    return (m < a.S) ? a.xyz1 + 3 * (size_t)ub : a.xyz2 + 3 * (size_t)(ub - a.n1_total);
}
__device__ __forceinline__ const double* mesh_lm(const LipArgs& a, int m) {
    int lb = a.mesh_bin[m];
    return (m < a.S) ? a.lm1 + 3 * (size_t)lb : a.lm2 + 3 * (size_t)(lb - a.sumK);
}
__device__ __forceinline__ int pair_mesh(const LipArgs& a, int m) { return m < a.S ? m + a.S : m - a.S; }
__device__ __forceinline__ int pair_bin(const LipArgs& a, int b) { return b < a.sumK ? b + a.sumK : b - a.sumK; }
// is mesh m the SOURCE of a requested direction?  T21 (bit 0): src = mesh2 (m >= S)
__device__ __forceinline__ bool src_active(const LipArgs& a, int m) {
    return (m >= a.S) ? (a.dir_mask & 1) : (a.dir_mask & 2);
}

// ---------------------------------------------------------------------------------------------
// warp reductions (xor butterfly: every lane ends with the bit-identical result)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LIP_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(LIP_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(LIP_FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(LIP_FULL, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// exact float64 squared distance in the operation order of sklearn's rdist (x, then y, then z;
// every product and sum rounded separately — no FMA contraction), so that exact ties resolve on
// the same values the reference sees.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double rdist64(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = ax - bx, dy = ay - by, dz = az - bz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// fp32 top-2 tracking: best, second best and index of the best
__device__ __forceinline__ void nn_update(float d, int j, float& best, float& sec, int& idx) {
    sec = fminf(sec, fmaxf(d, best));
    idx = (d < best) ? j : idx;
    best = fminf(best, d);
}
// bound on |d2_fp32 - d2_exact| given the per-component error e of a coordinate difference
__device__ __forceinline__ float d2_err(float d2, float e) {
    return 3.5f * sqrtf(d2) * e + 3.0f * e * e + 5e-7f * d2;
}

// ---------------------------------------------------------------------------------------------
// 5th-generation tensor cores (tcgen05): PTX wrappers used by blend_tile_tc_kernel.  Operands are K-major,
// not swizzled: a core matrix is 8 rows x 16 bytes stored as 128 contiguous bytes; the descriptor holds the
// byte distance between 8-row groups (sbo) and between the two 16-byte K halves of a K = 8 tf32 step (lbo).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a protocol error must end in a trap (reported as a launch failure), never in a hung GPU
__device__ __noinline__ void mbar_timeout() { __trap(); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
#pragma unroll 1
    for (int i = 0; i < (1 << 26); ++i) if (mbar_try_wait(bar, parity)) return;
    mbar_timeout();
}
__device__ __forceinline__ unsigned long long umma_desc(const void* p, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((tc_smem_u32(p) >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr unsigned umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&r)[32]) {
    unsigned* u = reinterpret_cast<unsigned*>(r);
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// upper part of an fp32 that a tf32 operand represents exactly (10 explicit mantissa bits): whatever the tensor core does with
// the low 13 bits (truncate or round) is then a no-op
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---------------------------------------------------------------------------------------------
// 3x3 helpers (row-major)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double* A, const double* Bm, double* C) {  // C = A*B (C may alias B)
    double r[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r[3 * i + j] = A[3 * i] * Bm[j] + A[3 * i + 1] * Bm[3 + j] + A[3 * i + 2] * Bm[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) C[i] = r[i];
}
__device__ __forceinline__ void mat3_vec(const double* A, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = A[0] * x + A[1] * y + A[2] * z;
    oy = A[3] * x + A[4] * y + A[5] * z;
    oz = A[6] * x + A[7] * y + A[8] * z;
}

__host__ __device__ __forceinline__ double rcp_approx(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return (double)(1.0f / (float)x);     // host build (unit test harness): same ~20-bit precision class
#endif
}
__host__ __device__ __forceinline__ double rsqrt_approx(double x) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
#else
    return (double)(1.0f / sqrtf((float)x));
#endif
}
__host__ __device__ __forceinline__ double rsqrt_full(double x) {
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// One-sided Jacobi step on columns p,q of A (and V).  tan(theta) only has to be approximately the
// annihilating one (the sweeps are iterated to convergence), so it comes from the ~20-bit MUFU
// reciprocal / rsqrt; cos and sin are then derived from it in full precision so that every step is
// an exact rotation and V stays orthogonal.
#define LIP_JACOBI_PAIR(AP0, AP1, AP2, AQ0, AQ1, AQ2, VP0, VP1, VP2, VQ0, VQ1, VQ2)              \
    {                                                                                            \
        double al = AP0 * AP0 + AP1 * AP1 + AP2 * AP2;                                           \
        double be = AQ0 * AQ0 + AQ1 * AQ1 + AQ2 * AQ2;                                           \
        double ga = AP0 * AQ0 + AP1 * AQ1 + AP2 * AQ2;                                           \
        if (ga * ga > 1e-30 * al * be) {                                                         \
            double ze = (be - al) * rcp_approx(2.0 * ga);                                        \
            double az = fabs(ze), tt;                                                            \
            if (az > 1e7) tt = 0.5 * rcp_approx(ze);                                             \
            else {                                                                               \
                double w1 = 1.0 + ze * ze;                                                       \
                tt = copysign(rcp_approx(az + w1 * rsqrt_approx(w1)), ze);                       \
            }                                                                                    \
            double cs = rsqrt_full(1.0 + tt * tt), sn = cs * tt, x;                               \
            x = AP0; AP0 = cs * x - sn * AQ0; AQ0 = sn * x + cs * AQ0;                           \
            x = AP1; AP1 = cs * x - sn * AQ1; AQ1 = sn * x + cs * AQ1;                           \
            x = AP2; AP2 = cs * x - sn * AQ2; AQ2 = sn * x + cs * AQ2;                           \
            x = VP0; VP0 = cs * x - sn * VQ0; VQ0 = sn * x + cs * VQ0;                           \
            x = VP1; VP1 = cs * x - sn * VQ1; VQ1 = sn * x + cs * VQ1;                           \
            x = VP2; VP2 = cs * x - sn * VQ2; VQ2 = sn * x + cs * VQ2;                           \
            any = true;                                                                          \
        }                                                                                        \
    }

#define LIP_SWAP_COL(SA, A0, A1, A2, V0, V1, V2, SB, B0, B1, B2, W0, W1, W2)                      \
    if (SA < SB) {                                                                               \
        double x;                                                                                \
        x = SA; SA = SB; SB = x; x = A0; A0 = B0; B0 = x; x = A1; A1 = B1; B1 = x;               \
        x = A2; A2 = B2; B2 = x; x = V0; V0 = W0; W0 = x; x = V1; V1 = W1; W1 = x;               \
        x = V2; V2 = W2; W2 = x;                                                                 \
    }

// Kabsch rotation from the cross-covariance H (row-major, H = sum (a - ca)(b - cb)^T):
// the proper rotation R (det +1) maximising trace(R H), i.e. numpy's V U^T with the row of Vt of the
// smallest singular value flipped when det < 0 (registration_icp.py:30-36).  Only the two leading
// singular pairs are needed; the third follows by cross products, which IS the reflection fix.
// Returns 0, or 1 when H has rank < 2 (R not unique; a minimal rotation u0 -> v0 is returned).
__host__ __device__ __noinline__ int kabsch_rotation(const double* H, double* R) {
    // a_k = column k of H (left singular side), v_k = column k of V
    double a00 = H[0], a01 = H[3], a02 = H[6];
    double a10 = H[1], a11 = H[4], a12 = H[7];
    double a20 = H[2], a21 = H[5], a22 = H[8];
    double v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool any = false;
        LIP_JACOBI_PAIR(a00, a01, a02, a10, a11, a12, v00, v01, v02, v10, v11, v12)
        LIP_JACOBI_PAIR(a00, a01, a02, a20, a21, a22, v00, v01, v02, v20, v21, v22)
        LIP_JACOBI_PAIR(a10, a11, a12, a20, a21, a22, v10, v11, v12, v20, v21, v22)
        if (!any) break;
    }
    // H V = [sigma_k u_k]: squared singular values are the squared column norms
    double s0 = a00 * a00 + a01 * a01 + a02 * a02;
    double s1 = a10 * a10 + a11 * a11 + a12 * a12;
    double s2 = a20 * a20 + a21 * a21 + a22 * a22;
    LIP_SWAP_COL(s0, a00, a01, a02, v00, v01, v02, s1, a10, a11, a12, v10, v11, v12)
    LIP_SWAP_COL(s0, a00, a01, a02, v00, v01, v02, s2, a20, a21, a22, v20, v21, v22)
    LIP_SWAP_COL(s1, a10, a11, a12, v10, v11, v12, s2, a20, a21, a22, v20, v21, v22)
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    if (!(s0 > 1e-280)) return 1;                            // H == 0: identity
    double x;
    x = 1.0 / sqrt(a00 * a00 + a01 * a01 + a02 * a02); double u00 = a00 * x, u01 = a01 * x, u02 = a02 * x;
    x = 1.0 / sqrt(v00 * v00 + v01 * v01 + v02 * v02); v00 *= x; v01 *= x; v02 *= x;
    if (!(s1 > 1e-24 * s0)) {
        // rank 1: minimal rotation taking u0 to v0
        double c = u00 * v00 + u01 * v01 + u02 * v02;
        double wx = u01 * v02 - u02 * v01, wy = u02 * v00 - u00 * v02, wz = u00 * v01 - u01 * v00;
        if (c > -1.0 + 1e-12) {
            double k = 1.0 / (1.0 + c);
            R[0] = 1 - k * (wy * wy + wz * wz); R[1] = -wz + k * wx * wy;       R[2] = wy + k * wx * wz;
            R[3] = wz + k * wx * wy;            R[4] = 1 - k * (wx * wx + wz * wz); R[5] = -wx + k * wy * wz;
            R[6] = -wy + k * wx * wz;           R[7] = wx + k * wy * wz;        R[8] = 1 - k * (wx * wx + wy * wy);
        } else {
            // u0 = -v0: half turn about an axis orthogonal to u0
            double px = (fabs(u00) < 0.9) ? 1.0 : 0.0, py = 1.0 - px, pz = 0.0;
            double ax = u01 * pz - u02 * py, ay = u02 * px - u00 * pz, az = u00 * py - u01 * px;
            double k = 1.0 / sqrt(ax * ax + ay * ay + az * az);
            ax *= k; ay *= k; az *= k;
            R[0] = 2 * ax * ax - 1; R[1] = 2 * ax * ay;     R[2] = 2 * ax * az;
            R[3] = 2 * ay * ax;     R[4] = 2 * ay * ay - 1; R[5] = 2 * ay * az;
            R[6] = 2 * az * ax;     R[7] = 2 * az * ay;     R[8] = 2 * az * az - 1;
        }
        return 1;
    }
    // second pair: re-orthogonalise against the first (the Jacobi columns already are, to rounding)
    x = a10 * u00 + a11 * u01 + a12 * u02; a10 -= x * u00; a11 -= x * u01; a12 -= x * u02;
    x = 1.0 / sqrt(a10 * a10 + a11 * a11 + a12 * a12); double u10 = a10 * x, u11 = a11 * x, u12 = a12 * x;
    x = v10 * v00 + v11 * v01 + v12 * v02; v10 -= x * v00; v11 -= x * v01; v12 -= x * v02;
    x = 1.0 / sqrt(v10 * v10 + v11 * v11 + v12 * v12); v10 *= x; v11 *= x; v12 *= x;
    double u20 = u01 * u12 - u02 * u11, u21 = u02 * u10 - u00 * u12, u22 = u00 * u11 - u01 * u10;
    v20 = v01 * v12 - v02 * v11; v21 = v02 * v10 - v00 * v12; v22 = v00 * v11 - v01 * v10;
    // R = V U^T : R[i][j] = sum_k v_k[i] u_k[j]
    R[0] = v00 * u00 + v10 * u10 + v20 * u20; R[1] = v00 * u01 + v10 * u11 + v20 * u21; R[2] = v00 * u02 + v10 * u12 + v20 * u22;
    R[3] = v01 * u00 + v11 * u10 + v21 * u20; R[4] = v01 * u01 + v11 * u11 + v21 * u21; R[5] = v01 * u02 + v11 * u12 + v21 * u22;
    R[6] = v02 * u00 + v12 * u10 + v22 * u20; R[7] = v02 * u01 + v12 * u11 + v22 * u21; R[8] = v02 * u02 + v12 * u12 + v22 * u22;
    return 0;
}
