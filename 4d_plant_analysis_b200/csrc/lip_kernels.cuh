// Kernels of the local-ICP hot path (sm_100a).  See DESIGN.md for the data layout and the
// roofline of each kernel.  Reference file:line citations are relative to the reference root.
#pragma once
#include "lip_device.cuh"

#ifndef LIP_WARP_WORK
#define LIP_WARP_WORK 32768         // partitions with n_src * n_tgt <= this: icp_group_kernel (a warp per partition), else CTA ...
#define LIP_WARP_WORK_LARGE 98304   // ... in batches of at least LIP_WARP_WORK_BINS landmark bins this much (measured, profiles/README.md)
#define LIP_WARP_WORK_BINS 12000
#endif
#ifndef LIP_ICP_WARPS
#define LIP_ICP_WARPS 4       // warps per CTA of icp_group_kernel
#endif
#ifndef LIP_ICP_CHUNK
#define LIP_ICP_CHUNK 256     // target points per SMEM tile of a warp
#endif
#ifndef LIP_ICP_SUM17
#define LIP_ICP_SUM17 1       // 1: the 17 float64 sums of a pass are reduced through an SMEM transpose (the dead search tile), 0: 17 butterflies
#endif
#ifndef LIP_ICP_QCAP
#define LIP_ICP_QCAP 512      // queries of a partition whose winner is kept in SMEM (one byte each; the rest go through nn_idx)
#endif
#ifndef LIP_ICP_TASKS_PER_WARP
#define LIP_ICP_TASKS_PER_WARP 1   // group size cap: at least this many tasks per resident warp
#endif
#ifndef LIP_ICP_CLASS_SHIFT
#define LIP_ICP_CLASS_SHIFT 4       // work classes up to this one use groups of LIP_ICP_GMAX, each heavier class half as many
#endif
#ifndef LIP_ICP_GMAX
#define LIP_ICP_GMAX 8        // partitions per warp task: their 3x3 solves run SIMT, one lane per partition
#endif
#ifndef LIP_BIG_CHUNK
#define LIP_BIG_CHUNK 1024    // tgt tile (float4) of the CTA-per-partition variant
#endif
#ifndef LIP_BIG_THREADS
#define LIP_BIG_THREADS 128
#endif
#ifndef LIP_CLUSTER_SIZE
#define LIP_CLUSTER_SIZE 8        // CTAs per cluster of icp_cluster_kernel (runtime: LIP_CLUSTER_SIZE, 1 = off) ...
#define LIP_CLUSTER_CLASS 2       // ... which takes the CTA work classes from this one up (2: > 400 k pairs; runtime: LIP_CLUSTER_CLASS)
#define LIP_CLUSTER_COOP_PARTS 8192   // the cluster kernel runs (and its CTAs co-operate) only in batches of at most this many landmark bins (runtime: LIP_CLUSTER_COOP_PARTS)
#define LIP_CLUSTER_THREADS 256
#define LIP_CLUSTER_MINBLOCKS 2    // CTAs per SM the cluster kernel is compiled for (register cap 128)
#define LIP_CLUSTER_CHUNK 2048    // target points per SMEM tile (also the scratch of the 17-sum reduction: 4 KB per warp)
#endif
#ifndef LIP_ICP_SKIP_MIN
#define LIP_ICP_SKIP_MIN 384   // targets of at least this many points are scanned with the axis block skip (CTA kernels)
#endif
#ifndef LIP_ICP_SKIP
#define LIP_ICP_SKIP 32        // candidates per skip test of the ICP scan (a multiple of 8)
#endif
#define LIP_TILE 128          // vertices per CTA of the blend / query kernels
#ifndef LIP_MAP_THREADS
#define LIP_MAP_THREADS 64   // blend_tile_kernel: threads per CTA
#endif
#ifndef LIP_MAP_Q
#define LIP_MAP_Q 2          // ... and queries per thread.  Measured (profiles/README.md, round 2): 3 / 4 queries per thread (a ~150-point
#endif                       // bin as ONE tile, fewer LDS per score) lose to 2 — 1.24 / 1.34 ms against 1.13: the box cull works per tile
#ifndef LIP_MAP_MINBLOCKS
#define LIP_MAP_MINBLOCKS (LIP_MAP_Q > 2 ? 8 : 12) // CTAs per SM blend_tile_kernel is compiled for
#endif
#define LIP_MAP_TILE (LIP_MAP_Q * LIP_MAP_THREADS)
#ifndef LIP_MAP_CHUNK
#define LIP_MAP_CHUNK 384    // candidates staged per scan (a tile sees ~340 on average; 512 costs 2.5 KB more SMEM per CTA, which the pipelined host path pays in co-residency)
#endif
#define LIP_MAP_BOXBATCH 64
#ifndef LIP_MAP_BOXB
#define LIP_MAP_BOXB 64         // bins per box batch of blend_tile_kernel (32: final map 1.07 -> 1.11 ms on the bench batch)
#endif
#ifndef LIP_WARP_MINBLOCKS
#define LIP_WARP_MINBLOCKS 5   // CTAs per SM icp_group_kernel is compiled for (register cap 96)
#endif

// =============================================================================================
// K0  landmark assignment + stable partition of every mesh by nearest landmark
//     replaces KDTree(lm).query(xyz, k=2) and the boolean masks NN_skel[:,0]==i
//     (p2p_matching_in_organ/p2p_matching_seg_icp.py:388-392, 413-414).
//     Four grid-wide kernels over tiles of LIP_PART_CHUNK vertices of one mesh (no per-mesh CTA, no chain of
//     barrier-separated phases: the parallelism is the vertex count of the batch, not the mesh count):
//       landmark_nn_kernel      two nearest landmarks of every vertex (index order, strict '<' = sklearn's single-leaf scan),
//                               counts per (tile, landmark), coordinate sums per tile
//       partition_scan_kernel   per mesh: bin offsets (exclusive scan over landmarks) and, per (tile, landmark), the number of
//                               members in earlier tiles
//       partition_place_kernel  stable placement (ascending vertex id inside a bin, as the reference's boolean masks give it):
//                               writes every bin-ordered array the later kernels read coalesced
//       bin_stats_kernel        per bin: box of the fp32 offsets, float64 mean
// =============================================================================================
#ifndef LIP_PART_CHUNK
#define LIP_PART_CHUNK 512
#endif
#define LIP_PART_THREADS 256
__global__ void __launch_bounds__(LIP_PART_THREADS) landmark_nn_kernel(LipArgs a) {
    extern __shared__ double smem_d[];
    __shared__ double tsum[LIP_PART_THREADS / 32][3];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = a.ptile_mesh[t], first = a.ptile_first[t];
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
    const double* __restrict__ xyz = mesh_xyz(a, m);
    const double* __restrict__ lm = mesh_lm(a, m);
    double* lmS = smem_d;                                    // landmarks, float64
    float4* lmF = (float4*)(lmS + 3 * K + (K & 1));          // ... and as fp32 offsets from landmark 0 (the prefilter), 16-byte aligned
    int* hist = (int*)(lmF + K);
    constexpr int PER = LIP_PART_CHUNK / LIP_PART_THREADS;
    const int end = min(n, first + LIP_PART_CHUNK);
    double px[PER], py[PER], pz[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {                          // the vertex loads are in flight while the landmarks are staged
        const int v = min(first + r * LIP_PART_THREADS + tid, end - 1);
        px[r] = xyz[3 * (size_t)v]; py[r] = xyz[3 * (size_t)v + 1]; pz[r] = xyz[3 * (size_t)v + 2];
    }
    const double o0 = lm[0], o1 = lm[1], o2 = lm[2];
    for (int i = tid; i < K; i += LIP_PART_THREADS) {
        const double l0 = lm[3 * i], l1 = lm[3 * i + 1], l2 = lm[3 * i + 2];
        lmS[3 * i] = l0; lmS[3 * i + 1] = l1; lmS[3 * i + 2] = l2;
        lmF[i] = make_float4((float)(l0 - o0), (float)(l1 - o1), (float)(l2 - o2), 0.f);
        hist[i] = 0;
    }
    __syncthreads();
    // fp32 prefilter: the three nearest landmarks in fp32 (the tile's vertices of a thread side by side: independent chains,
    // one broadcast LDS.128 per landmark for all of them); if the third is clearly farther than the second (beyond a bound on
    // the fp32 evaluation error) the two nearest are among the three, and only those are evaluated in float64 — in index order
    // with strict '<', which is what the scan over all K gives.  Otherwise (rare) the full float64 scan decides.
    // The landmark index rides in the low mantissa bits of the fp32 score, so that the three smallest (score, index) keys come
    // out of a branch-free min / max network (landmarks lie along the skeleton: the running minimum improves at almost every
    // step up to the nearest one, a compare-and-branch update would diverge all the way).
    float xf[PER], yf[PER], zf[PER], b0[PER], b1[PER], b2[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        xf[r] = (float)(px[r] - o0); yf[r] = (float)(py[r] - o1); zf[r] = (float)(pz[r] - o2);
        b0[r] = b1[r] = b2[r] = INFINITY;
    }
    int ibits = 1;
    while ((1 << ibits) < K) ++ibits;                        // <= 11 (LIP_MAX_LANDMARKS = 2048)
    const unsigned imask = (1u << ibits) - 1u;
    if (K > 3) {
#pragma unroll 2
        for (int i = 0; i < K; ++i) {
            const float4 l = lmF[i];
#pragma unroll
            for (int r = 0; r < PER; ++r) {
                const float dx = xf[r] - l.x, dy = yf[r] - l.y, dz = zf[r] - l.z;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                float key = __uint_as_float((__float_as_uint(d) & ~imask) | (unsigned)i);     // d >= 0: keys order like (d rounded down, i)
                const float m0 = fminf(b0[r], key); key = fmaxf(b0[r], key); b0[r] = m0;
                const float m1 = fminf(b1[r], key); key = fmaxf(b1[r], key); b1[r] = m1;
                b2[r] = fminf(b2[r], key);
            }
        }
    }
    double sx = 0, sy = 0, sz = 0;
#pragma unroll
    for (int r = 0; r < PER; ++r) {
        const int v = first + r * LIP_PART_THREADS + tid;
        if (v >= end) continue;
        const double x = px[r], y = py[r], z = pz[r];
        sx += x; sy += y; sz += z;
        double best = INFINITY, sec = INFINITY;
        int i0 = 0, i1 = 0;
        bool decided = false;
        if (K > 3) {
            // |fp32 distance - true distance| <= e for every landmark at most sqrt(b2) away: input roundings 2^-24 (|q| + |l|) per
            // coordinate with |l| <= |q| + d, three roundings in the sum; 1e-6 (...) is > 8x that
            // the key drops `ibits` mantissa bits of d^2 (relative 2^(ibits-23), i.e. 2^(ibits-24) of the distance)
            const int ja = (int)(__float_as_uint(b0[r]) & imask), jb = (int)(__float_as_uint(b1[r]) & imask), jc = (int)(__float_as_uint(b2[r]) & imask);
            const float s1 = sqrtf(b1[r]), s2 = sqrtf(b2[r]);
            const float e = 1e-6f * (1.f + 2.f * (fabsf(xf[r]) + fabsf(yf[r]) + fabsf(zf[r])) + s2) + s2 * (float)(1u << ibits) * 1.2e-7f;
            if (s2 - s1 > 2.f * e) {
                const int lo = min(ja, min(jb, jc)), hi = max(ja, max(jb, jc)), mid = ja + jb + jc - lo - hi;
                const int cand[3] = {lo, mid, hi};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = cand[k];
                    const double d = rdist64(x, y, z, lmS[3 * i], lmS[3 * i + 1], lmS[3 * i + 2]);
                    if (d < best) { sec = best; i1 = i0; best = d; i0 = i; }
                    else if (d < sec) { sec = d; i1 = i; }
                }
                decided = true;
            }
        }
        if (!decided) {
            best = INFINITY; sec = INFINITY; i0 = 0; i1 = 0;
            for (int i = 0; i < K; ++i) {                   // index order, strict '<': sklearn single-leaf scan
                const double d = rdist64(x, y, z, lmS[3 * i], lmS[3 * i + 1], lmS[3 * i + 2]);
                if (d < best) { sec = best; i1 = i0; best = d; i0 = i; }
                else if (d < sec) { sec = d; i1 = i; }
            }
        }
        if (K == 1) { sec = best; i1 = i0; }
        a.nn0[ub + v] = i0; a.nn1[ub + v] = i1;
        a.d0[ub + v] = sqrt(best); a.d1[ub + v] = sqrt(sec);
        atomicAdd(&hist[i0], 1);                             // integer adds commute: deterministic
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { tsum[warp][0] = sx; tsum[warp][1] = sy; tsum[warp][2] = sz; }
    __syncthreads();
    if (tid < 3) {
        double s = 0;
        for (int w = 0; w < LIP_PART_THREADS / 32; ++w) s += tsum[w][tid];
        a.tile_sum[4 * (size_t)t + tid] = s;
    }
    int* __restrict__ cnt = a.tile_cnt + a.ptile_cnt[t];
    for (int i = tid; i < K; i += LIP_PART_THREADS) cnt[i] = hist[i];
}

// one warp per mesh: bin offsets; per (tile, landmark) the number of members in earlier tiles (tile_base)
__global__ void __launch_bounds__(128) partition_scan_kernel(LipArgs a) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (m >= a.M) return;
    const int t0 = a.mesh_ptile[m], t1 = a.mesh_ptile[m + 1];
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
    const int* __restrict__ cnt = a.tile_cnt + a.ptile_cnt[t0];
    int* __restrict__ tb = a.tile_base + a.ptile_cnt[t0];
    int running = 0;
    for (int base = 0; base < K; base += 32) {
        const int i = base + lane;
        int tot = 0;
        if (i < K)
            for (int t = t0; t < t1; ++t) {
                const size_t o = (size_t)(t - t0) * K + i;
                tb[o] = tot;
                tot += cnt[o];
            }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(LIP_FULL, inc, o); if (lane >= o) inc += u; }
        if (i < K) a.bin_off[bb + i] = ub + running + inc - tot;
        running += __shfl_sync(LIP_FULL, inc, 31);
    }
    if (m == a.M - 1 && lane == 0) a.bin_off[a.B] = a.n_total;
    if (lane < 3) {                                          // vertex sum of the mesh (mean-shift fallback of the fill step), fixed order
        double s = 0;
        for (int t = t0; t < t1; ++t) s += a.tile_sum[4 * (size_t)t + lane];
        a.mesh_sum[4 * (size_t)m + lane] = s;
    }
    if (lane == 3) a.mesh_sum[4 * (size_t)m + 3] = (double)n;
}

// Stable placement of a tile's vertices into the member lists of their bins (ascending vertex id inside a bin, as the boolean
// masks of the reference give it).  Rank of a vertex among the tile's members of its bin = rank inside the warp
// (__match_any_sync) + members in the earlier warps of the round (per-warp counts) + members in earlier rounds (cursors).  The
// records (fp32 offset from the bin's own landmark, vertex id) are first sorted by bin inside SMEM, so that the members a tile
// contributes to one bin leave as one contiguous run: coalesced 16-byte stores instead of one scattered store per vertex.
__global__ void __launch_bounds__(LIP_PART_THREADS) partition_place_kernel(LipArgs a) {
    extern __shared__ int smem_i[];
    __shared__ float4 rec[LIP_PART_CHUNK];
    __shared__ int dst[LIP_PART_CHUNK];
    __shared__ int wtot[LIP_PART_THREADS / 32];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LIP_PART_THREADS / 32;
    const int m = a.ptile_mesh[t], first = a.ptile_first[t];
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
    const double* __restrict__ xyz = mesh_xyz(a, m);
    const double* __restrict__ lm = mesh_lm(a, m);
    const int* __restrict__ cnt = a.tile_cnt + a.ptile_cnt[t];
    const int* __restrict__ tb = a.tile_base + a.ptile_cnt[t];
    int* lstart = smem_i;                                    // K: first SMEM slot of every bin's run, then its cursor
    int* gbase = lstart + K;                                 // K: global position of the run's first member minus its first slot
    int* wcnt = gbase + K;                                   // NW x K per-warp counts of the current round
    // exclusive scan of the tile's K counts (each thread a contiguous piece)
    const int per = (K + LIP_PART_THREADS - 1) / LIP_PART_THREADS;
    int mine = 0;
    for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < K) mine += cnt[i]; }
    int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(LIP_FULL, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) wtot[warp] = inc;
    for (int i = tid; i < NW * K; i += LIP_PART_THREADS) wcnt[i] = 0;
    __syncthreads();
    int run = inc - mine;
    for (int w = 0; w < warp; ++w) run += wtot[w];
    for (int k = 0; k < per; ++k) {
        const int i = tid * per + k;
        if (i < K) { lstart[i] = run; gbase[i] = a.bin_off[bb + i] + tb[i] - run; run += cnt[i]; }
    }
    const int end = min(n, first + LIP_PART_CHUNK);
    __syncthreads();
    for (int base = first; base < end; base += LIP_PART_THREADS) {
        const int v = base + tid;
        int bin = -1;
        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < end) {
            bin = a.nn0[ub + v];
            r4 = make_float4((float)(xyz[3 * (size_t)v] - lm[3 * bin]), (float)(xyz[3 * (size_t)v + 1] - lm[3 * bin + 1]),
                             (float)(xyz[3 * (size_t)v + 2] - lm[3 * bin + 2]), __int_as_float(v));
        }
        const unsigned peers = __match_any_sync(LIP_FULL, bin);
        const int rank = __popc(peers & ((1u << lane) - 1));
        if (bin >= 0 && rank == 0) wcnt[warp * K + bin] = __popc(peers);
        __syncthreads();
        if (bin >= 0) {
            int slot = lstart[bin] + rank;
            for (int w = 0; w < warp; ++w) slot += wcnt[w * K + bin];
            rec[slot] = r4;
            dst[slot] = gbase[bin] + slot;
        }
        __syncthreads();
        if (bin >= 0 && rank == 0) {                         // one lane per (warp, bin) present in the round
            atomicAdd(&lstart[bin], __popc(peers));          // integer adds commute: the cursors stay deterministic
            wcnt[warp * K + bin] = 0;
        }
        __syncthreads();
    }
    for (int j = tid; j < end - first; j += LIP_PART_THREADS) a.bin_tmp[dst[j]] = rec[j];
}

// per bin (one warp): bounding box of the fp32 offsets, float64 mean of the vertices (in member order: the means
// give my_icp's initial translation, registration_icp.py:70, and the centre of the ICP frame), and the FINAL
// order of the bin: members sorted along the axis of the box's largest extent (stable counting sort by an
// 8-bit key = position along the axis in 1/256 of the extent): spatially neighbouring queries share a warp (their
// winners' float64 gathers hit the same lines) and the boxes of query tiles are tight.  Ties are always decided by vertex id.
// Bins of up to 256 members (the usual case) are held in registers across the passes and permuted through SMEM: every
// global store coalesced.
// extent along the sort axis of each of the LIP_MAP_SUB consecutive ranges of a bin's final order (8 lanes per range; `at(pr)` =
// member pr of that order): with the bin's box in the other two dimensions these are the boxes the final map culls and stages by
template <typename F>
__device__ __forceinline__ void sub_boxes(const LipArgs& a, int b, int nb, int axis, int lane, F at) {
    static_assert(LIP_MAP_SUB == 4, "8 lanes per range");
    const int per = (nb + LIP_MAP_SUB - 1) / LIP_MAP_SUB, sub = lane >> 3;
    const int r0 = min(nb, sub * per), r1 = min(nb, r0 + per);
    float lo = INFINITY, hi = -INFINITY;
    for (int pr = r0 + (lane & 7); pr < r1; pr += 8) {
        const float4 t = at(pr);
        const float c = axis == 0 ? t.x : axis == 1 ? t.y : t.z;
        lo = fminf(lo, c); hi = fmaxf(hi, c);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(LIP_FULL, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(LIP_FULL, hi, o)); }
    if ((lane & 7) == 0) {
        float* sb = reinterpret_cast<float*>(a.sub_box + 2 * (size_t)b);
        sb[sub] = lo; sb[LIP_MAP_SUB + sub] = hi;            // empty range: lo > hi
    }
}

#ifndef LIP_SORT_BATCH
#define LIP_SORT_BATCH 2      // members per lane whose gathers are issued together in the final pass of bin_sort_kernel
#endif
#define LIP_SORT_WARPS 4
#ifndef LIP_SORT_WIDE
#define LIP_SORT_WIDE 4       // chunks of 32 members in flight per pass over a bin of more than 256 members
#endif
__global__ void __launch_bounds__(LIP_SORT_WARPS * 32) bin_sort_kernel(LipArgs a) {
    __shared__ unsigned skey_all[LIP_SORT_WARPS][256];
    __shared__ float4 sdat_all[LIP_SORT_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * LIP_SORT_WARPS + warp;
    if (b >= a.B) return;
    const int m = a.bin_mesh[b];
    const int ub = a.mesh_ubase[m];
    const double* __restrict__ xyz = mesh_xyz(a, m);
    unsigned* skey = skey_all[warp];
    float4* sdat = sdat_all[warp];
    const size_t N = (size_t)a.n_total;
    {
    const int s0 = a.bin_off[b], e0 = a.bin_off[b + 1], nb = e0 - s0;
    const bool fast = nb <= 256;
    float4 tt[8];
    float lo0 = INFINITY, lo1 = INFINITY, lo2 = INFINITY, hi0 = -INFINITY, hi1 = -INFINITY, hi2 = -INFINITY;
    double bx0 = 0, bx1 = 0, bx2 = 0;
    if (fast) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int e = lane + 32 * k; if (e < nb) tt[k] = a.bin_tmp[s0 + e]; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (lane + 32 * k >= nb) continue;
            const float4 t = tt[k];
            lo0 = fminf(lo0, t.x); lo1 = fminf(lo1, t.y); lo2 = fminf(lo2, t.z);
            hi0 = fmaxf(hi0, t.x); hi1 = fmaxf(hi1, t.y); hi2 = fmaxf(hi2, t.z);
            const size_t v = (size_t)__float_as_int(t.w);
            bx0 += xyz[3 * v]; bx1 += xyz[3 * v + 1]; bx2 += xyz[3 * v + 2];
        }
    } else {
        // large bins (whole organs of K == 1 segments): the slowest warp of the launch in a small batch, so LIP_SORT_WIDE chunks
        // of 32 members are in flight together in every pass (same per-lane summation order as one chunk at a time)
        for (int p0 = s0 + lane; p0 < e0; p0 += 32 * LIP_SORT_WIDE) {
            float4 t[LIP_SORT_WIDE];
            double gx[LIP_SORT_WIDE], gy[LIP_SORT_WIDE], gz[LIP_SORT_WIDE];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k) if (p0 + 32 * k < e0) t[k] = a.bin_tmp[p0 + 32 * k];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k)
                if (p0 + 32 * k < e0) {
                    const size_t v = (size_t)__float_as_int(t[k].w);
                    gx[k] = xyz[3 * v]; gy[k] = xyz[3 * v + 1]; gz[k] = xyz[3 * v + 2];
                }
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k)
                if (p0 + 32 * k < e0) {
                    lo0 = fminf(lo0, t[k].x); lo1 = fminf(lo1, t[k].y); lo2 = fminf(lo2, t[k].z);
                    hi0 = fmaxf(hi0, t[k].x); hi1 = fmaxf(hi1, t[k].y); hi2 = fmaxf(hi2, t[k].z);
                    bx0 += gx[k]; bx1 += gy[k]; bx2 += gz[k];
                }
        }
    }
    lo0 = -warp_maxf(-lo0); lo1 = -warp_maxf(-lo1); lo2 = -warp_maxf(-lo2);
    hi0 = warp_maxf(hi0); hi1 = warp_maxf(hi1); hi2 = warp_maxf(hi2);
    bx0 = warp_sum(bx0); bx1 = warp_sum(bx1); bx2 = warp_sum(bx2);
    const float ex0 = hi0 - lo0, ex1 = hi1 - lo1, ex2 = hi2 - lo2;
    const int axis = (ex0 >= ex1 && ex0 >= ex2) ? 0 : (ex1 >= ex2) ? 1 : 2;
    if (lane == 0) {
        float* bx = a.bin_box + 8 * (size_t)(b);
        bx[0] = lo0; bx[1] = lo1; bx[2] = lo2; bx[3] = hi0; bx[4] = hi1; bx[5] = hi2;
        bx[6] = fmaxf(fmaxf(fmaxf(fabsf(lo0), fabsf(hi0)), fmaxf(fabsf(lo1), fabsf(hi1))), fmaxf(fabsf(lo2), fabsf(hi2)));
        bx[7] = (float)axis;
        double* bs = a.bin_sum + 4 * (size_t)(b);
        const double cnt = (double)nb;
        bs[0] = bx0 / cnt; bs[1] = bx1 / cnt; bs[2] = bx2 / cnt; bs[3] = cnt;     // (NaN for an empty bin: never read)
    }
    const float alo = axis == 0 ? lo0 : axis == 1 ? lo1 : lo2, aex = axis == 0 ? ex0 : axis == 1 ? ex1 : ex2;
    const float scale = 255.999f / fmaxf(aex, 1e-30f);
    auto bucket_of = [&](const float4& t) {
        const float c = axis == 0 ? t.x : axis == 1 ? t.y : t.z;
        return (int)fminf(fmaxf((c - alo) * scale, 0.f), 255.f);
    };
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) skey[lane + 32 * k] = 0u;
    __syncwarp();
    if (fast) {
#pragma unroll
        for (int k = 0; k < 8; ++k) if (lane + 32 * k < nb) atomicAdd(&skey[bucket_of(tt[k])], 1u);
    } else {
        for (int e0w = lane; e0w < nb; e0w += 32 * LIP_SORT_WIDE) {
            float4 t[LIP_SORT_WIDE];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k) if (e0w + 32 * k < nb) t[k] = a.bin_tmp[s0 + e0w + 32 * k];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k) if (e0w + 32 * k < nb) atomicAdd(&skey[bucket_of(t[k])], 1u);
        }
    }
    __syncwarp();
    {                                                    // exclusive scan of the 256 counts: lane l owns buckets 8l .. 8l+7
        unsigned c8[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c8[k] = skey[8 * lane + k]; tot += c8[k]; }
        unsigned inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(LIP_FULL, inc, o); if (lane >= o) inc += t; }
        unsigned run = inc - tot;
#pragma unroll
        for (int k = 0; k < 8; ++k) { skey[8 * lane + k] = run; run += c8[k]; }
    }
    __syncwarp();
    // placement, chunk by chunk in member order: rank among the chunk's members of the same bucket + bucket cursor
    if (fast) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (32 * k >= nb) break;                     // uniform
            const int e = lane + 32 * k;
            const int bucket = e < nb ? bucket_of(tt[k]) : -1;
            const unsigned peers = __match_any_sync(LIP_FULL, bucket);
            const int rk = __popc(peers & ((1u << lane) - 1));
            if (e < nb) sdat[(int)skey[bucket] + rk] = tt[k];
            __syncwarp();
            if (e < nb && rk == 0) skey[bucket] += __popc(peers);
            __syncwarp();
        }
        sub_boxes(a, b, nb, axis, lane, [&](int pr) { return sdat[pr]; });
#if LIP_SORT_BATCH > 1
        // coalesced stores of everything that is kept in bin order; the gathers by vertex id of LIP_SORT_BATCH members are in
        // flight together (the kernel waits on L2 latency, not on bandwidth)
        for (int pr0 = lane; pr0 < nb; pr0 += 32 * LIP_SORT_BATCH) {
            double gx[LIP_SORT_BATCH], gy[LIP_SORT_BATCH], gz[LIP_SORT_BATCH], g0[LIP_SORT_BATCH], g1[LIP_SORT_BATCH];
            int gn[LIP_SORT_BATCH];
            float4 t[LIP_SORT_BATCH];
#pragma unroll
            for (int k = 0; k < LIP_SORT_BATCH; ++k) {
                const int pr = pr0 + 32 * k;
                if (pr < nb) {
                    t[k] = sdat[pr];
                    const size_t v = (size_t)__float_as_int(t[k].w);
                    gx[k] = xyz[3 * v]; gy[k] = xyz[3 * v + 1]; gz[k] = xyz[3 * v + 2];
                    gn[k] = a.nn1[ub + v]; g0[k] = a.d0[ub + v]; g1[k] = a.d1[ub + v];
                }
            }
#pragma unroll
            for (int k = 0; k < LIP_SORT_BATCH; ++k) {
                const int pr = pr0 + 32 * k;
                if (pr < nb) {
                    const size_t p = (size_t)s0 + pr;
                    a.part_vid[p] = __float_as_int(t[k].w);
                    a.bin_pts[p] = t[k];
                    a.bin_xyz[p] = gx[k]; a.bin_xyz[N + p] = gy[k]; a.bin_xyz[2 * N + p] = gz[k];
                    a.pos_bin[p] = b; a.pos_nn1[p] = gn[k];
                    a.pos_d[p] = g0[k]; a.pos_d[N + p] = g1[k];
                }
            }
        }
#else
        for (int pr = lane; pr < nb; pr += 32) {         // coalesced stores of everything that is kept in bin order
            const float4 t = sdat[pr];
            const size_t v = (size_t)__float_as_int(t.w), p = (size_t)s0 + pr;
            a.part_vid[p] = (int)v;
            a.bin_pts[p] = t;
            a.bin_xyz[p] = xyz[3 * v]; a.bin_xyz[N + p] = xyz[3 * v + 1]; a.bin_xyz[2 * N + p] = xyz[3 * v + 2];
            a.pos_bin[p] = b; a.pos_nn1[p] = a.nn1[ub + v];
            a.pos_d[p] = a.d0[ub + v]; a.pos_d[N + p] = a.d1[ub + v];
        }
#endif
    } else {
        for (int base = 0; base < nb; base += 32 * LIP_SORT_WIDE) {
            float4 t[LIP_SORT_WIDE];
            size_t dstp[LIP_SORT_WIDE];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k) {
                t[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + 32 * k + lane < nb) t[k] = a.bin_tmp[s0 + base + 32 * k + lane];
            }
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k) {       // the cursors advance chunk by chunk, in member order
                if (base + 32 * k >= nb) break;             // uniform
                const bool in = base + 32 * k + lane < nb;
                const int bucket = in ? bucket_of(t[k]) : -1;
                const unsigned peers = __match_any_sync(LIP_FULL, bucket);
                const int rk = __popc(peers & ((1u << lane) - 1));
                if (in) dstp[k] = (size_t)s0 + skey[bucket] + rk;
                __syncwarp();
                if (in && rk == 0) skey[bucket] += __popc(peers);
                __syncwarp();
            }
            double gx[LIP_SORT_WIDE], gy[LIP_SORT_WIDE], gz[LIP_SORT_WIDE], g0[LIP_SORT_WIDE], g1[LIP_SORT_WIDE];
            int gn[LIP_SORT_WIDE];
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k)
                if (base + 32 * k + lane < nb) {
                    const size_t v = (size_t)__float_as_int(t[k].w);
                    gx[k] = xyz[3 * v]; gy[k] = xyz[3 * v + 1]; gz[k] = xyz[3 * v + 2];
                    gn[k] = a.nn1[ub + v]; g0[k] = a.d0[ub + v]; g1[k] = a.d1[ub + v];
                }
#pragma unroll
            for (int k = 0; k < LIP_SORT_WIDE; ++k)
                if (base + 32 * k + lane < nb) {
                    const size_t p = dstp[k];
                    a.part_vid[p] = __float_as_int(t[k].w);
                    a.bin_pts[p] = t[k];
                    a.bin_xyz[p] = gx[k]; a.bin_xyz[N + p] = gy[k]; a.bin_xyz[2 * N + p] = gz[k];
                    a.pos_bin[p] = b; a.pos_nn1[p] = gn[k];
                    a.pos_d[p] = g0[k]; a.pos_d[N + p] = g1[k];
                }
        }
        __syncwarp();                                        // the warp's own stores to bin_pts, read back in final order
        sub_boxes(a, b, nb, axis, lane, [&](int pr) { return a.bin_pts[(size_t)s0 + pr]; });
    }

    }
}

// =============================================================================================
// ICP scheduling: one thread per bin; empty src or tgt -> "no ICP" (p2p_matching_seg_icp.py:415-419)
// =============================================================================================
__device__ __forceinline__ IcpFrame icp_frame(const LipArgs& a, int b);
__global__ void icp_schedule_kernel(LipArgs a) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    int m = a.bin_mesh[b];
    a.icp_iters[b] = 0;
    a.icp_status[b] = 0;
    if (!src_active(a, m)) return;
    int pb = pair_bin(a, b);
    int ns = a.bin_off[b + 1] - a.bin_off[b], nt = a.bin_off[pb + 1] - a.bin_off[pb];
    if (ns > 0) {                                            // tiles of the final map: up to LIP_MAP_TILE queries of one bin
        const int nt_map = (ns + LIP_MAP_TILE - 1) / LIP_MAP_TILE;
        const int slot = atomicAdd(&a.sched[20], nt_map);
        MapTile t;
        const int mt = pair_mesh(a, m);
        t.b = b; t.m = m; t.bb = a.mesh_bin[m]; t.K = a.mesh_bin[m + 1] - t.bb; t.bbt = a.mesh_bin[mt];
        t.so = a.bin_off[b]; t.ub = a.mesh_ubase[m];
        t.ts = a.bin_off[pb]; t.te = a.bin_off[pb + 1];
        t.home_mabs = (t.te > t.ts) ? a.bin_box[8 * (size_t)pb + 6] : 0.f;
        const double* lmh = mesh_lm(a, mt) + 3 * (size_t)(b - t.bb);
        t.C[0] = lmh[0]; t.C[1] = lmh[1]; t.C[2] = lmh[2];
        for (int i = 0; i < nt_map; ++i) { t.first = i * LIP_MAP_TILE; t.nq = min(LIP_MAP_TILE, ns - t.first); a.map_tiles[slot + i] = t; }
    }
    if (ns == 0 || nt == 0) return;
    a.icp_frames[b] = icp_frame(a, b);                       // one 64-byte record per ICP instead of a chain of look-ups per pass
    const long long w = (long long)ns * nt;
    if (w <= a.warp_work) {
        // eight work classes (factor 2 apart); the persistent warps take groups of partitions heaviest class first
        int cls = 0;
        for (long long t = 1024; t <= w && cls < 7; t <<= 1) ++cls;
        a.warp_list[(size_t)cls * a.B + atomicAdd(&a.sched[24 + cls], 1)] = b;
        atomicAdd(&a.sched[0], 1);
    } else {
        const int cls = (w > 1000000) ? 3 : (w > 400000) ? 2 : (w > 150000) ? 1 : 0;
        a.big_list[(size_t)cls * a.B + atomicAdd(&a.sched[12 + cls], 1)] = b;
        atomicAdd(&a.sched[1], 1);
    }
}

// =============================================================================================
// Scan primitives shared by the ICP kernels and the final map: an SMEM tile of fp32 candidates (float4,
// DOT form: .w = |y|^2 and the query passed as m = -2q, score c = |y|^2 - 2 q.y = d^2 - |q|^2 in 3 FFMA;
// difference form: m = q, score d^2) is scanned by all lanes at once (broadcast reads).  Per block of 8
// candidates only the block minimum enters the top-2 tracking (FMNMX3), the winner's position is
// recovered afterwards by re-evaluating its block with the identical instruction sequence.
// =============================================================================================
template <bool DOT>
__device__ __forceinline__ float tile_score(const float4 t, const float* m) {
    if (DOT) return fmaf(m[0], t.x, fmaf(m[1], t.y, fmaf(m[2], t.z, t.w)));
    const float dx = t.x - m[0], dy = t.y - m[1], dz = t.z - m[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
template <int Q, bool DOT, int QMAX>
__device__ __forceinline__ void tile_scan(const float4* __restrict__ tp, int npad, const float (&m)[QMAX][3], float (&best)[QMAX],
                                          float (&sec)[QMAX], int (&blk)[QMAX]) {
#pragma unroll 1
    for (int jb = 0; jb < npad; jb += 8) {
        float4 t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = tp[jb + k];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float c[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) c[k] = tile_score<DOT>(t[k], m[q]);
            float bm = fminf(fminf(c[0], c[1]), c[2]);
            bm = fminf(fminf(bm, c[3]), c[4]);
            bm = fminf(fminf(bm, c[5]), c[6]);
            bm = fminf(bm, c[7]);
            sec[q] = fminf(sec[q], fmaxf(bm, best[q]));
            blk[q] = (bm < best[q]) ? jb : blk[q];
            best[q] = fminf(best[q], bm);
        }
    }
}
// The same scan over a tile whose candidates are consecutive members of an axis-sorted bin (bin_sort_kernel): a block of 8 whose
// interval along that axis lies farther from every query of the warp than that query's bound R (an upper bound of its nearest-
// neighbour distance, margins and the sort's bucket width included) cannot hold a nearest neighbour or a tie partner and is not
// evaluated.  The last block of a tile holds the padding and is always evaluated.  Returns the number of blocks evaluated.
template <int Q, bool DOT, int QMAX>
__device__ __forceinline__ int tile_scan_skip(const float4* __restrict__ tp, int npad, const float (&m)[QMAX][3], float (&best)[QMAX],
                                              float (&sec)[QMAX], int (&blk)[QMAX], int axis, const float (&xq)[QMAX], const float (&R)[QMAX]) {
    const float* __restrict__ ax = reinterpret_cast<const float*>(tp) + axis;
    int done = 0;
#pragma unroll 1
    for (int jg = 0; jg < npad; jg += LIP_ICP_SKIP) {       // groups of LIP_ICP_SKIP candidates share one test
        const int je = min(npad, jg + LIP_ICP_SKIP);
        if (je < npad) {
            const float cf = ax[4 * jg], cl = ax[4 * (je - 1)];
            bool need = false;
#pragma unroll
            for (int q = 0; q < Q; ++q) need = need || !(fmaxf(cf - xq[q], xq[q] - cl) > R[q]);
            if (!__any_sync(LIP_FULL, need)) continue;
        }
#pragma unroll 1
        for (int jb = jg; jb < je; jb += 8) {
            ++done;
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = tp[jb + k];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float c[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) c[k] = tile_score<DOT>(t[k], m[q]);
                float bm = fminf(fminf(c[0], c[1]), c[2]);
                bm = fminf(fminf(bm, c[3]), c[4]);
                bm = fminf(fminf(bm, c[5]), c[6]);
                bm = fminf(bm, c[7]);
                sec[q] = fminf(sec[q], fmaxf(bm, best[q]));
                blk[q] = (bm < best[q]) ? jb : blk[q];
                best[q] = fminf(best[q], bm);
            }
        }
    }
    return done;
}
// position of the winner of block `blk` (first candidate whose score equals `best` bitwise) and the smallest
// score among the other seven
template <bool DOT>
__device__ __forceinline__ int tile_recover(const float4* __restrict__ tp, int blk, float best, const float* m, float& sec_in) {
    int pos = -1;
    float s2 = INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float c = tile_score<DOT>(tp[blk + k], m);
        if (c == best && pos < 0) pos = blk + k;
        else s2 = fminf(s2, c);
    }
    sec_in = s2;
    return pos < 0 ? blk : pos;
}
__device__ __forceinline__ float4 tile_sentinel(bool dot) {       // never the best, never a runner-up
    return dot ? make_float4(0.f, 0.f, 0.f, 3.0e38f) : make_float4(1e18f, 1e18f, 1e18f, 0.f);
}
// Bound on |evaluated - exact| of the dot-form score of a candidate at distance <= sqrt(d2) of a query with
// |q|^2 = qq: three FMAs and the stored |y|^2 (three roundings), every partial result bounded by
// |y|^2 + 2 |q| |y| with |y| <= |q| + d.   6 u (.) with u = 2^-24, times 1.5 for slack.
__device__ __forceinline__ float dot_err(float qq, float d2) {
    const float nq = sqrtf(qq), ny = nq + 1.001f * sqrtf(d2) + 1e-3f;
    return 5.4e-7f * fmaf(ny, ny, 2.f * nq * ny);
}

// =============================================================================================
// K1-K3  partition ICP, whole iteration loop on device (registration/registration_icp.py:60-98)
//   NN (get_NN_in_target :49-57): fp32 brute force over the SMEM-staged target in coordinates
//   centred on the target centroid, top-2 tracking; a query whose two best candidates are closer
//   than the fp32 error bound is re-evaluated in float64 (lowest index wins an exact tie).
//   Kabsch (get_best_fit_transform :8-46): float64 sums reduced over the warp with a fixed butterfly,
//   3x3 Jacobi SVD in registers; accumulate R,t; cost/stop test (:86-88) — all uniform per warp.
// =============================================================================================
// ICP state of one partition.  It lives in SHARED memory (one per warp / CTA): every lane computes the
// identical update and stores the identical values, so the state costs no registers across the NN scan.
struct IcpState {
    double R[9], t[3];       // accumulated transform in centred coordinates
    double Ri[9], ti[3];     // last increment
};

__device__ __forceinline__ void icp_state_init(IcpState* s, double tix, double tiy, double tiz) {
#pragma unroll
    for (int i = 0; i < 9; ++i) { s->R[i] = (i % 4 == 0) ? 1.0 : 0.0; s->Ri[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    s->t[0] = s->t[1] = s->t[2] = 0.0;
    s->ti[0] = tix; s->ti[1] = tiy; s->ti[2] = tiz;
}
// R <- Ri R ; t <- Ri t + ti   (registration_icp.py:83-84)
__device__ __forceinline__ void icp_state_accumulate(IcpState* s) {
    double Ri[9], R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { Ri[i] = s->Ri[i]; R[i] = s->R[i]; }
    const double t0 = s->t[0], t1 = s->t[1], t2 = s->t[2];
    mat3_mul(Ri, R, R);
    double x, y, z;
    mat3_vec(Ri, t0, t1, t2, x, y, z);
    x += s->ti[0]; y += s->ti[1]; z += s->ti[2];
#pragma unroll
    for (int i = 0; i < 9; ++i) s->R[i] = R[i];
    s->t[0] = x; s->t[1] = y; s->t[2] = z;
}
// acc: 0 sum dist | 1..3 sum p | 4..6 sum b | 7..15 sum p_i b_j | 16 pairs kept
// Kabsch increment from the reduced sums; returns 1 if the solve was rank deficient
__device__ __forceinline__ int icp_kabsch_step(const double* acc, IcpState* s) {
    double n = acc[16];
    double Ri[9];
    if (!(n > 0.0)) {
#pragma unroll
        for (int i = 0; i < 9; ++i) s->Ri[i] = (i % 4 == 0) ? 1.0 : 0.0;
        s->ti[0] = s->ti[1] = s->ti[2] = 0.0;
        return 1;
    }
    double inv = 1.0 / n;
    double ca[3] = {acc[1] * inv, acc[2] * inv, acc[3] * inv};
    double cb[3] = {acc[4] * inv, acc[5] * inv, acc[6] * inv};
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) H[3 * i + j] = acc[7 + 3 * i + j] - n * ca[i] * cb[j];
    int deg = kabsch_rotation(H, Ri);
    double x, y, z;
    mat3_vec(Ri, ca[0], ca[1], ca[2], x, y, z);
#pragma unroll
    for (int i = 0; i < 9; ++i) s->Ri[i] = Ri[i];
    s->ti[0] = cb[0] - x; s->ti[1] = cb[1] - y; s->ti[2] = cb[2] - z;
    return deg;
}
// final transform back to the original frame: y - c = R (x - c) + t~  =>  t = t~ + c - R c
__device__ __forceinline__ void icp_store(const LipArgs& a, int b, const IcpState* s, double cx, double cy, double cz,
                                          int iters, int status) {
    double x, y, z;
    mat3_vec(s->R, cx, cy, cz, x, y, z);
#pragma unroll
    for (int i = 0; i < 9; ++i) a.icp_R[9 * (size_t)b + i] = s->R[i];
    a.icp_t[3 * (size_t)b] = s->t[0] + cx - x;
    a.icp_t[3 * (size_t)b + 1] = s->t[1] + cy - y;
    a.icp_t[3 * (size_t)b + 2] = s->t[2] + cz - z;
    a.icp_iters[b] = iters;
    a.icp_status[b] = status;
}

// ---- one NN pass + Kabsch sums of a partition, by a warp (NT = 32) or by a CTA (NT = blockDim) -------------
// The partition's points are read in bin order: a.bin_pts (fp32 offsets from the target's landmark, 16-byte
// coalesced loads) for the search tile, a.bin_xyz (float64 planes x|y|z in bin order) for everything float64.
template <int NT>
__device__ __forceinline__ void coop_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}
template <bool DOT, int NT>
__device__ __forceinline__ void icp_stage(const LipArgs& a, float4* tf, int to, int c0, int cnt, int tid,
                                          float sh0, float sh1, float sh2) {
    const int npad = (cnt + 7) & ~7;
    for (int j = tid; j < npad; j += NT) {
        float4 t = tile_sentinel(DOT);
        if (j < cnt) {
            t = a.bin_pts[to + c0 + j];
            t.x += sh0; t.y += sh1; t.z += sh2;
            t.w = DOT ? fmaf(t.x, t.x, fmaf(t.y, t.y, t.z * t.z)) : 0.f;
        }
        tf[j] = t;
    }
}
// source bin at so/ns, target bin at to/nt; frame centred on c (the target centroid); the tile points are the
// stored offsets shifted by sh = landmark - c.  Returns this thread's partial sums in acc[17]:
// 0 sum dist | 1..3 sum p | 4..6 sum b | 7..15 sum p_i b_j | 16 pairs kept.
template <bool DOT, int NT, int CHUNK, bool TS>
__device__ __forceinline__ void icp_pass(const LipArgs& a, float4* tf, unsigned char* nns, const IcpState* st, int tid,
                                         int so, int ns, int to, int nt, double c0, double c1, double c2,
                                         float sh0, float sh1, float sh2, float mabs, double* __restrict__ out17,
                                         int axis, float wb, bool warm) {
    const size_t N = (size_t)a.n_total;
    const double* __restrict__ sx = a.bin_xyz + so;
    const double* __restrict__ tx = a.bin_xyz + to;
    int* __restrict__ nn_idx = a.nn_idx + so;
    const float sc = DOT ? -2.0f : 1.0f;
    const bool one_chunk = nt <= CHUNK;
    const bool staged = TS && one_chunk;                     // winners fit one byte: they stay in SMEM between the two phases
    // Axis block skip (tile_scan_skip): only the CTA kernels, and only on targets long enough for it to pay — measured: the warp
    // kernel loses 19 % to the tests (its targets are ~150 points, half of whose blocks every 64-query round needs), a CTA kernel
    // with 150 k - 1 M pair partitions gains 39 %, with 100 k pair partitions (targets of ~300) it loses 13 %.
    const bool skip = !TS && NT > 32 && nt >= LIP_ICP_SKIP_MIN;
    if (one_chunk) { coop_sync<NT>(); icp_stage<DOT, NT>(a, tf, to, 0, nt, tid, sh0, sh1, sh2); coop_sync<NT>(); }
    unsigned long long scanned = 0;
    for (int q0 = 0; q0 < ns; q0 += 2 * NT) {
        const bool two = ns - q0 > NT;                       // uniform
        const int qi[2] = {q0 + tid, q0 + tid + NT};
        double pq[2][3];
        float m[2][3], best[2], sec[2], sec_in[2], xq[2], Rw[2];
        int blk[2], idx[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int q = min(qi[k], ns - 1);
            mat3_vec(st->R, sx[q] - c0, sx[N + q] - c1, sx[2 * N + q] - c2, pq[k][0], pq[k][1], pq[k][2]);
            pq[k][0] += st->t[0]; pq[k][1] += st->t[1]; pq[k][2] += st->t[2];
            m[k][0] = sc * (float)pq[k][0]; m[k][1] = sc * (float)pq[k][1]; m[k][2] = sc * (float)pq[k][2];
            best[k] = INFINITY; sec[k] = INFINITY; sec_in[k] = INFINITY; blk[k] = 0; idx[k] = 0;
            // bound for the block skip: the exact distance to a candidate that is probably near — the previous pass's winner, in the
            // first pass the member at the same relative position of the (equally sorted) target
            xq[k] = 0.f; Rw[k] = INFINITY;
            if (skip) {
                const int jw = min(max(warm ? nn_idx[q] : (int)(((long long)q * nt) / ns), 0), nt - 1);
                const float fxa = (float)(axis == 0 ? pq[k][0] : axis == 1 ? pq[k][1] : pq[k][2]);
                xq[k] = fxa;
                float dwf;
                if (one_chunk) {                             // the staged tile holds it: no second trip to global memory
                    const float4 tw = tf[jw];
                    const float dx = tw.x - (float)pq[k][0], dy = tw.y - (float)pq[k][1], dz = tw.z - (float)pq[k][2];
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float qa = fmaxf(fmaxf(fabsf((float)pq[k][0]), fabsf((float)pq[k][1])), fabsf((float)pq[k][2]));
                    dwf = sqrtf(d2 + d2_err(d2, 2.4e-7f * (qa + mabs)));
                } else {
                    dwf = (float)sqrt(rdist64(pq[k][0], pq[k][1], pq[k][2], tx[jw] - c0, tx[N + jw] - c1, tx[2 * N + jw] - c2));
                }
                Rw[k] = dwf * 1.00001f + 1e-6f * (fabsf(fxa) + mabs) + wb;
            }
        }
        for (int t0 = 0; t0 < nt; t0 += CHUNK) {
            const int cnt = min(CHUNK, nt - t0), npad = (cnt + 7) & ~7;
            if (!one_chunk) { coop_sync<NT>(); icp_stage<DOT, NT>(a, tf, to, t0, cnt, tid, sh0, sh1, sh2); coop_sync<NT>(); }
            const float prev0 = best[0], prev1 = best[1];
            int nblk = npad / 8;
            if (skip) {
                nblk = two ? tile_scan_skip<2, DOT>(tf, npad, m, best, sec, blk, axis, xq, Rw)
                           : tile_scan_skip<1, DOT>(tf, npad, m, best, sec, blk, axis, xq, Rw);
            } else if (two) tile_scan<2, DOT>(tf, npad, m, best, sec, blk);
            else tile_scan<1, DOT>(tf, npad, m, best, sec, blk);
            if (best[0] < prev0) idx[0] = t0 + tile_recover<DOT>(tf, blk[0], best[0], m[0], sec_in[0]);
            if (best[1] < prev1) idx[1] = t0 + tile_recover<DOT>(tf, blk[1], best[1], m[1], sec_in[1]);
            scanned += (unsigned long long)nblk * (two ? 8 * 64 : 8 * 32);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float fx = (float)pq[k][0], fy = (float)pq[k][1], fz = (float)pq[k][2];
            const float qabs = fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz));
            const float e = 2.4e-7f * (qabs + mabs);         // stored offset, shift and sum: three roundings per coordinate
            const float qq = DOT ? fmaf(fx, fx, fmaf(fy, fy, fz * fz)) : 0.f;
            const float d2b = fmaxf(best[k] + qq, 0.f), d2s = fmaxf(fminf(sec[k], sec_in[k]) + qq, 0.f);
            float tol = d2_err(d2b, e) + d2_err(d2s, e);
            if (DOT) tol += 2.f * dot_err(qq, d2s);
            // near ties are decided in float64 (lowest index), the warp working together on one query at a time
            unsigned ties = __ballot_sync(LIP_FULL, qi[k] < ns && !(d2s - d2b > tol));
            if (ties && (tid & 31) == 0) atomicAdd(&a.counters[2], (unsigned long long)__popc(ties));
            while (ties) {
                const int src = __ffs(ties) - 1;
                ties &= ties - 1;
                const double p0 = __shfl_sync(LIP_FULL, pq[k][0], src), p1 = __shfl_sync(LIP_FULL, pq[k][1], src),
                             p2 = __shfl_sync(LIP_FULL, pq[k][2], src);
                double bd = INFINITY;
                int bj = 0, bv = 0x7fffffff;                 // position and vertex id: the lowest vertex id wins an exact tie
                for (int j = tid & 31; j < nt; j += 32) {
                    const double d = rdist64(p0, p1, p2, tx[j] - c0, tx[N + j] - c1, tx[2 * N + j] - c2);
                    if (d <= bd) {
                        const int vj = a.part_vid[to + j];
                        if (d < bd || vj < bv) { bd = d; bj = j; bv = vj; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(LIP_FULL, bd, o);
                    const int oj = __shfl_xor_sync(LIP_FULL, bj, o), ov = __shfl_xor_sync(LIP_FULL, bv, o);
                    if (od < bd || (od == bd && ov < bv)) { bd = od; bj = oj; bv = ov; }
                }
                if ((tid & 31) == src) idx[k] = bj;
            }
            if (qi[k] < ns) {
                if (staged && qi[k] < LIP_ICP_QCAP) nns[qi[k]] = (unsigned char)idx[k];
                else nn_idx[qi[k]] = idx[k];                 // (with `skip` also the next pass's bound: skip implies !staged)
            }
        }
    }
    if (a.profile && (tid & 31) == 0) atomicAdd(&a.counters[NT == 32 ? 4 : 5], scanned);   // candidate pairs evaluated (incl. padding)
    // ---- sums of get_best_fit_transform + cost over the (source, winner) pairs, float64 ----
    double acc[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = 0.0;
    const double reject = a.reject_dist;
    for (int q = tid; q < ns; q += NT) {
        const int w = (staged && q < LIP_ICP_QCAP) ? (int)nns[q] : nn_idx[q];     // written by this very lane
        double px, py, pz;
        mat3_vec(st->R, sx[q] - c0, sx[N + q] - c1, sx[2 * N + q] - c2, px, py, pz);
        px += st->t[0]; py += st->t[1]; pz += st->t[2];
        const double bx = tx[w] - c0, by = tx[N + w] - c1, bz = tx[2 * N + w] - c2;
        const double d = sqrt(rdist64(px, py, pz, bx, by, bz));
        acc[0] += d;
        if (d <= reject) {
            acc[1] += px; acc[2] += py; acc[3] += pz;
            acc[4] += bx; acc[5] += by; acc[6] += bz;
            acc[7] += px * bx; acc[8] += px * by; acc[9] += px * bz;
            acc[10] += py * bx; acc[11] += py * by; acc[12] += py * bz;
            acc[13] += pz * bx; acc[14] += pz * by; acc[15] += pz * bz;
            acc[16] += 1.0;
        }
    }
    // ---- the warp's 17 sums -> out17[0..16] ----
    const int lane = tid & 31;
#if LIP_ICP_SUM17
    // reduce-scatter through SMEM instead of 17 butterflies: the search tile is dead by now and serves as scratch, 16 rows
    // (sums 1..16) of 32 doubles per warp, column l of row i stored at l ^ i (conflict-free both ways); lane (r, h) adds
    // half h of row r, one shuffle joins the halves.  Sum 0 (the cost) takes the butterfly.
    coop_sync<NT>();
    double* scr = reinterpret_cast<double*>(tf) + (size_t)(tid >> 5) * (16 * 32);
#pragma unroll
    for (int i = 0; i < 16; ++i) scr[i * 32 + (lane ^ i)] = acc[i + 1];
    __syncwarp();
    {
        const int r = lane & 15, h = lane & 16;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += scr[r * 32 + ((h + k) ^ r)];
        s += __shfl_xor_sync(LIP_FULL, s, 16);
        if (lane < 16) out17[1 + r] = s;
    }
    const double s0 = warp_sum(acc[0]);
    if (lane == 0) out17[0] = s0;
    __syncwarp();
#else
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 17; ++i) out17[i] = acc[i];
    }
#endif
}

// frame of one ICP: centre = target centroid, tile shift = target landmark - centre, bound on |tile coordinate|
#if LIP_ICP_SUM17
static_assert((LIP_ICP_CHUNK + 8) * 16 >= 16 * 32 * 8 && (LIP_BIG_CHUNK + 8) * 16 >= (LIP_BIG_THREADS / 32) * 16 * 32 * 8,
              "the search tile doubles as the scratch of the 17-sum reduction: 4 KB per warp");
#endif
// (IcpFrame: lip_device.cuh) derived from the partition kernel's outputs
__device__ __forceinline__ IcpFrame icp_frame(const LipArgs& a, int b) {
    IcpFrame f;
    const int pb = pair_bin(a, b), mt = pair_mesh(a, a.bin_mesh[b]);
    f.so = a.bin_off[b]; f.ns = a.bin_off[b + 1] - f.so;
    f.to = a.bin_off[pb]; f.nt = a.bin_off[pb + 1] - f.to;
    const double* lmt = mesh_lm(a, mt) + 3 * (size_t)(pb - a.mesh_bin[mt]);
    const double* ct = a.bin_sum + 4 * (size_t)pb;
    f.c0 = ct[0]; f.c1 = ct[1]; f.c2 = ct[2];
    f.sh0 = (float)(lmt[0] - ct[0]); f.sh1 = (float)(lmt[1] - ct[1]); f.sh2 = (float)(lmt[2] - ct[2]);
    f.mabs = a.bin_box[8 * (size_t)pb + 6] + fmaxf(fmaxf(fabsf(f.sh0), fabsf(f.sh1)), fabsf(f.sh2));
    // bin_sort_kernel ordered the target bin by an 8-bit key along this axis: member i comes before member j => x_i < x_j + wb
    f.axis = min(max((int)a.bin_box[8 * (size_t)pb + 7], 0), 2);
    f.wb = (a.bin_box[8 * (size_t)pb + 3 + f.axis] - a.bin_box[8 * (size_t)pb + f.axis]) * (1.001f / 255.999f) + 1e-30f;
    return f;
}

// ---- warp-per-partition, groups of partitions per warp task ---------------------------------------------
// A persistent warp takes a task of g <= LIP_ICP_GMAX partitions of one work class.  Every NN pass is
// warp-cooperative (one partition after the other, lanes own the queries); the serial part of an iteration —
// cost / stop test, 3x3 Kabsch solve, transform accumulation — runs SIMT across the task's partitions, lane p
// working for partition p, so its cost is shared by g partitions instead of being paid 32-fold redundantly.
struct IcpGroupSmem {
    float4 tf[LIP_ICP_CHUNK + 8];                         // fp32 search tile
    IcpState st[LIP_ICP_GMAX];
    double acc[LIP_ICP_GMAX][17];
    unsigned char nn[LIP_ICP_QCAP ? LIP_ICP_QCAP : 1];                       // winner of every query of the current pass
};
// partitions per task of work class c (n_src * n_tgt in [512 * 2^c, 1024 * 2^c)): 8 up to 16k pairs, then 4, 2, 1
__device__ __forceinline__ int icp_class_group(int c, int gcap, int shift) { return max(1, min(gcap, LIP_ICP_GMAX >> max(c - shift, 0))); }

__global__ void __launch_bounds__(LIP_ICP_WARPS * 32, LIP_WARP_MINBLOCKS) icp_group_kernel(LipArgs a) {
    extern __shared__ __align__(16) unsigned char icp_group_smem[];
    IcpGroupSmem& sm = reinterpret_cast<IcpGroupSmem*>(icp_group_smem)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int n_items = a.sched[0];
    // group size cap: as large as leaves every resident warp a few tasks
    const int gcap = max(1, min(n_items / (int)(gridDim.x * LIP_ICP_WARPS * a.icp_tasks_per_warp), LIP_ICP_GMAX));
    const int gshift = a.icp_class_shift;
    int n_tasks = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { const int G = icp_class_group(c, gcap, gshift); n_tasks += (a.sched[24 + c] + G - 1) / G; }
    for (;;) {                                              // persistent warps, dynamic task queue, heaviest class first
        int w = 0;
        if (lane == 0) w = atomicAdd(&a.sched[2], 1);
        w = __shfl_sync(LIP_FULL, w, 0);
        if (w >= n_tasks) break;
        int c = 7, G = icp_class_group(7, gcap, gshift), tc = (a.sched[24 + 7] + G - 1) / G;
        while (c > 0 && w >= tc) { w -= tc; --c; G = icp_class_group(c, gcap, gshift); tc = (a.sched[24 + c] + G - 1) / G; }
        const int first = w * G, g = min(G, a.sched[24 + c] - first);
        // ---- lane p < g owns partition p of the task -------------------------------------------------------
        int b = 0, own_ns = 1, iters = 0, status = 1;
        bool active = lane < g;
        double prev = 0.0;
        if (active) {
            b = a.warp_list[(size_t)c * a.B + first + lane];
            own_ns = a.bin_off[b + 1] - a.bin_off[b];
            const double* ss = a.bin_sum + 4 * (size_t)b;
            const double* ts = a.bin_sum + 4 * (size_t)pair_bin(a, b);
            // increment 0: identity, t = mean(tgt) - mean(src)   (registration_icp.py:70)
            icp_state_init(&sm.st[lane], ts[0] - ss[0], ts[1] - ss[1], ts[2] - ss[2]);
        }
        __syncwarp();
        for (int it = 0; it < a.max_iter; ++it) {
            unsigned todo = __ballot_sync(LIP_FULL, active);
            if (!todo) break;
            if (active) icp_state_accumulate(&sm.st[lane]);     // R <- Ri R ; t <- Ri t + ti   (:83-84)
            __syncwarp();
            while (todo) {                                      // NN pass + sums, one partition after the other
                const int p = __ffs(todo) - 1;
                todo &= todo - 1;
                const IcpFrame f = a.icp_frames[__shfl_sync(LIP_FULL, b, p)];
                // dot-form scores only while the cancellation error stays far below the point spacing
                if (f.mabs <= 24.0f) icp_pass<true, 32, LIP_ICP_CHUNK, true>(a, sm.tf, sm.nn, &sm.st[p], lane, f.so, f.ns, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, sm.acc[p], f.axis, f.wb, it > 0);
                else icp_pass<false, 32, LIP_ICP_CHUNK, true>(a, sm.tf, sm.nn, &sm.st[p], lane, f.so, f.ns, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, sm.acc[p], f.axis, f.wb, it > 0);
            }
            __syncwarp();
            if (active) {                                       // SIMT over the task's partitions
                const double cost = sm.acc[lane][0] / own_ns;   // np.mean(distance)   (:86)
                iters = it + 1;
                if (it > 1 && (cost > prev + a.rise_tol || fabs(cost - prev) < a.flat_tol)) active = false;   // (:87)
                else {
                    prev = cost;
                    if (icp_kabsch_step(sm.acc[lane], &sm.st[lane])) status = 2;
                }
            }
            __syncwarp();
        }
        if (lane < g) {
            const double* ct = a.bin_sum + 4 * (size_t)pair_bin(a, b);
            icp_store(a, b, &sm.st[lane], ct[0], ct[1], ct[2], iters, status);
        }
        __syncwarp();
    }
}

// ---- CTA-per-partition variant for the heavy partitions (n_src * n_tgt > LIP_WARP_WORK: K == 1 whole-segment
// ICP, p2p_matching_seg_icp.py:440-449, direct my_icp calls, long leaves): same pass, the CTA's threads own the
// queries, the target streams through one SMEM tile, warp 0 does the serial part.
__device__ __forceinline__ double block_sum(double v, double* red, int tid) {   // 256 threads; all get the sum
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    return s;
}

__global__ void __launch_bounds__(LIP_BIG_THREADS, 4) icp_big_kernel(LipArgs a) {
    __shared__ float4 tf[LIP_BIG_CHUNK + 8];
    __shared__ double red17[LIP_BIG_THREADS / 32][17];
    __shared__ double tot[17];
    __shared__ int s_item, s_go, s_status;
    __shared__ IcpState st_sh;
    IcpState* st = &st_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int top = a.big_cluster_cls - 1;                  // classes above `top` belong to icp_cluster_kernel
    int n_big = 0;
    for (int k = 0; k <= top; ++k) n_big += a.sched[12 + k];
    for (;;) {                                              // persistent CTAs, heaviest class first
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&a.sched[16], 1);
        __syncthreads();
        const int wi = s_item;
        if (wi >= n_big) break;
        int cls = top, ri = wi;
        while (cls > 0 && ri >= a.sched[12 + cls]) { ri -= a.sched[12 + cls]; --cls; }
        const int b = a.big_list[(size_t)cls * a.B + ri];
        const IcpFrame f = a.icp_frames[b];
        if (warp == 0) {
            const double* ss = a.bin_sum + 4 * (size_t)b;
            const double* ts = a.bin_sum + 4 * (size_t)pair_bin(a, b);
            icp_state_init(st, ts[0] - ss[0], ts[1] - ss[1], ts[2] - ss[2]);
            if (lane == 0) { s_go = 1; s_status = 1; }
        }
        double prev = 0.0;                                   // warp 0 only
        int iters = 0;
        for (int it = 0; it < a.max_iter; ++it) {
            __syncthreads();
            if (!s_go) break;
            if (warp == 0) icp_state_accumulate(st);
            __syncthreads();
            if (f.mabs <= 24.0f) icp_pass<true, LIP_BIG_THREADS, LIP_BIG_CHUNK, false>(a, tf, nullptr, st, tid, f.so, f.ns, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, red17[warp], f.axis, f.wb, it > 0);
            else icp_pass<false, LIP_BIG_THREADS, LIP_BIG_CHUNK, false>(a, tf, nullptr, st, tid, f.so, f.ns, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, red17[warp], f.axis, f.wb, it > 0);
            __syncthreads();
            if (warp == 0) {                                 // serial part: all lanes of warp 0 identically
                if (lane < 17) {
                    double s = 0;
#pragma unroll
                    for (int k = 0; k < LIP_BIG_THREADS / 32; ++k) s += red17[k][lane];
                    tot[lane] = s;
                }
                __syncwarp();
                const double cost = tot[0] / f.ns;
                iters = it + 1;
                if (it > 1 && (cost > prev + a.rise_tol || fabs(cost - prev) < a.flat_tol)) { if (lane == 0) s_go = 0; }
                else {
                    prev = cost;
                    if (icp_kabsch_step(tot, st) && lane == 0) s_status = 2;
                }
            }
        }
        __syncthreads();
        if (tid == 0) icp_store(a, b, st, f.c0, f.c1, f.c2, iters, s_status);
    }
}

// ---- thread-block-cluster variant for the heaviest partitions (work classes >= a.big_cluster_cls: whole-organ ICPs of
// K == 1 segments, p2p_matching_seg_icp.py:440-449).  The serial chain of NN passes of ONE such partition is the critical
// path of the ICP stage of a small batch (one frame pair, as the reference's driver submits them) when it runs on the 128
// threads of an icp_big_kernel CTA.  Co-operative mode — as long as every heavy partition of the batch can have a cluster of
// its own at once: the CTAs of a cluster split the queries of a pass; every CTA stages the whole target, keeps its own copy of
// the ICP state and runs the identical serial part on the identical totals (each CTA adds the partial sums of all ranks, in
// rank order, out of their shared memory), so one cluster barrier per pass is the whole exchange.  With more heavy partitions
// than clusters, or in a batch of more than a.cluster_coop_parts partitions, the stage is throughput-bound, not latency-bound,
// and every CTA works for itself (one partition per CTA).
// Measured (profiles/README.md): one 979 x 1,058 ICP 292 -> 107 us, one 5,000 x 5,000 ICP 7.4 -> 1.1 ms.
template <int NT, int CHUNK>
__global__ void __launch_bounds__(NT, LIP_CLUSTER_MINBLOCKS) icp_cluster_kernel(LipArgs a) {
    namespace cg = cooperative_groups;
    __shared__ float4 tf[CHUNK + 8];
    __shared__ double red17[NT / 32][17];
    __shared__ double part[2][17];                          // this CTA's sums of the current pass (two buffers: see the barrier note)
    __shared__ double tot[17];
    __shared__ int s_item, s_go, s_status;
    __shared__ IcpState st_sh;
    static_assert(!LIP_ICP_SUM17 || (CHUNK + 8) * 16 >= (NT / 32) * 16 * 32 * 8, "search tile doubles as the 17-sum scratch");
    IcpState* st = &st_sh;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lo_cls = a.big_cluster_cls;
    int n_items = 0;
    for (int k = lo_cls; k < 4; ++k) n_items += a.sched[12 + k];
    if (n_items == 0) return;
    // uniform over the grid: co-operate only while no partition would have to wait for a cluster, and only in batches too small
    // to fill the GPU anyway — a co-operating cluster spends about six times the SM time of a lone CTA on the same partition
    // (107 us on 8 half-SMs against 292 us on a quarter), which a throughput-bound batch pays in full
    const bool coop = cluster.num_blocks() > 1 && (long long)n_items * cluster.num_blocks() <= (long long)gridDim.x &&
                      a.sched[0] + a.sched[1] <= a.cluster_coop_parts;
    const int rank = coop ? (int)cluster.block_rank() : 0, CL = coop ? (int)cluster.num_blocks() : 1;
    const int* item0 = coop ? cluster.map_shared_rank(&s_item, 0) : &s_item;
    for (;;) {                                              // persistent clusters / CTAs, heaviest class first
        __syncthreads();
        if (rank == 0 && tid == 0) s_item = atomicAdd(&a.sched[17], 1);
        if (coop) cluster.sync(); else __syncthreads();
        const int wi = *item0;
        if (coop) cluster.sync();                            // everybody has read it: rank 0 may draw again (or leave)
        if (wi >= n_items) break;
        int cls = 3, ri = wi;
        while (cls > lo_cls && ri >= a.sched[12 + cls]) { ri -= a.sched[12 + cls]; --cls; }
        const int b = a.big_list[(size_t)cls * a.B + ri];
        const IcpFrame f = a.icp_frames[b];
        // this CTA's queries: whole warps' worth, the ranks in order
        const int per = ((f.ns + 32 * CL - 1) / (32 * CL)) * 32;
        const int q_lo = min(f.ns, rank * per), q_n = min(f.ns, q_lo + per) - q_lo;
        if (warp == 0) {
            const double* ss = a.bin_sum + 4 * (size_t)b;
            const double* ts = a.bin_sum + 4 * (size_t)pair_bin(a, b);
            icp_state_init(st, ts[0] - ss[0], ts[1] - ss[1], ts[2] - ss[2]);
            if (lane == 0) { s_go = 1; s_status = 1; }
        }
        double prev = 0.0;                                   // warp 0 only
        int iters = 0;
        for (int it = 0; it < a.max_iter; ++it) {
            __syncthreads();
            if (!s_go) break;                                // identical in every CTA of the cluster
            if (warp == 0) icp_state_accumulate(st);
            __syncthreads();
            if (q_n > 0) {
                if (f.mabs <= 24.0f) icp_pass<true, NT, CHUNK, false>(a, tf, nullptr, st, tid, f.so + q_lo, q_n, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, red17[warp], f.axis, f.wb, it > 0);
                else icp_pass<false, NT, CHUNK, false>(a, tf, nullptr, st, tid, f.so + q_lo, q_n, f.to, f.nt, f.c0, f.c1, f.c2, f.sh0, f.sh1, f.sh2, f.mabs, red17[warp], f.axis, f.wb, it > 0);
            } else if (lane < 17) red17[warp][lane] = 0.0;
            __syncthreads();
            if (tid < 17) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < NT / 32; ++k) s += red17[k][tid];
                part[it & 1][tid] = s;
            }
            // One barrier per pass.  part[it & 1] is rewritten two passes (or one partition: two barriers) later, by when every
            // CTA has passed the next barrier and therefore finished reading it.
            if (coop) cluster.sync(); else __syncthreads();
            if (warp == 0) {                                 // serial part: all lanes of warp 0 of every CTA identically
                if (lane < 17) {
                    double s = 0;
                    for (int r = 0; r < CL; ++r) s += (coop ? cluster.map_shared_rank(&part[it & 1][0], r) : &part[it & 1][0])[lane];
                    tot[lane] = s;
                }
                __syncwarp();
                const double cost = tot[0] / f.ns;
                iters = it + 1;
                if (it > 1 && (cost > prev + a.rise_tol || fabs(cost - prev) < a.flat_tol)) { if (lane == 0) s_go = 0; }
                else {
                    prev = cost;
                    if (icp_kabsch_step(tot, st) && lane == 0) s_status = 2;
                }
            }
        }
        __syncthreads();
        if (rank == 0 && tid == 0) icp_store(a, b, st, f.c0, f.c1, f.c2, iters, s_status);
    }
}

// =============================================================================================
// fill of missing transforms (p2p_matching_seg_icp.py:422-437): one thread per source mesh
// =============================================================================================
__global__ void fill_kernel(LipArgs a) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.M || !src_active(a, m)) return;
    const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
    for (int i = 0; i < K; ++i) {
        const int b = bb + i;
        int from = -1;                                       // -1: identity + mean shift
        bool own = a.icp_status[b] != 0;
        if (!own) {
            if (i > 0) from = b - 1;                         // already filled (ascending pass)
            else if (K > 1 && a.icp_status[b + 1] != 0) from = b + 1;
        }
        if (own) {
            for (int k = 0; k < 9; ++k) a.fill_R[9 * (size_t)b + k] = a.icp_R[9 * (size_t)b + k];
            for (int k = 0; k < 3; ++k) a.fill_t[3 * (size_t)b + k] = a.icp_t[3 * (size_t)b + k];
        } else if (from == b - 1) {
            for (int k = 0; k < 9; ++k) a.fill_R[9 * (size_t)b + k] = a.fill_R[9 * (size_t)from + k];
            for (int k = 0; k < 3; ++k) a.fill_t[3 * (size_t)b + k] = a.fill_t[3 * (size_t)from + k];
        } else if (from == b + 1) {
            for (int k = 0; k < 9; ++k) a.fill_R[9 * (size_t)b + k] = a.icp_R[9 * (size_t)from + k];
            for (int k = 0; k < 3; ++k) a.fill_t[3 * (size_t)b + k] = a.icp_t[3 * (size_t)from + k];
        } else {
            const double* hs = a.mesh_sum + 4 * (size_t)m;
            const double* ht = a.mesh_sum + 4 * (size_t)pair_mesh(a, m);
            for (int k = 0; k < 9; ++k) a.fill_R[9 * (size_t)b + k] = (k % 4 == 0) ? 1.0 : 0.0;
            for (int k = 0; k < 3; ++k) a.fill_t[3 * (size_t)b + k] = ht[k] / ht[3] - hs[k] / hs[3];
        }
    }
}

// =============================================================================================
// K5  uniform-grid spatial hash of every mesh (replaces KDTree(xyz1), p2p_matching_seg_icp.py:452):
//     bbox + vertex sum, cell size from a sampled nearest-neighbour spacing, cell hash -> counting
//     sort (histogram, prefix scan, scatter) into cell-ordered float4 points.  One CTA per mesh.
// =============================================================================================
__device__ __forceinline__ int cell_hash(int cx, int cy, int cz, int mask) {
    return (int)(((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u)) & mask;
}
__device__ __forceinline__ int cell_coord(float x, float g, float inv_h, int dim) {
    float c = floorf((x - g) * inv_h);
    c = fminf(fmaxf(c, -4.0f), (float)(dim + 3));
    return (int)c;
}

__global__ void __launch_bounds__(256) grid_build_kernel(LipArgs a, int need_all) {
    __shared__ double redd[8][9];
    __shared__ float samp[32];
    __shared__ GridHdr hdr;
    __shared__ int carry;
    __shared__ int wsum[8];
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // only targets of a requested direction need a grid, but every mesh needs hdr.sum / hdr.n
    const bool want_grid = need_all || src_active(a, pair_mesh(a, m));
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const double* xyz = mesh_xyz(a, m);
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, sm[3] = {0, 0, 0};
    for (int v = tid; v < n; v += 256) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double x = xyz[3 * (size_t)v + k];
            mn[k] = fmin(mn[k], x); mx[k] = fmax(mx[k], x); sm[k] += x;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); sm[k] = warp_sum(sm[k]); }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { redd[warp][k] = mn[k]; redd[warp][3 + k] = mx[k]; redd[warp][6 + k] = sm[k]; }
    }
    __syncthreads();
    if (tid == 0) {
        float ext = 0.f;
        for (int k = 0; k < 3; ++k) {
            double lo = redd[0][k], hi = redd[0][3 + k], s = redd[0][6 + k];
            for (int w = 1; w < 8; ++w) { lo = fmin(lo, redd[w][k]); hi = fmax(hi, redd[w][3 + k]); s += redd[w][6 + k]; }
            hdr.c[k] = 0.5 * (lo + hi);
            hdr.sum[k] = s;
            hdr.gmin[k] = (float)(lo - hdr.c[k]);
            ext = fmaxf(ext, (float)(hi - lo));
        }
        hdr.mabs = 0.5f * ext * 1.0001f + 1e-30f;
        hdr.h = ext;                                         // provisional: extent, finalised below
        hdr.n = n;
        hdr.mask = a.mesh_tab[m + 1] - a.mesh_tab[m] - 3;
    }
    __syncthreads();
    if (!want_grid) {
        if (tid == 0) a.grid_hdr[m] = hdr;
        return;
    }
    const double c0 = hdr.c[0], c1 = hdr.c[1], c2 = hdr.c[2];
    // sampled nearest-neighbour spacing -> cell size h = 4 * mean spacing (about 4 points per cell on a surface)
    const int nsamp = min(32, n);
    for (int k = warp; k < nsamp; k += 8) {
        const int sv = (int)(((long long)k * n) / nsamp);
        const float qx = (float)(xyz[3 * (size_t)sv] - c0), qy = (float)(xyz[3 * (size_t)sv + 1] - c1),
                    qz = (float)(xyz[3 * (size_t)sv + 2] - c2);
        float best = INFINITY;
        for (int j = lane; j < n; j += 32) {
            float dx = (float)(xyz[3 * (size_t)j] - c0) - qx, dy = (float)(xyz[3 * (size_t)j + 1] - c1) - qy,
                  dz = (float)(xyz[3 * (size_t)j + 2] - c2) - qz;
            float d = dx * dx + dy * dy + dz * dz;
            if (j != sv) best = fminf(best, d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(LIP_FULL, best, o));
        if (lane == 0) samp[k] = best;
    }
    __syncthreads();
    if (tid == 0) {
        float ext = hdr.h, s = 0.f;
        int cnt = 0;
        for (int k = 0; k < nsamp; ++k)
            if (samp[k] < INFINITY) { s += sqrtf(samp[k]); ++cnt; }
        float h = (cnt > 0) ? 4.0f * s / cnt : 1.0f;
        h = fmaxf(h, ext * (1.0f / 1000.0f));
        if (!(h > 1e-30f)) h = 1.0f;
        hdr.h = h;
        hdr.inv_h = 1.0f / h;
        for (int k = 0; k < 3; ++k) hdr.dim[k] = 0;          // filled from the points below (max cell + 1)
        a.grid_hdr[m] = hdr;
    }
    __syncthreads();
    const float inv_h = hdr.inv_h, g0 = hdr.gmin[0], g1 = hdr.gmin[1], g2 = hdr.gmin[2];
    const int mask = hdr.mask;
    int* tab = a.grid_start + a.mesh_tab[m];                 // mask+3 entries, zeroed by the host
    int dmax[3] = {0, 0, 0};
    for (int v = tid; v < n; v += 256) {
        float x = (float)(xyz[3 * (size_t)v] - c0), y = (float)(xyz[3 * (size_t)v + 1] - c1), z = (float)(xyz[3 * (size_t)v + 2] - c2);
        int cx = max(cell_coord(x, g0, inv_h, 1 << 20), 0), cy = max(cell_coord(y, g1, inv_h, 1 << 20), 0),
            cz = max(cell_coord(z, g2, inv_h, 1 << 20), 0);
        dmax[0] = max(dmax[0], cx); dmax[1] = max(dmax[1], cy); dmax[2] = max(dmax[2], cz);
        atomicAdd(&tab[cell_hash(cx, cy, cz, mask) + 2], 1);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax[k] = max(dmax[k], __shfl_xor_sync(LIP_FULL, dmax[k], o));
        if (lane == 0) atomicMax(&a.grid_hdr[m].dim[k], dmax[k] + 1);
    }
    if (tid == 0) carry = 0;
    __syncthreads();
    // inclusive scan of tab[0 .. mask+2] in place: afterwards tab[b+1] = first slot of bucket b
    const int total = mask + 3;
    for (int base = 0; base < total; base += 256) {
        int i = base + tid;
        int c = (i < total) ? tab[i] : 0, inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(LIP_FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (i < total) tab[i] = inc + off;
        __syncthreads();
        if (tid == 255) carry = inc + off;
        __syncthreads();
    }
    // scatter: atomicAdd on tab[b+1] hands out the slots; afterwards tab[b] .. tab[b+1] brackets bucket b
    float4* pts = a.grid_pts + ub;
    for (int v = tid; v < n; v += 256) {
        float x = (float)(xyz[3 * (size_t)v] - c0), y = (float)(xyz[3 * (size_t)v + 1] - c1), z = (float)(xyz[3 * (size_t)v + 2] - c2);
        int cx = max(cell_coord(x, g0, inv_h, 1 << 20), 0), cy = max(cell_coord(y, g1, inv_h, 1 << 20), 0),
            cz = max(cell_coord(z, g2, inv_h, 1 << 20), 0);
        int slot = atomicAdd(&tab[cell_hash(cx, cy, cz, mask) + 1], 1);
        pts[slot] = make_float4(x, y, z, __int_as_float(v));
    }
}

// ---------------------------------------------------------------------------------------------
// exact nearest neighbour of one query in a hashed grid: ring 1 (27 cells), ring 2 (98 more), then a
// full scan; a ring result is accepted only if the best distance is inside the searched block.
// ---------------------------------------------------------------------------------------------
struct NNResult { int idx; int level; };

__device__ __forceinline__ void grid_visit32(const int* __restrict__ tab, const float4* __restrict__ pts, int bucket,
                                             float qx, float qy, float qz, float& best, float& sec, int& idx,
                                             int& visited) {
    const int s = __ldg(&tab[bucket]), e = __ldg(&tab[bucket + 1]);
    visited += e - s;
    for (int j = s; j < e; ++j) {
        const float4 p = __ldg(&pts[j]);
        float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
        float d = dx * dx; d = fmaf(dy, dy, d); d = fmaf(dz, dz, d);
        nn_update(d, __float_as_int(p.w), best, sec, idx);
    }
}
__device__ __forceinline__ void grid_visit64(const int* __restrict__ tab, const float4* __restrict__ pts,
                                             const double* __restrict__ xyz, int bucket, double qx, double qy, double qz,
                                             double& best, int& idx) {
    const int s = __ldg(&tab[bucket]), e = __ldg(&tab[bucket + 1]);
    for (int j = s; j < e; ++j) {
        const int v = __float_as_int(__ldg(&pts[j]).w);
        double d = rdist64(qx, qy, qz, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
        if (d < best || (d == best && v < idx)) { best = d; idx = v; }
    }
}

// scan the cells at Chebyshev distance r_lo..r_hi from (cx,cy,cz)
__device__ __forceinline__ void grid_ring_scan32(const GridHdr& h, const int* __restrict__ tab, const float4* __restrict__ pts,
                                                 int cx, int cy, int cz, int r_lo, int r_hi, float qx, float qy, float qz,
                                                 float& best, float& sec, int& idx, int& visited) {
    for (int dz = -r_hi; dz <= r_hi; ++dz) {
        const int z = cz + dz;
        if (z < 0 || z >= h.dim[2]) continue;
        for (int dy = -r_hi; dy <= r_hi; ++dy) {
            const int y = cy + dy;
            if (y < 0 || y >= h.dim[1]) continue;
            for (int dx = -r_hi; dx <= r_hi; ++dx) {
                const int x = cx + dx;
                if (x < 0 || x >= h.dim[0]) continue;
                if (max(max(abs(dx), abs(dy)), abs(dz)) < r_lo) continue;
                grid_visit32(tab, pts, cell_hash(x, y, z, h.mask), qx, qy, qz, best, sec, idx, visited);
            }
        }
    }
}

// Given the fp32 result of a search that covered every cell within `level` rings of the query's cell and
// found best <= (0.999 level h)^2: resolve near ties in float64 on the original coordinates and return
// the index of the exact nearest neighbour (lowest index on an exact tie).
__device__ __forceinline__ int grid_finish(const GridHdr& h, const int* __restrict__ tab, const float4* __restrict__ pts,
                                           const double* __restrict__ xyz, double q0, double q1, double q2, float qx,
                                           float qy, float qz, int cx, int cy, int cz, int level, float best, float sec,
                                           int idx, unsigned long long* counters) {
    const float g = 0.999f * h.h * (float)level;
    // anything outside the searched block is at least g away: it competes with `sec`
    sec = fminf(sec, g * g);
    const float e = 1.2e-7f * (fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + h.mabs);
    if (sec - best <= d2_err(best, e) + d2_err(sec, e)) {
        // near tie (or best close to the certified radius): decide in float64
        atomicAdd(&counters[3], 1ull);
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int dz = -level; dz <= level; ++dz) {
            const int z = cz + dz;
            if (z < 0 || z >= h.dim[2]) continue;
            for (int dy = -level; dy <= level; ++dy) {
                const int y = cy + dy;
                if (y < 0 || y >= h.dim[1]) continue;
                for (int dx = -level; dx <= level; ++dx) {
                    const int x = cx + dx;
                    if (x < 0 || x >= h.dim[0]) continue;
                    grid_visit64(tab, pts, xyz, cell_hash(x, y, z, h.mask), q0, q1, q2, bd, bi);
                }
            }
        }
        const double gg = (double)g * 0.999;
        if (!(bd <= gg * gg)) {                              // not certifiable in float64 either: full scan
            bd = INFINITY; bi = 0x7fffffff;
            for (int v = 0; v < h.n; ++v) {
                double d = rdist64(q0, q1, q2, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
                if (d < bd) { bd = d; bi = v; }
            }
        }
        idx = bi;
    }
    return idx;
}

// One thread, one query (ORIGINAL float64 coordinates): ring 1 (27 cells), then ring 2 (98 more); a ring
// result is accepted only if the best distance is inside the searched block.  idx = -1: not certifiable
// within two rings — the caller hands the query to far_nn_kernel.
__device__ __forceinline__ NNResult grid_nearest(const GridHdr& h, const int* __restrict__ tab,
                                                 const float4* __restrict__ pts, const double* __restrict__ xyz,
                                                 double q0, double q1, double q2, unsigned long long* counters,
                                                 double* dist_out, int profile) {
    const float qx = (float)(q0 - h.c[0]), qy = (float)(q1 - h.c[1]), qz = (float)(q2 - h.c[2]);
    const int cx = cell_coord(qx, h.gmin[0], h.inv_h, h.dim[0]), cy = cell_coord(qy, h.gmin[1], h.inv_h, h.dim[1]),
              cz = cell_coord(qz, h.gmin[2], h.inv_h, h.dim[2]);
    float best = INFINITY, sec = INFINITY;
    int idx = -1, level = 1, visited = 0;
    grid_ring_scan32(h, tab, pts, cx, cy, cz, 0, 1, qx, qy, qz, best, sec, idx, visited);
    if (profile) {                                           // ring-1 candidates = the algorithmic pair count
        unsigned tot = __reduce_add_sync(__activemask(), (unsigned)visited);
        if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicAdd(&counters[7], (unsigned long long)tot);
    }
    float g = 0.999f * h.h;
    if (!(best <= g * g)) {
        level = 2;
        atomicAdd(&counters[4], 1ull);
        grid_ring_scan32(h, tab, pts, cx, cy, cz, 2, 2, qx, qy, qz, best, sec, idx, visited);
        g = 2.0f * 0.999f * h.h;
        if (!(best <= g * g)) {
            NNResult r; r.idx = -1; r.level = 3;
            return r;
        }
    }
    idx = grid_finish(h, tab, pts, xyz, q0, q1, q2, qx, qy, qz, cx, cy, cz, level, best, sec, idx, counters);
    if (dist_out) {
        *dist_out = sqrt(rdist64(q0, q1, q2, xyz[3 * (size_t)idx], xyz[3 * (size_t)idx + 1], xyz[3 * (size_t)idx + 2]));
    }
    NNResult r; r.idx = idx; r.level = level;
    return r;
}

__device__ __forceinline__ void far_push(const LipArgs& a, int u, int mt, double q0, double q1, double q2) {
    const int slot = atomicAdd(&a.sched[3], 1);
    a.far_list[2 * (size_t)slot] = u; a.far_list[2 * (size_t)slot + 1] = mt;
    a.far_q[3 * (size_t)slot] = q0; a.far_q[3 * (size_t)slot + 1] = q1; a.far_q[3 * (size_t)slot + 2] = q2;
}
__device__ __forceinline__ void retry_push(const LipArgs& a, int u, int mt, double q0, double q1, double q2) {
    const int slot = atomicAdd(&a.sched[7], 1);
    a.retry_list[2 * (size_t)slot] = u; a.retry_list[2 * (size_t)slot + 1] = mt;
    a.retry_q[3 * (size_t)slot] = q0; a.retry_q[3 * (size_t)slot + 1] = q1; a.retry_q[3 * (size_t)slot + 2] = q2;
}

// Which ICP result does landmark i of a mesh use after the fill of p2p_matching_seg_icp.py:422-437?
// Ascending pass: a missing entry takes entry i-1 (already filled), entry 0 takes entry 1 if that ran,
// otherwise identity + mean shift.  Returns the bin whose raw (R,t) applies, or -1 for the default.
__device__ __forceinline__ int fill_source(const LipArgs& a, int bb, int K, int i) {
    int j = i;
    while (j > 0 && a.icp_status[bb + j] == 0) --j;
    if (a.icp_status[bb + j] != 0) return bb + j;
    return (K > 1 && a.icp_status[bb + 1] != 0) ? bb + 1 : -1;
}
__device__ __forceinline__ void fill_apply(const LipArgs& a, int m, int src, double x, double y, double z, double& ox,
                                           double& oy, double& oz) {
    if (src >= 0) {
        const double* R = a.icp_R + 9 * (size_t)src;
        const double* t = a.icp_t + 3 * (size_t)src;
        mat3_vec(R, x, y, z, ox, oy, oz);
        ox += t[0]; oy += t[1]; oz += t[2];
    } else {                                                 // identity, t = mean(target mesh) - mean(source mesh)
        const double* hs = a.mesh_sum + 4 * (size_t)m;
        const double* ht = a.mesh_sum + 4 * (size_t)pair_mesh(a, m);
        ox = x + (ht[0] / ht[3] - hs[0] / hs[3]);
        oy = y + (ht[1] / ht[3] - hs[1] / hs[3]);
        oz = z + (ht[2] / ht[3] - hs[2] / hs[3]);
    }
}

// The blended position of vertex v of source mesh m (p2p_matching_seg_icp.py:454-487): weights from the
// triangle (landmark gap, d_a, d_b) — get_triangle_angles :27-34 fused, cos(arccos(x)) == x — and the two
// landmark transforms after the fill of :422-437, all in float64.
__device__ __forceinline__ void blend_point(const LipArgs& a, int m, int bb, int K, double x, double y, double z, int ia, int ib,
                                            double da, double db, double& q0, double& q1, double& q2) {
    double f0 = 0.5, f1 = 0.5;
    if (K > 1) {
        const double* lm = mesh_lm(a, m);
        const double gx = lm[3 * ib] - lm[3 * ia], gy = lm[3 * ib + 1] - lm[3 * ia + 1], gz = lm[3 * ib + 2] - lm[3 * ia + 2];
        const double g = sqrt(gx * gx + gy * gy + gz * gz);  // |lm[b] - lm[a]|            (:467)
        // get_triangle_angles(g, da, db): B opposite da, C opposite db
        double cosB = (db * db + g * g - da * da) / (2.0 * g * db + a.cos_den_eps);
        double cosC = (g * g + da * da - db * db) / (2.0 * g * da + a.cos_den_eps);
        cosB = fmin(fmax(cosB, -1.0), 1.0);
        cosC = fmin(fmax(cosC, -1.0), 1.0);
        if (cosC < a.obtuse_cos) { f0 = 1.0; f1 = 0.0; }     // C > pi/2 - 1e-5                (:471)
        else {
            const double cb = fabs(cosB), cc = fabs(cosC);
            f0 = cb / (cb + cc); f1 = cc / (cb + cc);
        }
    }
    double ax, ay, az, bx, by, bz;
    fill_apply(a, m, fill_source(a, bb, K, ia), x, y, z, ax, ay, az);
    if (ib == ia) { bx = ax; by = ay; bz = az; }
    else fill_apply(a, m, fill_source(a, bb, K, ib), x, y, z, bx, by, bz);
    q0 = f0 * ax + f1 * bx; q1 = f0 * ay + f1 * by; q2 = f0 * az + f1 * bz;
}

__device__ __forceinline__ void blend_query(const LipArgs& a, int m, int bb, int K, int v, double& q0, double& q1, double& q2) {
    const int u = a.mesh_ubase[m] + v;
    const double* xyz = mesh_xyz(a, m);
    blend_point(a, m, bb, K, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2], a.nn0[u], a.nn1[u], a.d0[u], a.d1[u],
                q0, q1, q2);
}

// One thread per landmark bin, after the ICPs: everything blend_kernel needs of a bin in ONE 128-byte record — the
// transform of the bin after the fill of p2p_matching_seg_icp.py:422-437 (an identity rotation with the mean shift as
// translation when no ICP of the mesh ran), its landmark, the first bin and the landmark count of its mesh — so that a
// vertex reaches its two transforms in two dependent loads instead of six.
__global__ void __launch_bounds__(128) bin_desc_kernel(LipArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    double* D = a.bin_desc + 16 * (size_t)b;
    const int m = a.bin_mesh[b];
    int2 meta = make_int2(0, 0);
    if (src_active(a, m)) {
        const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
        const int src = fill_source(a, bb, K, b - bb);
        if (src >= 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) D[i] = a.icp_R[9 * (size_t)src + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) D[9 + i] = a.icp_t[3 * (size_t)src + i];
        } else {                                             // identity, t = mean(target mesh) - mean(source mesh)
            const double* hs = a.mesh_sum + 4 * (size_t)m;
            const double* ht = a.mesh_sum + 4 * (size_t)pair_mesh(a, m);
#pragma unroll
            for (int i = 0; i < 9; ++i) D[i] = (i % 4 == 0) ? 1.0 : 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) D[9 + i] = ht[i] / ht[3] - hs[i] / hs[3];
        }
        const double* lm = mesh_lm(a, m) + 3 * (size_t)(b - bb);
        D[12] = lm[0]; D[13] = lm[1]; D[14] = lm[2];
        meta = make_int2(bb, K);
    }
    *reinterpret_cast<int2*>(D + 15) = meta;
}

// One thread per source vertex (bin order): the blended position of p2p_matching_seg_icp.py:454-487 in float64,
// written as planes x | y | z like bin_xyz.  Streaming and fully parallel, so that the latency of its dependent
// loads and float64 divisions is not paid inside the search tiles.  Same arithmetic as blend_point, on the bin records.
__global__ void __launch_bounds__(256) blend_kernel(LipArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_total) return;
    const int b = a.pos_bin[p];                              // everything per vertex is read in bin order: coalesced
    const double* __restrict__ Da = a.bin_desc + 16 * (size_t)b;
    const int2 meta = *reinterpret_cast<const int2*>(Da + 15);
    const int bb = meta.x, K = meta.y;
    if (K == 0) return;                                      // not a source of a requested direction
    const size_t N = (size_t)a.n_total;
    const double x = a.bin_xyz[p], y = a.bin_xyz[N + p], z = a.bin_xyz[2 * N + p];
    const int ia = b - bb, ib = a.pos_nn1[p];
    const double da = a.pos_d[p], db = a.pos_d[N + p];
    const double* __restrict__ Db = a.bin_desc + 16 * (size_t)(bb + ib);
    double f0 = 0.5, f1 = 0.5;
    if (K > 1) {
        const double gx = Db[12] - Da[12], gy = Db[13] - Da[13], gz = Db[14] - Da[14];
        const double g = sqrt(gx * gx + gy * gy + gz * gz);  // |lm[b] - lm[a]|            (:467)
        // get_triangle_angles(g, da, db): B opposite da, C opposite db
        double cosB = (db * db + g * g - da * da) / (2.0 * g * db + a.cos_den_eps);
        double cosC = (g * g + da * da - db * db) / (2.0 * g * da + a.cos_den_eps);
        cosB = fmin(fmax(cosB, -1.0), 1.0);
        cosC = fmin(fmax(cosC, -1.0), 1.0);
        if (cosC < a.obtuse_cos) { f0 = 1.0; f1 = 0.0; }     // C > pi/2 - 1e-5                (:471)
        else {
            const double cb = fabs(cosB), cc = fabs(cosC);
            f0 = cb / (cb + cc); f1 = cc / (cb + cc);
        }
    }
    double ax, ay, az, bx, by, bz;
    mat3_vec(Da, x, y, z, ax, ay, az);
    ax += Da[9]; ay += Da[10]; az += Da[11];
    if (ib == ia) { bx = ax; by = ay; bz = az; }
    else { mat3_vec(Db, x, y, z, bx, by, bz); bx += Db[9]; by += Db[10]; bz += Db[11]; }
    a.bin_q[p] = f0 * ax + f1 * bx; a.bin_q[N + p] = f0 * ay + f1 * by; a.bin_q[2 * N + p] = f0 * az + f1 * bz;
}

// =============================================================================================
// K4  per-point two-transform blend + final nearest neighbour in the whole target mesh
//     (p2p_matching_seg_icp.py:452-489; get_triangle_angles :27-34 fused, cos(arccos(x)) == x;
//     the fill of :422-437 is resolved on the fly).
//     Queries are taken in partition order (part_vid): the threads of a CTA work on one landmark blob,
//     so the target cells and points they touch stay in L1.  Each thread first issues the 27 cell
//     look-ups of its ring (independent loads), then walks the candidates of the non-empty cells in one
//     flattened loop; queries that ring 1 cannot certify go to a compacted retry list.
// =============================================================================================
__global__ void __launch_bounds__(LIP_TILE) blend_nn_kernel(LipArgs a) {
    const int m = a.tile_mesh[blockIdx.x];
    const int slot = a.tile_first[blockIdx.x] + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const bool valid = slot < n;
    const int bb = a.mesh_bin[m], K = a.mesh_bin[m + 1] - bb;
    const int v = valid ? a.part_vid[ub + slot] : 0;
    const int u = ub + v;
    const int mt = pair_mesh(a, m);
    double q0 = 0, q1 = 0, q2 = 0;
    if (valid) blend_query(a, m, bb, K, v, q0, q1, q2);
    if (!valid) return;
    // ---- ring-1 search in the target's hashed grid ------------------------------------------------------
    const GridHdr& h = a.grid_hdr[mt];
    const int* __restrict__ tab = a.grid_start + a.mesh_tab[mt];
    const float4* __restrict__ pts = a.grid_pts + a.mesh_ubase[mt];
    const float qx = (float)(q0 - h.c[0]), qy = (float)(q1 - h.c[1]), qz = (float)(q2 - h.c[2]);
    const int d0 = h.dim[0], d1 = h.dim[1], d2 = h.dim[2], mask = h.mask;
    const int cx = cell_coord(qx, h.gmin[0], h.inv_h, d0), cy = cell_coord(qy, h.gmin[1], h.inv_h, d1),
              cz = cell_coord(qz, h.gmin[2], h.inv_h, d2);
    // pass A: the 27 table look-ups, branch-free and independent (all loads in flight together);
    // non-empty cells are compacted into a per-thread list
    int rs[27], re[27];
    int nc = 0;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const int x = cx + (k % 3) - 1, y = cy + ((k / 3) % 3) - 1, z = cz + (k / 9) - 1;
        const bool in = (unsigned)x < (unsigned)d0 && (unsigned)y < (unsigned)d1 && (unsigned)z < (unsigned)d2;
        const int bucket = in ? cell_hash(x, y, z, mask) : 0;
        const int s = __ldg(&tab[bucket]), e = __ldg(&tab[bucket + 1]);
        if (in && e > s) { rs[nc] = s; re[nc] = e; ++nc; }
    }
    // pass B: the non-empty cells, four candidates per step under predication (a cell holds ~4 points);
    // the only divergent branch is the step to the next cell
    float best = INFINITY, sec = INFINITY;
    int idx = -1, visited = 0;
    if (nc > 0) {
        int c = 0, j = rs[0], e = re[0];
        visited = e - j;
        for (;;) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (j + k < e) {
                    const float4 p = __ldg(&pts[j + k]);
                    float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
                    float d = dx * dx; d = fmaf(dy, dy, d); d = fmaf(dz, dz, d);
                    nn_update(d, __float_as_int(p.w), best, sec, idx);
                }
            }
            j += 4;
            if (j >= e) {
                if (++c == nc) break;
                j = rs[c]; e = re[c];
                visited += e - j;
            }
        }
    }
    if (a.profile) {                                         // ring-1 candidates = the algorithmic pair count
        const unsigned act = __activemask();
        const unsigned tot = __reduce_add_sync(act, (unsigned)visited);
        if (lane == __ffs(act) - 1) atomicAdd(&a.counters[7], (unsigned long long)tot);
    }
    const float g = 0.999f * h.h;
    if (best <= g * g) {                                     // certified: nothing outside ring 1 can be closer
        idx = grid_finish(h, tab, pts, mesh_xyz(a, mt), q0, q1, q2, qx, qy, qz, cx, cy, cz, 1, best, sec, idx, a.counters);
        store_map(a, u, idx);
    } else {
        retry_push(a, u, mt, q0, q1, q2);                    // ring 2 / far: second pass over the compacted list
    }
}

// =============================================================================================
// K4' final map on landmark-bin tiles (default; the hashed-grid variant above is kept for A/B evidence)
//     The partition step has already sorted every mesh by nearest landmark, so the target mesh comes as
//     K spatially compact bins with known bounding boxes — that IS the spatial index.  A CTA takes up to
//     LIP_MAP_TILE queries of ONE source bin (icp_schedule_kernel builds the tile list), stages the target
//     bin of the same landmark into SMEM and scans it (round 0); then every target bin whose box is closer
//     to some query than that query's current best is staged and scanned too (round 1).  After round 1
//     every query has seen every point that could be its nearest neighbour, so the result is exact without
//     any fallback path.  All lanes read the same candidate (SMEM broadcast); the per-pair cost is 3 FFMA
//     + ~1.1 min/max (scores c = |y|^2 - 2 q.y, block minima, winner recovered per block of 8).
// =============================================================================================
struct MapSmem {
    float4 tp[LIP_MAP_CHUNK];
    int tv[LIP_MAP_CHUNK];
    unsigned need[LIP_MAX_LANDMARKS * LIP_MAP_SUB / 32];              // one bit per (bin, range): staged in round 1
    float sbox[LIP_MAP_BOXB][8];                                  // box of every bin of the batch (tile frame), max |coordinate|, 0 / 1 + sort axis
    alignas(16) float srange[LIP_MAP_BOXB][2 * LIP_MAP_SUB];                  // ... and the extents of its ranges along that axis (lo[], hi[])
    float red[LIP_MAP_THREADS / 32][8];
    unsigned cand[2];
    struct { int start, end; float sh[3]; } sdesc[LIP_MAP_BOXB];   // where the bins of the last box batch are, and their shift
};

__device__ __forceinline__ void blend_tile(const LipArgs& a, MapSmem& S, int tile) {
    float4* tp = S.tp;
    int* tv = S.tv;
    unsigned* need = S.need;
    float (*sbox)[8] = S.sbox;
    float (*red)[8] = S.red;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const MapTile T = a.map_tiles[tile];                     // one 80-byte descriptor instead of a chain of dependent look-ups
    const int first = T.first, m = T.m, ub = T.ub, bb = T.bb, K = T.K, bbt = T.bbt, so = T.so, nq = T.nq;
    const int mt = pair_mesh(a, m);
    const int home = T.b - bb;                               // the queries' nearest landmark (local index)
    const double* __restrict__ lmt = mesh_lm(a, mt);
    // tile origin: the target's landmark (the home bin's stored offsets need no shift) — unless that bin is empty: the queries of
    // such a tile lie far from everything they will be compared with, and the cancellation error of the dot-form scores (which
    // grows with |q|^2 about the origin) would send most of them to the float64 rescan.  Their own first query is the origin then.
    const bool orphan = T.te <= T.ts;
    const size_t pq0 = (size_t)so + first;
    const double C0 = orphan ? a.bin_q[pq0] : T.C[0], C1 = orphan ? a.bin_q[(size_t)a.n_total + pq0] : T.C[1],
                 C2 = orphan ? a.bin_q[2 * (size_t)a.n_total + pq0] : T.C[2];
    constexpr int Q = LIP_MAP_Q;
    float mq[Q][3], qq[Q], qabs[Q], best[Q], sec[Q], sec_in[Q];
    int idx[Q], blk[Q], vq[Q];
    bool valid[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const int slot = first + tid + k * LIP_MAP_THREADS;
        valid[k] = tid + k * LIP_MAP_THREADS < nq;
        vq[k] = valid[k] ? a.part_vid[so + slot] : 0;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (valid[k]) {                                      // blended positions in bin order, written by blend_kernel
            const size_t N = (size_t)a.n_total, p = (size_t)so + slot;
            fx = (float)(a.bin_q[p] - C0); fy = (float)(a.bin_q[N + p] - C1); fz = (float)(a.bin_q[2 * N + p] - C2);
        }
        mq[k][0] = -2.f * fx; mq[k][1] = -2.f * fy; mq[k][2] = -2.f * fz;
        qq[k] = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
        qabs[k] = fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz));
        best[k] = INFINITY; sec[k] = INFINITY; sec_in[k] = INFINITY; idx[k] = -1; blk[k] = 0;
    }
    // queries in use by this WARP (units of 32 queries are dealt warp 0, warp 1, warp 0, warp 1)
    const int qw = min(Q, max(0, (nq - 32 * warp + LIP_MAP_THREADS - 1) / LIP_MAP_THREADS));
    __syncthreads();                                         // the previous tile of this CTA is done with the shared arrays
    const int nwords = (K * LIP_MAP_SUB + 31) / 32;
    for (int i = tid; i < nwords; i += LIP_MAP_THREADS) need[i] = 0u;
    float mabs = 0.f;                                        // max |coordinate| staged so far (uniform)
    int fill = 0;
    long long staged = 0;

    // Two rounds through ONE staging / scanning loop (a single copy of the scan code keeps the kernel small):
    // round 0 stages the target bin of the queries' own landmark, round 1 every other bin that could hold
    // something closer.
    for (int round = 0; round < 2; ++round) {
        if (round == 1) {
            if (K <= 1) break;
            // CTA-wide bounds: box of the queries, largest current best distance, largest |q|
            float v[8];
            v[6] = 0.f; v[7] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { v[c] = INFINITY; v[3 + c] = -INFINITY; }
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                if (!valid[k]) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) { const float f = -0.5f * mq[k][c]; v[c] = fminf(v[c], f); v[3 + c] = fmaxf(v[3 + c], f); }
                v[6] = fmaxf(v[6], fmaxf(best[k] + qq[k], 0.f));
                v[7] = fmaxf(v[7], qabs[k]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = -warp_maxf(-v[c]);
#pragma unroll
            for (int c = 3; c < 8; ++c) v[c] = warp_maxf(v[c]);
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) red[warp][c] = v[c];
            }
            __syncthreads();
#pragma unroll
            for (int w = 0; w < LIP_MAP_THREADS / 32; ++w) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = fminf(v[c], red[w][c]);
#pragma unroll
                for (int c = 3; c < 8; ++c) v[c] = fmaxf(v[c], red[w][c]);
            }
            const float rmax2 = v[6], qabs_max = v[7];
            const float qq_max = 3.f * qabs_max * qabs_max;
            for (int j0 = 0; j0 < K; j0 += LIP_MAP_BOXB) {
                __syncthreads();
                if (tid < LIP_MAP_BOXB) {                    // boxes of this batch of bins in the tile frame + CTA-level cull
                    const int j = j0 + tid;
                    float flag = 0.f, lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f}, mb = 0.f;
                    const int js = j < K ? a.bin_off[bbt + j] : 0, je = j < K ? a.bin_off[bbt + j + 1] : 0;
                    S.sdesc[tid].start = js; S.sdesc[tid].end = je;
                    if (j < K && j != home && je > js) {
                        const float4* bx4 = reinterpret_cast<const float4*>(a.bin_box + 8 * (size_t)(bbt + j));
                        const float4 b0 = bx4[0], b1 = bx4[1];               // box, max |coordinate|, axis ...
                        {
                            const float4 r0 = a.sub_box[2 * (size_t)(bbt + j)], r1 = a.sub_box[2 * (size_t)(bbt + j) + 1];   // ... and the ranges, one trip
                            const int ax = min(max((int)b1.w, 0), 2);          // bin_sort_kernel's sort axis
                            const float sha = (float)(lmt[3 * j + ax] - (ax == 0 ? C0 : ax == 1 ? C1 : C2));
                            *reinterpret_cast<float4*>(&S.srange[tid][0]) = make_float4(r0.x + sha, r0.y + sha, r0.z + sha, r0.w + sha);
                            *reinterpret_cast<float4*>(&S.srange[tid][4]) = make_float4(r1.x + sha, r1.y + sha, r1.z + sha, r1.w + sha);
                        }
                        const float bx[6] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y};
                        float bb2 = 0.f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float sh = (float)(lmt[3 * j + c] - (c == 0 ? C0 : c == 1 ? C1 : C2));
                            S.sdesc[tid].sh[c] = sh;
                            lo[c] = bx[c] + sh; hi[c] = bx[3 + c] + sh;
                            mb = fmaxf(mb, fmaxf(fabsf(lo[c]), fabsf(hi[c])));
                            const float d = fmaxf(fmaxf(lo[c] - v[3 + c], v[c] - hi[c]), 0.f);
                            bb2 = fmaf(d, d, bb2);
                        }
                        const float e = 4e-7f * (qabs_max + fmaxf(mb, mabs));
                        const float rhs = rmax2 + d2_err(rmax2, e) + dot_err(qq_max, rmax2);
                        const float lhs = bb2 * 0.9999f - d2_err(bb2, e);
                        flag = (lhs > rhs) ? 0.f : 1.f + fminf(fmaxf(b1.w, 0.f), 2.f);     // (NaN compares false -> kept)
                    }
                    sbox[tid][0] = lo[0]; sbox[tid][1] = lo[1]; sbox[tid][2] = lo[2];
                    sbox[tid][3] = hi[0]; sbox[tid][4] = hi[1]; sbox[tid][5] = hi[2];
                    sbox[tid][6] = mb; sbox[tid][7] = flag;
                    const unsigned fm = __ballot_sync(LIP_FULL, flag != 0.f);
                    if (lane == 0) S.cand[warp] = fm;
                }
                __syncthreads();
                unsigned long long cm = LIP_MAP_BOXB > 32 ? ((unsigned long long)S.cand[1] << 32) | S.cand[0] : (unsigned long long)S.cand[0];     // uniform
                while (cm) {
                    const int t = __ffsll((long long)cm) - 1;
                    cm &= cm - 1;
                    const int axis = (int)sbox[t][7] - 1;
                    unsigned nd = 0u;                        // ranges of bin j0 + t that some query of this lane still has to see
#pragma unroll
                    for (int k = 0; k < Q; ++k) {
                        if (!valid[k]) continue;
                        const float bd = fmaxf(best[k] + qq[k], 0.f);
                        const float e = 4e-7f * (qabs[k] + fmaxf(sbox[t][6], mabs));
                        const float tol0 = d2_err(bd, e) + dot_err(qq[k], bd);
                        float g[3];                          // the whole bin on its own box first
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float f = -0.5f * mq[k][c];
                            g[c] = fmaxf(fmaxf(sbox[t][c] - f, f - sbox[t][3 + c]), 0.f);
                        }
                        const float bd2w = fmaf(g[2], g[2], fmaf(g[1], g[1], g[0] * g[0]));
                        if (bd2w - bd > tol0 + d2_err(bd2w, e)) continue;
                        // its ranges: the bin's box in the two other dimensions, their own extent along the sort axis
                        const float fa = -0.5f * (axis == 0 ? mq[k][0] : axis == 1 ? mq[k][1] : mq[k][2]);
                        const float other = (axis == 0 ? 0.f : g[0] * g[0]) + (axis == 1 ? 0.f : g[1] * g[1]) + (axis == 2 ? 0.f : g[2] * g[2]);
#pragma unroll
                        for (int r = 0; r < LIP_MAP_SUB; ++r) {
                            const float rl = S.srange[t][r], rh = S.srange[t][LIP_MAP_SUB + r];
                            if (rl > rh) continue;                              // empty range
                            const float da = fmaxf(fmaxf(rl - fa, fa - rh), 0.f);
                            const float bd2 = fmaf(da, da, other);
                            if (!(bd2 - bd > tol0 + d2_err(bd2, e))) nd |= 1u << r;
                        }
                    }
                    nd = __reduce_or_sync(LIP_FULL, nd);
                    const int u = (j0 + t) * LIP_MAP_SUB;
                    if (lane == 0 && nd) atomicOr(&need[u >> 5], nd << (u & 31));
                }
            }
            __syncthreads();
        }
        int wi = 0, base = 0, cur_p = 0, cur_e = 0;
        unsigned bits = 0u;
        float sh0 = 0.f, sh1 = 0.f, sh2 = 0.f;
        bool pending_home = round == 0, more = true;
        while (more) {
            while (fill < LIP_MAP_CHUNK) {
                if (cur_p >= cur_e) {                        // next bin of this round (uniform)
                    int j = -1;
                    if (round == 0) {
                        if (!pending_home) { more = false; break; }
                        pending_home = false;
                        cur_p = T.ts; cur_e = T.te; sh0 = sh1 = sh2 = 0.f;
                        mabs = fmaxf(mabs, T.home_mabs);
                        continue;
                    }
                    int u = -1;                              // next (bin, range) of this round
                    while (bits == 0u && wi < nwords) { bits = need[wi]; base = 32 * wi; ++wi; }
                    int nr = 1;                              // ... and the ranges of the same bin that follow it directly: one piece
                    if (bits) {
                        u = base + __ffs(bits) - 1; bits &= bits - 1;
                        while ((u + nr) % LIP_MAP_SUB != 0 && ((bits >> ((u + nr) & 31)) & 1u)) { bits &= ~(1u << ((u + nr) & 31)); ++nr; }
                    }
                    if (u < 0) { more = false; break; }
                    j = u / LIP_MAP_SUB;
                    int bs, be;
                    if (K <= LIP_MAP_BOXB) {             // descriptors left in SMEM by the box pass
                        bs = S.sdesc[j].start; be = S.sdesc[j].end;
                        sh0 = S.sdesc[j].sh[0]; sh1 = S.sdesc[j].sh[1]; sh2 = S.sdesc[j].sh[2];
                        mabs = fmaxf(mabs, S.sbox[j][6]);
                    } else {
                        bs = a.bin_off[bbt + j]; be = a.bin_off[bbt + j + 1];
                        if (be > bs) {
                            sh0 = (float)(lmt[3 * j] - C0); sh1 = (float)(lmt[3 * j + 1] - C1); sh2 = (float)(lmt[3 * j + 2] - C2);
                            mabs = fmaxf(mabs, a.bin_box[8 * (size_t)(bbt + j) + 6] + fmaxf(fmaxf(fabsf(sh0), fabsf(sh1)), fabsf(sh2)));
                        }
                    }
                    const int per = (be - bs + LIP_MAP_SUB - 1) / LIP_MAP_SUB;
                    cur_p = min(be, bs + (u % LIP_MAP_SUB) * per); cur_e = min(be, cur_p + nr * per);
                    continue;
                }
                const int cnt = min(cur_e - cur_p, LIP_MAP_CHUNK - fill);
                for (int i0 = tid; i0 < cnt; i0 += 4 * LIP_MAP_THREADS) {
                    float4 t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int i = i0 + u * LIP_MAP_THREADS; if (i < cnt) t[u] = a.bin_pts[cur_p + i]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * LIP_MAP_THREADS;
                        if (i < cnt) {
                            const float x = t[u].x + sh0, y = t[u].y + sh1, z = t[u].z + sh2;
                            tp[fill + i] = make_float4(x, y, z, fmaf(x, x, fmaf(y, y, z * z)));
                            tv[fill + i] = __float_as_int(t[u].w);
                        }
                    }
                }
                fill += cnt; cur_p += cnt;
            }
            if (fill > 0) {                                  // scan the staged chunk
                const int npad = (fill + 7) & ~7;
                if (tid < npad - fill) { tp[fill + tid] = make_float4(0.f, 0.f, 0.f, 3.0e38f); tv[fill + tid] = 0x7fffffff; }
                __syncthreads();
                float prev[Q];
#pragma unroll
                for (int k = 0; k < Q; ++k) prev[k] = best[k];
                if (qw > (Q + 1) / 2) tile_scan<Q, true>(tp, npad, mq, best, sec, blk);
                else if (qw > 0) tile_scan<(Q + 1) / 2, true>(tp, npad, mq, best, sec, blk);
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (best[k] < prev[k]) idx[k] = tv[tile_recover<true>(tp, blk[k], best[k], mq[k], sec_in[k])];
                __syncthreads();
                staged += fill;
                fill = 0;
            }
        }
    }
    if (a.profile && tid == 0) atomicAdd(&a.counters[7], (unsigned long long)(staged * nq));
#ifdef LIP_DEBUG_SLOW_TILES
    const long long dbg_t1 = clock64();
    int dbg_ties = 0;
#endif

    // ---- result: near ties are decided in float64 over everything that was staged, a warp per query -------
    const double* __restrict__ xt = mesh_xyz(a, mt);
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        const float d2b = fmaxf(best[k] + qq[k], 0.f), d2s = fmaxf(fminf(sec[k], sec_in[k]) + qq[k], 0.f);
        const float e = 4e-7f * (qabs[k] + mabs);
        const float tol = d2_err(d2b, e) + d2_err(d2s, e) + 2.f * dot_err(qq[k], d2s);
        int res = idx[k];
        const bool tie = valid[k] && !(d2s - d2b > tol);
        unsigned ties = __ballot_sync(LIP_FULL, tie);
        if (ties && lane == 0) atomicAdd(&a.counters[3], (unsigned long long)__popc(ties));
#ifdef LIP_DEBUG_SLOW_TILES
        dbg_ties += __popc(ties);
#endif
        // A rescan reads the staged ranges again as fp32 offsets (coalesced) and takes the float64 coordinates (a gather by vertex id)
        // only of candidates whose fp32 distance, less its error bound, is not above the best one's upper bound.
        const float upper = (d2b + d2_err(d2b, e) + dot_err(qq[k], d2b)) * 1.000001f;
        while (ties) {
            const int src = __ffs(ties) - 1;
            ties &= ties - 1;
            const size_t N = (size_t)a.n_total;
            const size_t pq = (size_t)so + first + (tid - lane + src) + k * LIP_MAP_THREADS;     // slot of lane `src`
            const double q0 = a.bin_q[pq], q1 = a.bin_q[N + pq], q2 = a.bin_q[2 * N + pq];
            const float fx = -0.5f * __shfl_sync(LIP_FULL, mq[k][0], src), fy = -0.5f * __shfl_sync(LIP_FULL, mq[k][1], src),
                        fz = -0.5f * __shfl_sync(LIP_FULL, mq[k][2], src);
            const float es = __shfl_sync(LIP_FULL, e, src), us = __shfl_sync(LIP_FULL, upper, src);
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int j = 0; j < K; ++j) {
                const unsigned m4 = j == home ? (1u << LIP_MAP_SUB) - 1 : (need[(j * LIP_MAP_SUB) >> 5] >> ((j * LIP_MAP_SUB) & 31)) & ((1u << LIP_MAP_SUB) - 1);
                if (!m4) continue;
                const int bs = a.bin_off[bbt + j], be = a.bin_off[bbt + j + 1], per = (be - bs + LIP_MAP_SUB - 1) / LIP_MAP_SUB;
                const float s0 = j == home ? 0.f : (float)(lmt[3 * j] - C0), s1 = j == home ? 0.f : (float)(lmt[3 * j + 1] - C1),
                            s2 = j == home ? 0.f : (float)(lmt[3 * j + 2] - C2);         // as staged
                const int b1 = bs + per, b2 = bs + 2 * per, b3 = bs + 3 * per;
                static_assert(LIP_MAP_SUB == 4, "range of a position from three comparisons");
#pragma unroll 1
                for (int p = bs + lane; p < be; p += 32) {
                    const int r = (p >= b1) + (p >= b2) + (p >= b3);
                    if (!((m4 >> r) & 1u)) continue;         // ranges that were not staged cannot hold anything nearer
                    const float4 t = a.bin_pts[p];
                    const float dx = (t.x + s0) - fx, dy = (t.y + s1) - fy, dz = (t.z + s2) - fz;
                    const float d2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (d2f - d2_err(d2f, es) > us) continue;
                    const int vv = __float_as_int(t.w);
                    const double d = rdist64(q0, q1, q2, xt[3 * (size_t)vv], xt[3 * (size_t)vv + 1], xt[3 * (size_t)vv + 2]);
                    if (d < bd || (d == bd && vv < bi)) { bd = d; bi = vv; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(LIP_FULL, bd, o);
                const int oi = __shfl_xor_sync(LIP_FULL, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (lane == src && bi != 0x7fffffff) res = bi;
        }
        if (valid[k]) {
            const int u = ub + vq[k];
            store_map(a, u, res);
        }
    }
#ifdef LIP_DEBUG_SLOW_TILES
    if (tid == 0 && clock64() - dbg_t1 > LIP_DEBUG_SLOW_TILES / 2)
        printf("  tile %d: tie phase %lld clk, staged %lld, ties(warp0) %d, K=%d nq=%d mabs=%g\n", tile, clock64() - dbg_t1, staged, dbg_ties, K, nq, mabs);
#endif
}

__global__ void __launch_bounds__(LIP_MAP_THREADS, LIP_MAP_MINBLOCKS) blend_tile_kernel(LipArgs a) {
    __shared__ MapSmem S;
    __shared__ int s_tile;
    const int n_tiles = a.sched[20];
    for (;;) {                                              // persistent CTAs over the tile list
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&a.sched[21], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
#ifdef LIP_DEBUG_SLOW_TILES
        const long long t0 = clock64();
        blend_tile(a, S, tile);
        const long long dt = clock64() - t0;
        if (threadIdx.x == 0 && dt > LIP_DEBUG_SLOW_TILES) {
            const MapTile T = a.map_tiles[tile];
            printf("slow tile %d of %d: %lld clk K=%d nq=%d home=%d home bin %d..%d\n", tile, n_tiles, dt, T.K, T.nq, T.b - T.bb, T.ts, T.te);
        }
#else
        blend_tile(a, S, tile);
#endif
    }
}

// =============================================================================================
// K4 on the 5th-generation tensor cores (map mode 2).  Same tiles, same two rounds (home bin, then every bin the box
// test cannot rule out), same float64 decision of near ties as blend_tile_kernel — but the scores
//     c(q, y) = |y|^2 - 2 q.y
// of 128 queries x 128 candidates come out of two tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8) into TMEM:
//     q = qh + ql, y = yh + yl, |y|^2 = nh + nl   (every part exactly representable in tf32)
//     A row (query)        [-2qh.x -2qh.y -2qh.z 1 | -2ql.x -2ql.y -2ql.z 0]
//     B row (candidate)    hi plane [yh.x yh.y yh.z nh], lo plane [yl.x yl.y yl.z nl], each read for BOTH K halves (lbo = 0)
//     D = A Bhi^T ; D += A Blo^T  =  nh + nl - 2 (qh + ql).(yh + yl)                  (tools/micro/tcgen05_scan_bench.cu)
// Warp roles (192 threads): warps 0-3 own the 128 queries (TMEM lane = query): tcgen05.ld of 32 scores at a time, block
// minima of 8 + top-2 tracking exactly as tile_scan, winner recovery from the staged candidates, box cull, final decision;
// warps 4 and 5 are producers: each owns one SMEM stage + one TMEM buffer, stages 128 candidates (shift into the tile frame,
// |y|^2, hi / lo split), issues the two MMAs and commits to the stage's mbarrier.  Stages are handed back by the query warps
// through a second mbarrier once the scores are read and the winner recovered; no CTA-wide barrier inside a tile.
// The tensor-core score differs from the exact one by at most tc_err (measured 2^-20.4 relative to |y|^2 + 2|q||y|; bound used
// 2^-18): the tolerance of the final decision is widened by it, so ambiguous queries go to the float64 scan as before and the
// maps stay exact.
// =============================================================================================
#define LIP_TC_THREADS 192
#define LIP_TC_N 128
struct __align__(128) MapTcSmem {
    float4 A[2][128];                                   // query operand: K halves as planes
    float4 hi[2][LIP_TC_N], lo[2][LIP_TC_N];            // candidates of the two stages
    int tv[2][LIP_TC_N];                                // ... and their vertex ids
    unsigned need[LIP_MAX_LANDMARKS / 32];
    float sbox[LIP_MAP_BOXBATCH][8];
    float red[4][8];
    unsigned cand[2];
    struct { int start, end; float sh[3]; } sdesc[LIP_MAP_BOXBATCH];
    unsigned long long t_full[2], t_empty[2], a_ready, cull_done, r1_ready;
    int tot1;
    float mabs1;
    unsigned tmem_base;
    int tile;
};
__device__ __forceinline__ float tc_err(float qq, float d2) {
    const float nq = sqrtf(qq), ny = nq + 1.001f * sqrtf(d2) + 1e-3f;
    return 4e-6f * fmaf(ny, ny, 2.f * nq * ny);
}

// ---- producer side: stage `take` candidates of bin_pts[src ..] (shifted by sh) at slot `at` of stage w ----
__device__ __forceinline__ void tc_stage(const LipArgs& a, MapTcSmem& S, int w, int at, int src, int take, float sh0, float sh1, float sh2, int lane) {
    for (int i = lane; i < take; i += 32) {
        const float4 t = a.bin_pts[src + i];
        const float x = t.x + sh0, y = t.y + sh1, z = t.z + sh2;
        const float n = fmaf(x, x, fmaf(y, y, z * z));
        const float hx = tf32_hi(x), hy = tf32_hi(y), hz = tf32_hi(z), hn = tf32_hi(n);
        S.hi[w][at + i] = make_float4(hx, hy, hz, hn);
        S.lo[w][at + i] = make_float4(tf32_hi(x - hx), tf32_hi(y - hy), tf32_hi(z - hz), tf32_hi(n - hn));
        S.tv[w][at + i] = __float_as_int(t.w);
    }
}
// chunk `c` of this CTA's chunk sequence is complete in stage w = c & 1: pad, make it visible to the tensor core, multiply
__device__ __forceinline__ void tc_fire(MapTcSmem& S, int w, int fill, unsigned tmem, int lane) {
    const int pad_to = min(LIP_TC_N, (fill + 31) & ~31);
    for (int i = fill + lane; i < pad_to; i += 32) {
        S.hi[w][i] = make_float4(0.f, 0.f, 0.f, 1e30f); S.lo[w][i] = make_float4(0.f, 0.f, 0.f, 0.f); S.tv[w][i] = 0x7fffffff;
    }
    fence_async_smem();
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        tc_fence_after();
        constexpr unsigned idesc = umma_idesc_tf32(128, LIP_TC_N);
        const unsigned long long ad = umma_desc(&S.A[0][0], 128 * 16, 128);
        umma_tf32(tmem + w * LIP_TC_N, ad, umma_desc(&S.hi[w][0], 0, 128), idesc, 0);
        umma_tf32(tmem + w * LIP_TC_N, ad, umma_desc(&S.lo[w][0], 0, 128), idesc, 1);
        umma_commit(&S.t_full[w]);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(LIP_TC_THREADS, 2) blend_tile_tc_kernel(LipArgs a) {
    __shared__ MapTcSmem S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = a.sched[20];
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&S.tmem_base)), "r"(2 * LIP_TC_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        mbar_init(&S.t_full[0], 1); mbar_init(&S.t_full[1], 1); mbar_init(&S.t_empty[0], 4); mbar_init(&S.t_empty[1], 4);
        mbar_init(&S.a_ready, 1); mbar_init(&S.cull_done, 1); mbar_init(&S.r1_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = S.tmem_base;
    unsigned chunk = 0, n_tile = 0, n_cull = 0;              // identical in every role: they give the barrier phases
    for (;;) {                                              // persistent CTAs over the tile list
        __syncthreads();
        if (tid == 0) S.tile = atomicAdd(&a.sched[21], 1);
        __syncthreads();
        const int tile = S.tile;
        if (tile >= n_tiles) break;
        const MapTile T = a.map_tiles[tile];
        const int first = T.first, m = T.m, ub = T.ub, bb = T.bb, K = T.K, bbt = T.bbt, so = T.so, nq = T.nq;
        const int mt = pair_mesh(a, m);
        const int home = T.b - bb;
        const double* __restrict__ lmt = mesh_lm(a, mt);
        const double C0 = T.C[0], C1 = T.C[1], C2 = T.C[2];
        const int tot0 = T.te - T.ts, n0 = (tot0 + LIP_TC_N - 1) / LIP_TC_N;
        const int nwords = (K + 31) / 32;
        int tot1 = 0;
        if (warp >= 4) {
            // ================================ producers ================================
            const int w = warp - 4;
            mbar_wait(&S.a_ready, n_tile & 1);               // the queries of this tile are in S.A
            int pos = 0;                                     // position in the tile's candidate stream
            auto feed = [&](int src, int cnt, float sh0, float sh1, float sh2, int total) {
                while (cnt > 0) {
                    const int at = pos & (LIP_TC_N - 1), take = min(cnt, LIP_TC_N - at);
                    const unsigned c = chunk + (unsigned)(pos / LIP_TC_N);
                    const bool mine = (int)(c & 1u) == w;
                    if (mine) {
                        if (at == 0 && (c >> 1) >= 1u) mbar_wait(&S.t_empty[w], ((c >> 1) - 1u) & 1u);     // stage and TMEM buffer drained
                        tc_stage(a, S, w, at, src, take, sh0, sh1, sh2, lane);
                    }
                    pos += take; src += take; cnt -= take;
                    if (mine && ((pos & (LIP_TC_N - 1)) == 0 || pos == total)) tc_fire(S, w, ((pos - 1) & (LIP_TC_N - 1)) + 1, tmem, lane);
                }
            };
            if (tot0 > 0) feed(T.ts, tot0, 0.f, 0.f, 0.f, tot0);
            chunk += (unsigned)n0;
            if (K > 1) {
                mbar_wait(&S.cull_done, n_cull & 1);         // S.need (and S.sdesc for K <= LIP_MAP_BOXBATCH) are final
                int cnt1 = 0;
                float mabs1 = 0.f;
                for (int j = lane; j < K; j += 32) {
                    if (!((S.need[j >> 5] >> (j & 31)) & 1u)) continue;
                    const int js = a.bin_off[bbt + j], je = a.bin_off[bbt + j + 1];
                    if (je <= js) continue;
                    cnt1 += je - js;
                    const float s0 = (float)(lmt[3 * j] - C0), s1 = (float)(lmt[3 * j + 1] - C1), s2 = (float)(lmt[3 * j + 2] - C2);
                    mabs1 = fmaxf(mabs1, a.bin_box[8 * (size_t)(bbt + j) + 6] + fmaxf(fmaxf(fabsf(s0), fabsf(s1)), fabsf(s2)));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { cnt1 += __shfl_xor_sync(LIP_FULL, cnt1, o); mabs1 = fmaxf(mabs1, __shfl_xor_sync(LIP_FULL, mabs1, o)); }
                tot1 = cnt1;
                if (w == 0 && lane == 0) { S.tot1 = tot1; S.mabs1 = mabs1; __threadfence_block(); mbar_arrive(&S.r1_ready); }
                pos = 0;
                for (int wi = 0; wi < nwords; ++wi) {
                    unsigned bits = S.need[wi];
                    while (bits) {
                        const int j = 32 * wi + __ffs(bits) - 1;
                        bits &= bits - 1;
                        const int js = a.bin_off[bbt + j], je = a.bin_off[bbt + j + 1];
                        if (je <= js) continue;
                        feed(js, je - js, (float)(lmt[3 * j] - C0), (float)(lmt[3 * j + 1] - C1), (float)(lmt[3 * j + 2] - C2), tot1);
                    }
                }
                chunk += (unsigned)((tot1 + LIP_TC_N - 1) / LIP_TC_N);
                ++n_cull;
            }
        } else {
            // ================================ query warps ================================
            const bool valid = tid < nq;
            const int slot = first + tid;
            const int vq = valid ? a.part_vid[so + slot] : 0;
            float fx = 0.f, fy = 0.f, fz = 0.f;
            if (valid) {                                     // blended positions in bin order, written by blend_kernel
                const size_t N = (size_t)a.n_total, p = (size_t)so + slot;
                fx = (float)(a.bin_q[p] - C0); fy = (float)(a.bin_q[N + p] - C1); fz = (float)(a.bin_q[2 * N + p] - C2);
            }
            const float mq[3] = {-2.f * fx, -2.f * fy, -2.f * fz};
            const float qq = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
            const float qabs = fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz));
            {
                const float hx = tf32_hi(fx), hy = tf32_hi(fy), hz = tf32_hi(fz);
                S.A[0][tid] = make_float4(-2.f * hx, -2.f * hy, -2.f * hz, 1.f);
                S.A[1][tid] = make_float4(-2.f * tf32_hi(fx - hx), -2.f * tf32_hi(fy - hy), -2.f * tf32_hi(fz - hz), 0.f);
            }
            for (int i = tid; i < nwords; i += 128) S.need[i] = 0u;
            fence_async_smem();
            __threadfence_block();
            bar_sync_named(1, 128);
            if (tid == 0) mbar_arrive(&S.a_ready);
            float best = INFINITY, sec = INFINITY, bestf = INFINITY, sec_in = INFINITY;      // best / sec: tensor-core scores; bestf / sec_in: fp32 FFMA
            int idx = -1;                                                                   // scores of the winner's block of 8
            float mabs = tot0 > 0 ? T.home_mabs : 0.f;
            const unsigned lane_addr = (unsigned)(warp * 32) << 16;
            auto consume = [&](int total) {
                const int nch = (total + LIP_TC_N - 1) / LIP_TC_N;
                for (int i = 0; i < nch; ++i) {
                    const unsigned c = chunk + (unsigned)i;
                    const int s = (int)(c & 1u);
                    const int cnt = min(LIP_TC_N, total - LIP_TC_N * i), ng = (cnt + 31) >> 5;
                    mbar_wait(&S.t_full[s], (c >> 1) & 1u);
                    tc_fence_after();
                    const unsigned tbase = tmem + lane_addr + (unsigned)(s * LIP_TC_N);
                    const float prev = best;
                    int blk = 0;
                    float r[2][32];
                    tmem_ld32(tbase, r[0]);
                    tmem_wait_ld();
#pragma unroll
                    for (int g = 0; g < LIP_TC_N / 32; ++g) {
                        if (g < ng) {
                            if (g + 1 < ng) tmem_ld32(tbase + (g + 1) * 32, r[(g + 1) & 1]);
                            const float* v = r[g & 1];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                float bm = fminf(fminf(v[8 * k], v[8 * k + 1]), v[8 * k + 2]);
                                bm = fminf(fminf(bm, v[8 * k + 3]), v[8 * k + 4]);
                                bm = fminf(fminf(bm, v[8 * k + 5]), v[8 * k + 6]);
                                bm = fminf(bm, v[8 * k + 7]);
                                sec = fminf(sec, fmaxf(bm, best));
                                blk = (bm < best) ? g * 32 + 8 * k : blk;
                                best = fminf(best, bm);
                            }
                            if (g + 1 < ng) tmem_wait_ld();
                        }
                    }
                    if (best < prev) {                       // the winner so far sits in this chunk: its position, from the staged candidates
                        float b1 = INFINITY, b2 = INFINITY;
                        int pos = blk;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 h = S.hi[s][blk + k], l = S.lo[s][blk + k];
                            const float cs = fmaf(mq[0], h.x + l.x, fmaf(mq[1], h.y + l.y, fmaf(mq[2], h.z + l.z, h.w + l.w)));
                            if (cs < b1) { b2 = b1; b1 = cs; pos = blk + k; } else b2 = fminf(b2, cs);
                        }
                        bestf = b1; sec_in = b2; idx = S.tv[s][pos];
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.t_empty[s]);
                }
                chunk += (unsigned)nch;
            };
            consume(tot0);
            if (K > 1) {
                // ---- box cull of the other target bins: CTA-wide bounds, then per query (as blend_tile) ----
                float (*sbox)[8] = S.sbox;
                float v[8];
                v[6] = 0.f; v[7] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) { v[c] = INFINITY; v[3 + c] = -INFINITY; }
                if (valid) {
                    v[0] = v[3] = fx; v[1] = v[4] = fy; v[2] = v[5] = fz;
                    v[6] = fmaxf(bestf + qq, 0.f);           // the distance of a real candidate (or infinity): an upper bound of the answer
                    v[7] = qabs;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = -warp_maxf(-v[c]);
#pragma unroll
                for (int c = 3; c < 8; ++c) v[c] = warp_maxf(v[c]);
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) S.red[warp][c] = v[c];
                }
                bar_sync_named(1, 128);
#pragma unroll
                for (int w = 0; w < 4; ++w) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = fminf(v[c], S.red[w][c]);
#pragma unroll
                    for (int c = 3; c < 8; ++c) v[c] = fmaxf(v[c], S.red[w][c]);
                }
                const float rmax2 = v[6], qabs_max = v[7];
                const float qq_max = 3.f * qabs_max * qabs_max;
                for (int j0 = 0; j0 < K; j0 += LIP_MAP_BOXBATCH) {
                    bar_sync_named(1, 128);
                    if (tid < LIP_MAP_BOXBATCH) {            // boxes of this batch of bins in the tile frame + CTA-level cull
                        const int j = j0 + tid;
                        float flag = 0.f, lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f}, mb = 0.f;
                        const int js = j < K ? a.bin_off[bbt + j] : 0, je = j < K ? a.bin_off[bbt + j + 1] : 0;
                        if (j < K && j != home && je > js) {
                            const float* bx = a.bin_box + 8 * (size_t)(bbt + j);
                            float bb2 = 0.f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float sh = (float)(lmt[3 * j + c] - (c == 0 ? C0 : c == 1 ? C1 : C2));
                                lo[c] = bx[c] + sh; hi[c] = bx[3 + c] + sh;
                                mb = fmaxf(mb, fmaxf(fabsf(lo[c]), fabsf(hi[c])));
                                const float d = fmaxf(fmaxf(lo[c] - v[3 + c], v[c] - hi[c]), 0.f);
                                bb2 = fmaf(d, d, bb2);
                            }
                            const float e = 8e-7f * (qabs_max + fmaxf(mb, mabs));
                            const float rhs = rmax2 + d2_err(rmax2, e) + 2.f * dot_err(qq_max, rmax2);
                            const float lhs = bb2 * 0.9999f - d2_err(bb2, e);
                            flag = (lhs > rhs) ? 0.f : 1.f;      // (NaN compares false -> kept)
                        }
                        sbox[tid][0] = lo[0]; sbox[tid][1] = lo[1]; sbox[tid][2] = lo[2];
                        sbox[tid][3] = hi[0]; sbox[tid][4] = hi[1]; sbox[tid][5] = hi[2];
                        sbox[tid][6] = mb; sbox[tid][7] = flag;
                        const unsigned fm = __ballot_sync(LIP_FULL, flag != 0.f);
                        if (lane == 0) S.cand[warp] = fm;
                    }
                    bar_sync_named(1, 128);
                    unsigned w0 = 0u, w1 = 0u;
                    unsigned long long cm = ((unsigned long long)S.cand[1] << 32) | S.cand[0];     // uniform
                    while (cm) {
                        const int t = __ffsll((long long)cm) - 1;
                        cm &= cm - 1;
                        bool nd = false;
                        if (valid) {
                            float bd2 = 0.f;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float f = c == 0 ? fx : c == 1 ? fy : fz;
                                const float d = fmaxf(fmaxf(sbox[t][c] - f, f - sbox[t][3 + c]), 0.f);
                                bd2 = fmaf(d, d, bd2);
                            }
                            const float bd = fmaxf(bestf + qq, 0.f);
                            const float e = 8e-7f * (qabs + fmaxf(sbox[t][6], mabs));
                            const float tol = d2_err(bd, e) + d2_err(bd2, e) + 2.f * dot_err(qq, bd);
                            nd = !(bd2 - bd > tol);
                        }
                        if (nd) { if (t < 32) w0 |= 1u << t; else w1 |= 1u << (t - 32); }
                    }
                    w0 = __reduce_or_sync(LIP_FULL, w0); w1 = __reduce_or_sync(LIP_FULL, w1);
                    if (lane == 0) {
                        if (w0) atomicOr(&S.need[j0 / 32], w0);
                        if (w1) atomicOr(&S.need[j0 / 32 + 1], w1);
                    }
                }
                __threadfence_block();
                bar_sync_named(1, 128);
                if (tid == 0) mbar_arrive(&S.cull_done);
                mbar_wait(&S.r1_ready, n_cull & 1);
                tot1 = S.tot1;
                mabs = fmaxf(mabs, S.mabs1);
                consume(tot1);
                ++n_cull;
            }
            if (a.profile && tid == 0) atomicAdd(&a.counters[7], (unsigned long long)(tot0 + tot1) * (unsigned long long)nq);

            // ---- result: near ties are decided in float64 over everything that was staged, a warp per query -------
            const double* __restrict__ xt = mesh_xyz(a, mt);
            const float d2b = fmaxf(bestf + qq, 0.f), d2s = fmaxf(fminf(sec, sec_in) + qq, 0.f);
            const float e = 8e-7f * (qabs + mabs);
            const float tol = d2_err(d2b, e) + d2_err(d2s, e) + 4.f * dot_err(qq, d2s) + 2.f * tc_err(qq, d2s);
            int res = idx;
            const bool tie = valid && !(d2s - d2b > tol);
            unsigned ties = __ballot_sync(LIP_FULL, tie);
            if (ties && lane == 0) atomicAdd(&a.counters[3], (unsigned long long)__popc(ties));
            while (ties) {
                const int src = __ffs(ties) - 1;
                ties &= ties - 1;
                const size_t N = (size_t)a.n_total;
                const size_t pq = (size_t)so + first + (tid - lane + src);       // slot of lane `src`
                const double q0 = a.bin_q[pq], q1 = a.bin_q[N + pq], q2 = a.bin_q[2 * N + pq];
                double bd = INFINITY;
                int bi = 0x7fffffff;
                for (int j = 0; j < K; ++j) {
                    if (j != home && !((S.need[j >> 5] >> (j & 31)) & 1u)) continue;
                    for (int p = a.bin_off[bbt + j] + lane; p < a.bin_off[bbt + j + 1]; p += 32) {
                        const int vv = a.part_vid[p];
                        const double d = rdist64(q0, q1, q2, xt[3 * (size_t)vv], xt[3 * (size_t)vv + 1], xt[3 * (size_t)vv + 2]);
                        if (d < bd || (d == bd && vv < bi)) { bd = d; bi = vv; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double od = __shfl_xor_sync(LIP_FULL, bd, o);
                    const int oi = __shfl_xor_sync(LIP_FULL, bi, o);
                    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                }
                if (lane == src) res = bi;
            }
            if (valid) store_map(a, ub + vq, res);
        }
        ++n_tile;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * LIP_TC_N));
}

// queries ring 1 could not certify: one thread per query, ring 1 + ring 2, else far list
__global__ void __launch_bounds__(128) retry_nn_kernel(LipArgs a) {
    const int n_retry = a.sched[7];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_retry; i += gridDim.x * blockDim.x) {
        const int u = a.retry_list[2 * (size_t)i], mt = a.retry_list[2 * (size_t)i + 1];
        const double q0 = a.retry_q[3 * (size_t)i], q1 = a.retry_q[3 * (size_t)i + 1], q2 = a.retry_q[3 * (size_t)i + 2];
        const GridHdr h = a.grid_hdr[mt];
        NNResult r = grid_nearest(h, a.grid_start + a.mesh_tab[mt], a.grid_pts + a.mesh_ubase[mt], mesh_xyz(a, mt),
                                  q0, q1, q2, a.counters, nullptr, 0);
        if (r.idx < 0) { far_push(a, u, mt, q0, q1, q2); continue; }
        store_map(a, u, r.idx);
    }
}

// raw query: nearest target (mesh1 side) of every mesh2-side vertex, with distance (get_NN_in_target)
__global__ void __launch_bounds__(LIP_TILE) nn_query_kernel(LipArgs a) {
    const int m = a.tile_mesh[blockIdx.x];
    const int v = a.tile_first[blockIdx.x] + threadIdx.x;
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    if (v >= n) return;
    const double* xyz = mesh_xyz(a, m);
    const int mt = pair_mesh(a, m);
    const GridHdr h = a.grid_hdr[mt];
    double dist;
    NNResult r = grid_nearest(h, a.grid_start + a.mesh_tab[mt], a.grid_pts + a.mesh_ubase[mt], mesh_xyz(a, mt),
                              xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2], a.counters, &dist, a.profile);
    const int u = ub + v;
    if (r.idx < 0) { far_push(a, u, mt, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]); return; }
    store_map(a, u, r.idx); if (m >= a.S && a.q_dist) a.q_dist[u - a.n1_total] = dist;
}

// ---------------------------------------------------------------------------------------------
// queries the ring search could not certify (farther than two cells from every target point, e.g.
// the grown tip of a leaf): one WARP per query scans the whole target mesh, lanes striding over the
// cell-sorted float4 points (coalesced), warp-shuffle argmin, same float64 tie rule.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) far_nn_kernel(LipArgs a) {
    const int lane = threadIdx.x & 31;
    const int n_far = a.sched[3];
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&a.sched[6], 1);
        w = __shfl_sync(LIP_FULL, w, 0);
        if (w >= n_far) break;
        const int u = a.far_list[2 * (size_t)w], mt = a.far_list[2 * (size_t)w + 1];
        const double q0 = a.far_q[3 * (size_t)w], q1 = a.far_q[3 * (size_t)w + 1], q2 = a.far_q[3 * (size_t)w + 2];
        const GridHdr& h = a.grid_hdr[mt];
        const float4* __restrict__ pts = a.grid_pts + a.mesh_ubase[mt];
        const double* __restrict__ xyz = mesh_xyz(a, mt);
        const int n = h.n;
        const float qx = (float)(q0 - h.c[0]), qy = (float)(q1 - h.c[1]), qz = (float)(q2 - h.c[2]);
        float best = INFINITY, sec = INFINITY;
        int idx = 0;
        for (int j = lane; j < n; j += 32) {
            const float4 p = __ldg(&pts[j]);
            float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            float d = dx * dx; d = fmaf(dy, dy, d); d = fmaf(dz, dz, d);
            nn_update(d, __float_as_int(p.w), best, sec, idx);
        }
        float ball = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ball = fminf(ball, __shfl_xor_sync(LIP_FULL, ball, o));
        const int win = __ffs(__ballot_sync(LIP_FULL, best == ball)) - 1;
        float sall = (lane == win) ? sec : best;             // runner-up over the whole warp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sall = fminf(sall, __shfl_xor_sync(LIP_FULL, sall, o));
        idx = __shfl_sync(LIP_FULL, idx, win);
        const float e = 1.2e-7f * (fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) + h.mabs);
        if (sall - ball <= d2_err(ball, e) + d2_err(sall, e)) {
            double bd = INFINITY;
            int bi = 0x7fffffff;
            for (int v = lane; v < n; v += 32) {
                double d = rdist64(q0, q1, q2, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
                if (d < bd) { bd = d; bi = v; }              // ascending v per lane: lowest index kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double od = __shfl_xor_sync(LIP_FULL, bd, o);
                int oi = __shfl_xor_sync(LIP_FULL, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            idx = bi;
            if (lane == 0) atomicAdd(&a.counters[3], 1ull);
        }
        if (lane == 0) {
            store_map(a, u, idx);
            if (u >= a.n1_total) {
                if (a.q_dist)
                    a.q_dist[u - a.n1_total] = sqrt(rdist64(q0, q1, q2, xyz[3 * (size_t)idx], xyz[3 * (size_t)idx + 1], xyz[3 * (size_t)idx + 2]));
            }
        }
    }
}

// =============================================================================================
// get_best_fit_transform on its own (registration_icp.py:8-46): one CTA
// =============================================================================================
__global__ void __launch_bounds__(256) kabsch_kernel(const double* __restrict__ A, const double* __restrict__ Bp, int n,
                                                     double* R_out, double* t_out, int* status) {
    __shared__ double red[8];
    const int tid = threadIdx.x;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s[k] += A[3 * (size_t)i + k]; s[3 + k] += Bp[3 * (size_t)i + k]; }
    }
    double ca[3], cb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ca[k] = block_sum(s[k], red, tid) / n; cb[k] = block_sum(s[3 + k], red, tid) / n; }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += 256) {
        double a0 = A[3 * (size_t)i] - ca[0], a1 = A[3 * (size_t)i + 1] - ca[1], a2 = A[3 * (size_t)i + 2] - ca[2];
        double b0 = Bp[3 * (size_t)i] - cb[0], b1 = Bp[3 * (size_t)i + 1] - cb[1], b2 = Bp[3 * (size_t)i + 2] - cb[2];
        H[0] += a0 * b0; H[1] += a0 * b1; H[2] += a0 * b2;
        H[3] += a1 * b0; H[4] += a1 * b1; H[5] += a1 * b2;
        H[6] += a2 * b0; H[7] += a2 * b1; H[8] += a2 * b2;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = block_sum(H[k], red, tid);
    if (tid == 0) {
        double R[9];
        int deg = kabsch_rotation(H, R);
        double x, y, z;
        mat3_vec(R, ca[0], ca[1], ca[2], x, y, z);
        for (int k = 0; k < 9; ++k) R_out[k] = R[k];
        t_out[0] = cb[0] - x; t_out[1] = cb[1] - y; t_out[2] = cb[2] - z;
        *status = deg ? 2 : 1;
    }
}

// get_triangle_angles (p2p_matching_seg_icp.py:27-34), n triangles
__global__ void triangle_angles_kernel(const double* __restrict__ abc, int n, double eps, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sa = abc[3 * (size_t)i], sb = abc[3 * (size_t)i + 1], sc = abc[3 * (size_t)i + 2];
    out[3 * (size_t)i] = acos((sb * sb + sc * sc - sa * sa) / (2 * sc * sb + eps));
    out[3 * (size_t)i + 1] = acos((sc * sc + sa * sa - sb * sb) / (2 * sa * sc + eps));
    out[3 * (size_t)i + 2] = acos((sa * sa + sb * sb - sc * sc) / (2 * sa * sb + eps));
}

// ---------------------------------------------------------------------------------------------
// Exact k nearest vertices of the same mesh in (distance, index) order, float64 (sklearn KDTree(x).query(x, k)).
// One CTA stages the mesh in shared memory as fp32 points relative to its first vertex, counting-sorted into buckets
// along the axis of its largest extent (a one-dimensional cell list, about 8 points per bucket); one THREAD per query
// then scans the query's own bucket and grows the range one bucket to the left / right for as long as a point beyond the
// scanned range could still enter the list: every unscanned point on the left lies strictly left of every scanned one
// (bucket index is a monotone function of the fp32 coordinate), so its distance exceeds the gap between the query and the
// leftmost scanned coordinate, minus the fp32 coordinate error.  A candidate whose fp32 squared distance exceeds the
// current k-th best float64 distance by more than a rigorous bound on the fp32 error (d2_err) is skipped; the others are
// evaluated in float64 from global memory with sklearn's operation order and ranked by (distance, index), so the result
// does not depend on the visiting order.  Meshes larger than the tile take the plain float64 loop in index order.
// side 0: mesh1 (xyz1), side 1: mesh2 (xyz2).  Query i of mesh m is its local vertex i * step.
// out[KNN * (qbase[m] + i) + k] = local index of the k-th neighbour (k = 0 is the vertex itself unless duplicated),
// -1 beyond the mesh size.  grid = (meshes, query blocks), dynamic SMEM = 16 * smem_pts + 4 * (LIP_KNN_BUCKETS + 1) bytes.
// ---------------------------------------------------------------------------------------------
#define LIP_KNN_BUCKETS 1024
// (d, v) ranks before (d2, v2)
__device__ __forceinline__ bool knn_before(double d, int v, double d2, int v2) { return d < d2 || (d == d2 && v < v2); }
template <int KNN>
__device__ __forceinline__ void knn_insert(double dn, int v, double (&d)[KNN], int (&id)[KNN]) {
    d[KNN - 1] = dn; id[KNN - 1] = v;
#pragma unroll
    for (int k = KNN - 1; k > 0; --k)
        if (knn_before(d[k], id[k], d[k - 1], id[k - 1])) {
            const double td = d[k]; d[k] = d[k - 1]; d[k - 1] = td; const int ti = id[k]; id[k] = id[k - 1]; id[k - 1] = ti;
        }
}
template <int KNN>
__global__ void __launch_bounds__(256) self_knn_kernel(LipArgs a, int side, int step, const int* __restrict__ qbase,
                                                       int* __restrict__ out, int smem_pts) {
    extern __shared__ float4 smem_knn[];
    __shared__ float s_red[8][7];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mi = blockIdx.x, m = side ? a.S + mi : mi;
    const int ub = a.mesh_ubase[m], n = a.mesh_ubase[m + 1] - ub;
    const double* __restrict__ xyz = mesh_xyz(a, m);
    const int nq = (n + step - 1) / step;
    int* o_base = out + (size_t)KNN * qbase[mi];
    if (n > smem_pts) {                                      // too large for the tile: float64 brute force, index order
        for (int qi = blockIdx.y * 256 + tid; qi < nq; qi += 256 * gridDim.y) {
            const int q = qi * step;
            const double q0 = xyz[3 * (size_t)q], q1 = xyz[3 * (size_t)q + 1], q2 = xyz[3 * (size_t)q + 2];
            double d[KNN];
            int id[KNN];
#pragma unroll
            for (int k = 0; k < KNN; ++k) { d[k] = INFINITY; id[k] = -1; }
            for (int v = 0; v < n; ++v) {
                const double dn = rdist64(q0, q1, q2, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
                if (dn < d[KNN - 1]) knn_insert<KNN>(dn, v, d, id);
            }
#pragma unroll
            for (int k = 0; k < KNN; ++k) o_base[(size_t)KNN * qi + k] = id[k];
        }
        return;
    }
    int* bstart = reinterpret_cast<int*>(smem_knn + smem_pts);            // LIP_KNN_BUCKETS + 1 bucket starts / cursors
    const double c0 = xyz[0], c1 = xyz[1], c2 = xyz[2];
    // ---- pass 1: box of the mesh -> axis, bucket scale, coordinate error -------------------------------------------
    float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int v = tid; v < n; v += 256) {
        const float f[3] = {(float)(xyz[3 * (size_t)v] - c0), (float)(xyz[3 * (size_t)v + 1] - c1), (float)(xyz[3 * (size_t)v + 2] - c2)};
#pragma unroll
        for (int c = 0; c < 3; ++c) { bl[c] = fminf(bl[c], f[c]); bh[c] = fmaxf(bh[c], f[c]); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { bl[c] = -warp_maxf(-bl[c]); bh[c] = warp_maxf(bh[c]); }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_red[warp][c] = bl[c]; s_red[warp][3 + c] = bh[c]; }
    }
    const int NB = max(1, min(LIP_KNN_BUCKETS, n / 8));
    for (int b = tid; b <= NB; b += 256) bstart[b] = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { bl[c] = fminf(bl[c], s_red[w][c]); bh[c] = fmaxf(bh[c], s_red[w][3 + c]); }
    }
    const float ex[3] = {bh[0] - bl[0], bh[1] - bl[1], bh[2] - bl[2]};
    const int axis = (ex[0] >= ex[1] && ex[0] >= ex[2]) ? 0 : (ex[1] >= ex[2]) ? 1 : 2;
    const float alo = axis == 0 ? bl[0] : axis == 1 ? bl[1] : bl[2];
    const float scale = (float)NB * 0.99999f / fmaxf(axis == 0 ? ex[0] : axis == 1 ? ex[1] : ex[2], 1e-30f);
    float rad = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) rad = fmaxf(rad, fmaxf(fabsf(bl[c]), fabsf(bh[c])));
    // per-coordinate error of an fp32 difference: both operands rounded from float64 (2^-24 |x| each) and the
    // subtraction (2^-24 |dx|), |x|, |dx| / 2 <= rad: 4 * 2^-24 * rad, times 1.5
    const float e = 3.6e-7f * rad;
    auto bucket_of = [&](float xa) { return min(NB - 1, max(0, (int)((xa - alo) * scale))); };   // monotone in xa
    // ---- pass 2: counts, starts, scatter (the order inside a bucket is arbitrary: the ranking below does not depend on it)
    for (int v = tid; v < n; v += 256) {
        const float xa = (float)(xyz[3 * (size_t)v + axis] - (axis == 0 ? c0 : axis == 1 ? c1 : c2));
        atomicAdd(&bstart[bucket_of(xa)], 1);
    }
    __syncthreads();
    {                                                        // exclusive scan of NB <= 1024 counts: 4 per thread
        int c4[4], tot = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int b = 4 * tid + k; c4[k] = b < NB ? bstart[b] : 0; tot += c4[k]; }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(LIP_FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_red[warp][6] = __int_as_float(inc);
        __syncthreads();
        int run = inc - tot;
        for (int w = 0; w < warp; ++w) run += __float_as_int(s_red[w][6]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int b = 4 * tid + k; if (b < NB) bstart[b] = run; run += c4[k]; }
        if (tid == 255) bstart[NB] = run;
    }
    __syncthreads();
    for (int v = tid; v < n; v += 256) {
        const float4 f = make_float4((float)(xyz[3 * (size_t)v] - c0), (float)(xyz[3 * (size_t)v + 1] - c1), (float)(xyz[3 * (size_t)v + 2] - c2),
                                     __int_as_float(v));
        const int pos = atomicAdd(&bstart[bucket_of(axis == 0 ? f.x : axis == 1 ? f.y : f.z)], 1);
        smem_knn[pos] = f;
    }
    __syncthreads();
    // the cursors now hold the bucket ENDS: start of bucket b = end of bucket b - 1.
    // This CTA answers the queries of a contiguous range of BUCKETS (every CTA of a mesh sorts it for itself and the order
    // inside a bucket differs from CTA to CTA, bucket membership does not); queries are taken in sorted order, so the lanes
    // of a warp are neighbours in space.
    const int b_lo = (int)((long long)NB * blockIdx.y / gridDim.y), b_hi = (int)((long long)NB * (blockIdx.y + 1) / gridDim.y);
    const int p_lo = b_lo ? bstart[b_lo - 1] : 0, p_hi = b_hi ? bstart[b_hi - 1] : 0;
    // ---- queries ----------------------------------------------------------------------------------------------------
    for (int p = p_lo + tid; p < p_hi; p += 256) {
        const float4 fq = smem_knn[p];
        const int q = __float_as_int(fq.w);
        if (q % step) continue;
        const double q0 = xyz[3 * (size_t)q], q1 = xyz[3 * (size_t)q + 1], q2 = xyz[3 * (size_t)q + 2];
        const float xq = axis == 0 ? fq.x : axis == 1 ? fq.y : fq.z;
        double d[KNN];
        int id[KNN];
#pragma unroll
        for (int k = 0; k < KNN; ++k) { d[k] = INFINITY; id[k] = -1; }
        float thr = INFINITY;                                 // fp32 distances above it cannot enter the list
        float xmin = xq, xmax = xq;                           // extent of everything scanned so far along the axis
        auto scan = [&](int b) {
            const int s = b ? bstart[b - 1] : 0, t = bstart[b];
            for (int i = s; i < t; ++i) {
                const float4 c = smem_knn[i];
                const float xa = axis == 0 ? c.x : axis == 1 ? c.y : c.z;
                xmin = fminf(xmin, xa); xmax = fmaxf(xmax, xa);
                const float dx = fq.x - c.x, dy = fq.y - c.y, dz = fq.z - c.z;
                if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > thr) continue;
                const int v = __float_as_int(c.w);
                const double dn = rdist64(q0, q1, q2, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
                if (knn_before(dn, v, d[KNN - 1], id[KNN - 1]) || id[KNN - 1] < 0) {
                    knn_insert<KNN>(dn, v, d, id);
                    const float dk = __double2float_ru(d[KNN - 1]);
                    thr = fmaf(dk + d2_err(dk, e), 1.000001f, 1e-30f);
                }
            }
        };
        int bl_ = bucket_of(xq), br_ = bl_;
        scan(bl_);
        for (;;) {
            // could a point left of / right of the scanned range still enter?  its |dx| exceeds the gap (minus the error)
            const float dkf = __double2float_ru(d[KNN - 1]);
            const float gl = fmaxf(xq - xmin - e, 0.f), gr = fmaxf(xmax - xq - e, 0.f);
            const bool goL = bl_ > 0 && !(gl * gl * 0.999999f > dkf);
            const bool goR = br_ < NB - 1 && !(gr * gr * 0.999999f > dkf);
            if (!goL && !goR) break;
            if (goL) scan(--bl_);
            if (goR) scan(++br_);
        }
        const int qi = q / step;
#pragma unroll
        for (int k = 0; k < KNN; ++k) o_base[(size_t)KNN * qi + k] = id[k];
    }
}

// =============================================================================================
// evaluate_p2p_matching (p2p_matching_in_organ/p2p_matching_evaluation.py:8-88) for a batch of segment pairs.
// All three scores are ratios of integer counts, so the kernels return the counts and the host divides:
// given equal maps the scores equal the reference's bit for bit.
//   cnt[6*s + 0..5] = landmark pairs within 2*threshold, landmarks, continuity passes, continuity samples,
//                     cycle passes, vertices of mesh2
// =============================================================================================
// get_mesh_landmarks (landmark_utils.py:5-9): nearest mesh vertex of every skeleton landmark, float64 brute force,
// one warp per landmark (lowest index on an exact tie).  lm_idx[b] for the unified landmark index b.
__global__ void __launch_bounds__(128) eval_landmark_nn_kernel(LipArgs a, int* __restrict__ lm_idx) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= a.B) return;
    const int m = a.bin_mesh[b];
    const int n = a.mesh_ubase[m + 1] - a.mesh_ubase[m];
    const double* __restrict__ xyz = mesh_xyz(a, m);
    const double* lm = mesh_lm(a, m) + 3 * (size_t)(b - a.mesh_bin[m]);
    const double l0 = lm[0], l1 = lm[1], l2 = lm[2];
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int v = lane; v < n; v += 32) {
        const double d = rdist64(l0, l1, l2, xyz[3 * (size_t)v], xyz[3 * (size_t)v + 1], xyz[3 * (size_t)v + 2]);
        if (d < bd) { bd = d; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(LIP_FULL, bd, o);
        const int oi = __shfl_xor_sync(LIP_FULL, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) lm_idx[b] = bi;
}

__device__ __forceinline__ double norm3(const double* __restrict__ p, const double* __restrict__ q) {
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// one CTA per segment pair; T21 / T12 are the caller's maps in batch order (local indices)
__global__ void __launch_bounds__(256) eval_scores_kernel(LipArgs a, const int* __restrict__ lm_idx, const int* __restrict__ nn2,
                                                          const int* __restrict__ nn2_base,
                                                          const long long* __restrict__ T21, const long long* __restrict__ T12,
                                                          double threshold, int sample_step, long long* __restrict__ cnt) {
    __shared__ int s_cnt[3];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int m1 = s, m2 = a.S + s;
    const int u1 = a.mesh_ubase[m1], u2 = a.mesh_ubase[m2] - a.n1_total;       // offsets into xyz1 / xyz2 and T12 / T21
    const int n2 = a.mesh_ubase[m2 + 1] - a.mesh_ubase[m2];
    const int K = a.mesh_bin[m1 + 1] - a.mesh_bin[m1];
    const double* __restrict__ v1 = mesh_xyz(a, m1);
    const double* __restrict__ v2 = mesh_xyz(a, m2);
    const long long* __restrict__ t21 = T21 + u2;
    if (tid < 3) s_cnt[tid] = 0;
    __syncthreads();
    // landmark precision (:35-37): |v1[T21[lm2]] - v1[lm1]| < 2 * threshold
    int c_lm = 0;
    for (int k = tid; k < K; k += 256) {
        const int i1 = lm_idx[a.mesh_bin[m1] + k], i2 = lm_idx[a.mesh_bin[m2] + k];
        if (norm3(v1 + 3 * (size_t)t21[i2], v1 + 3 * (size_t)i1) < 2.0 * threshold) ++c_lm;
    }
    // continuity (:41-57): every sample_step-th vertex of mesh2 and its two nearest vertices of mesh2 (KDTree k = 2,
    // self_knn_kernel<2>; the first is the vertex itself)
    int c_cont = 0;
    const int nsamp = (n2 + sample_step - 1) / sample_step;
    const int* __restrict__ nn = nn2 + 2 * (size_t)nn2_base[s];
    for (int i = tid; i < nsamp; i += 256) {
        const int i0 = nn[2 * (size_t)i], i1 = nn[2 * (size_t)i + 1];
        const long long ma = t21[i0], mb = (i1 >= 0) ? t21[i1] : ma;
        if (ma != mb && norm3(v1 + 3 * (size_t)ma, v1 + 3 * (size_t)mb) < threshold) ++c_cont;
    }
    // cycle consistency (:59-68): |v2[T12[T21[i]]] - v2[i]| < threshold
    int c_cyc = 0;
    if (T12) {
        const long long* __restrict__ t12 = T12 + u1;
        for (int i = tid; i < n2; i += 256)
            if (norm3(v2 + 3 * (size_t)t12[t21[i]], v2 + 3 * (size_t)i) < threshold) ++c_cyc;
    }
    atomicAdd(&s_cnt[0], c_lm); atomicAdd(&s_cnt[1], c_cont); atomicAdd(&s_cnt[2], c_cyc);
    __syncthreads();
    if (tid == 0) {
        long long* c = cnt + 6 * (size_t)s;
        c[0] = s_cnt[0]; c[1] = K; c[2] = s_cnt[1]; c[3] = (n2 + sample_step - 1) / sample_step; c[4] = s_cnt[2]; c[5] = n2;
    }
}

// =============================================================================================
// double_direction_refinement (p2p_matching_in_organ/p2p_matching_fm.py:75-132) for a batch of
// segment pairs.  The reference rewrites T21 in place while it iterates (np.unique / np.where are re-evaluated
// inside the loops), so the result depends on the visiting order: the kernel keeps that order — one warp per
// segment pair, lane 0 walking the targets in ascending order over inverse-index lists held in SMEM.
// =============================================================================================
// (the 5-NN self query of :78-79 is self_knn_kernel<5>; entries 1..4 of a row are the reference's NN_mesh1[t, 1:])
// The tables are 16-bit when every mesh of the batch has fewer than 65,535 vertices (IT = unsigned short, 0xFFFF = none):
// 2 * (3 n2 + 4 n1) bytes per segment pair, so that three maize-sized pairs share an SM instead of one; the mesh2
// coordinates stay in global memory (read only where two sources compete for a target).
template <typename IT>
struct RefineView {                                          // per-segment tables (SMEM when they fit, global scratch otherwise)
    IT *t21, *next21, *head21, *head12, *next12, *t12;
    const double* v2;                                        // 3 * n2, interleaved
    static constexpr IT NONE = (IT)~(IT)0;                   // -1 for int
};
__device__ __forceinline__ double refine_norm(const double* __restrict__ p, const double* __restrict__ q) {
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}
// the source of target t to keep: smallest |v2[ind] - v2[T12[t]]|, first on ties (:88-96, :113-121)
template <typename IT>
__device__ __forceinline__ IT refine_keep(const RefineView<IT>& w, int t) {
    const double* ref = w.v2 + 3 * (size_t)w.t12[t];
    double best = 10000.0;
    IT keep = w.head21[t];
    for (IT ind = w.head21[t]; ind != RefineView<IT>::NONE; ind = w.next21[ind]) {
        const double d = refine_norm(w.v2 + 3 * (size_t)ind, ref);
        if (d < best) { best = d; keep = ind; }
    }
    return keep;
}
// the same decision taken by the whole warp: lane l walks to position l of the list and fetches that source's coordinates, so the
// loads of all sources (and of the reference point) are in flight together instead of one L2 round trip after the other; the
// distances are the same float64 expressions, the winner is the first of the smallest (lowest position).  Uniform result.
// Lists of more than 32 sources take the serial walk.
template <typename IT>
__device__ __forceinline__ IT refine_keep_warp(const RefineView<IT>& w, int t, int lane) {
    constexpr IT NONE = RefineView<IT>::NONE;
    const IT head = w.head21[t];
    IT ind = head;
    for (int s = 0; s < lane && ind != NONE; ++s) ind = w.next21[ind];
    const bool longer = __shfl_sync(LIP_FULL, (int)(ind != NONE && w.next21[ind] != NONE), 31) != 0;
    if (longer) return refine_keep(w, t);                    // (every lane computes the same)
    const double* ref = w.v2 + 3 * (size_t)w.t12[t];
    double d = INFINITY;
    if (ind != NONE) d = refine_norm(w.v2 + 3 * (size_t)ind, ref);
    if (!(d < 10000.0)) d = INFINITY;                        // the reference starts from best = 10 ** 4: such a source never wins
    int pos = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(LIP_FULL, d, o);
        const int op = __shfl_xor_sync(LIP_FULL, pos, o);
        if (od < d || (od == d && op < pos)) { d = od; pos = op; }
    }
    const IT winner = (IT)__shfl_sync(LIP_FULL, (int)ind, pos);
    return d < INFINITY ? winner : head;
}
// move source ind from the list of t_from to its place (ascending) in the list of t_to
template <typename IT>
__device__ __forceinline__ void refine_move(const RefineView<IT>& w, IT ind, int t_from, int t_to) {
    IT* link = &w.head21[t_from];
    while (*link != ind) link = &w.next21[*link];
    *link = w.next21[ind];
    link = &w.head21[t_to];
    while (*link != RefineView<IT>::NONE && *link < ind) link = &w.next21[*link];
    w.next21[ind] = *link;
    *link = ind;
    w.t21[ind] = (IT)t_to;
}

// the refinement of one segment pair on tables at `base` (called once with the shared-memory tables and once with the
// global ones, so that the compiler emits shared-memory loads and stores for the usual case instead of generic ones)
// `hot`: the tables every step of the walk touches (T21, the source lists of every target, one snapshot bit per target:
// 2 n2 + n1 entries + n1 / 8 bytes); `cold`: T12 and its lists (n2 + 2 n1 entries), which a step only reads when it compares or
// re-routes sources.  The cold tables always live in the global scratch (they stay in L2); keeping only the hot ones in shared
// memory lets 7 instead of 3 maize-sized segment pairs share an SM, so that 64 frame pairs are one wave of warps instead of two
// (measured: 9.4 -> 8.4 ms; T12 in shared memory as well: 5 per SM, two waves again, 9.6 ms).
template <typename IT>
__device__ __forceinline__ void refine_pair(const LipArgs& a, IT* __restrict__ hot, IT* __restrict__ cold, const int* __restrict__ nn,
                                            long long* __restrict__ t21g, const long long* __restrict__ t12g, int u2, int n1, int n2, int lane) {
    constexpr IT NONE = RefineView<IT>::NONE;
    RefineView<IT> w;
    w.v2 = a.xyz2 + 3 * (size_t)u2;
    w.t21 = hot; w.next21 = hot + n2; w.head21 = hot + 2 * n2;
    w.head12 = cold; w.next12 = cold + n2; w.t12 = w.next12 + n1;
    unsigned* snap = reinterpret_cast<unsigned*>(hot + ((2 * (size_t)n2 + (size_t)n1 + 3) & ~(size_t)3));     // one bit per target, 4-byte aligned for either IT
    for (int i = lane; i < n2; i += 32) { w.t21[i] = (IT)t21g[i]; w.head12[i] = NONE; }
    for (int t = lane; t < n1; t += 32) { w.head21[t] = NONE; w.t12[t] = (IT)t12g[t]; }
    __syncwarp();
    if (lane == 0) {
        // inverse lists, ascending: np.where(T21 == t)[0] and np.where(T12 == ind)[0]
        for (int i = n2 - 1; i >= 0; --i) { const int t = w.t21[i]; w.next21[i] = w.head21[t]; w.head21[t] = (IT)i; }
        for (int t = n1 - 1; t >= 0; --t) { const int i = w.t12[t]; w.next12[t] = w.head12[i]; w.head12[i] = (IT)t; }
    }
    __syncwarp();
    for (int t0 = 0; t0 < n1; t0 += 32) {                                          // np.unique(T21), taken once (:85)
        const unsigned bits = __ballot_sync(LIP_FULL, t0 + lane < n1 && w.head21[t0 + lane] != NONE);
        if (lane == 0) snap[t0 >> 5] = bits;
    }
    __syncwarp();
    // Every lane walks the targets (all of them read the same table entries: broadcasts, uniform control flow); the loads a
    // decision waits for are spread over the lanes (refine_keep_warp), lane 0 alone rewrites the lists, in the reference's order.
    for (int t = 0; t < n1; ++t) {
        if (!((snap[t >> 5] >> (t & 31)) & 1u) || w.head21[t] == NONE || w.next21[w.head21[t]] == NONE) continue;   // a single source: nothing to do
        const IT keep = refine_keep_warp(w, t, lane);
        const int nrow = (lane < 4) ? nn[5 * (size_t)t + lane] : -1;     // the target's four neighbours, one load for the row
        IT ind = w.head21[t];
        while (ind != NONE) {
            const IT nxt = w.next21[ind];                    // the reference iterates over the list as it was
            if (ind != keep) {
                int t_keep = -1;
                if (w.head12[ind] != NONE) {                 // targets whose T12 points at this source (:100-101)
                    t_keep = w.head12[ind];
                    for (IT c = w.head12[ind]; c != NONE; c = w.next12[c]) if (w.head21[c] == NONE) t_keep = c;
                } else {                                     // the four nearest neighbours of the target (:102-103)
                    const int c0 = __shfl_sync(LIP_FULL, nrow, 0), c1 = __shfl_sync(LIP_FULL, nrow, 1), c2 = __shfl_sync(LIP_FULL, nrow, 2),
                              c3 = __shfl_sync(LIP_FULL, nrow, 3);
                    t_keep = c0;
                    if (c0 >= 0 && w.head21[c0] == NONE) t_keep = c0;
                    if (c1 >= 0 && w.head21[c1] == NONE) t_keep = c1;
                    if (c2 >= 0 && w.head21[c2] == NONE) t_keep = c2;
                    if (c3 >= 0 && w.head21[c3] == NONE) t_keep = c3;
                }
                if (t_keep >= 0 && t_keep != t) {
                    if (lane == 0) refine_move(w, ind, t, t_keep);
                    __syncwarp();                            // the lists as lane 0 left them, for every lane
                }
            }
            ind = nxt;
        }
    }
    __syncwarp();
    // targets nobody maps to take sources from their neighbours (:111-130).  The four neighbours of a target are
    // distinct, and moving the sources of one of them to the target does not touch the lists of the others, so the
    // four "which source stays" decisions are taken by four lanes at once; the moves stay in the reference's order.
    // The neighbour rows of eight targets are fetched with one coalesced load (lane l: entry l % 4 of target t0 + l / 4).
    for (int t0 = 0; t0 < n1; t0 += 8) {
        const int tr = t0 + (lane >> 2);
        const int row = (tr < n1) ? nn[5 * (size_t)tr + (lane & 3)] : -1;
        const int jn = min(8, n1 - t0);
        for (int j = 0; j < jn; ++j) {
            const int t = t0 + j;
            if (w.head21[t] != NONE) continue;               // uniform
            int nb = __shfl_sync(LIP_FULL, row, 4 * j + (lane & 3));
            IT keep = NONE;
            // (a neighbour with a single source keeps it whatever the distances are: nothing to compare, nothing to move)
            if (lane < 4 && nb >= 0 && nb != t && w.head21[nb] != NONE && w.next21[w.head21[nb]] != NONE) keep = refine_keep(w, nb); else nb = -1;
            __syncwarp();
            for (int k = 0; k < 4; ++k) {
                const int nbk = __shfl_sync(LIP_FULL, nb, k);
                const IT keepk = (IT)__shfl_sync(LIP_FULL, (int)keep, k);
                if (lane == 0 && nbk >= 0) {
                    IT ind = w.head21[nbk];
                    while (ind != NONE) {
                        const IT nxt = w.next21[ind];
                        if (ind != keepk) refine_move(w, ind, nbk, t);
                        ind = nxt;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    for (int i = lane; i < n2; i += 32) t21g[i] = (long long)w.t21[i];
}

// ALL_SMEM: the cold tables live in shared memory too (fewer segment pairs per SM, but every access is an LDS) — the host picks
// the layout that needs the fewer waves of CTAs for the batch at hand
template <typename IT, bool ALL_SMEM>
__global__ void __launch_bounds__(32) refine_kernel(LipArgs a, const int* __restrict__ nn5, long long* __restrict__ T21,
                                                    const long long* __restrict__ T12, IT* __restrict__ scratch, int smem_entries_cap) {
    extern __shared__ double smem_ref[];
    const int s = blockIdx.x, lane = threadIdx.x;
    const int u1 = a.mesh_ubase[s], n1 = a.mesh_ubase[s + 1] - u1;
    const int u2g = a.mesh_ubase[a.S + s], n2 = a.mesh_ubase[a.S + s + 1] - u2g, u2 = u2g - a.n1_total;
    long long* __restrict__ t21g = T21 + u2;
    const long long* __restrict__ t12g = T12 + u1;
    const int* __restrict__ nn = nn5 + 5 * (size_t)u1 + 1;          // entries 1..4 of every row: NN_mesh1[t, 1:]
    // tables: 3 n2 + 4 n1 + 16 entries per segment pair in the scratch: [hot: 2 n2 + n1 + snapshot bits | cold: n2 + 2 n1]
    IT* region = scratch + 3 * (size_t)u2 + 4 * (size_t)u1 + 16 * (size_t)s;
    const size_t need = 2 * (size_t)n2 + (size_t)n1 + 4 + ((size_t)n1 + 31) / 32 * (4 / sizeof(IT));
    const size_t need_all = need + (size_t)n2 + 2 * (size_t)n1;
    IT* sm = reinterpret_cast<IT*>(smem_ref);
    if (ALL_SMEM && need_all <= (size_t)smem_entries_cap) refine_pair<IT>(a, sm, sm + need, nn, t21g, t12g, u2, n1, n2, lane);
    else if (need <= (size_t)smem_entries_cap) refine_pair<IT>(a, sm, region + need, nn, t21g, t12g, u2, n1, n2, lane);
    else refine_pair<IT>(a, region, region + need, nn, t21g, t12g, u2, n1, n2, lane);
}


// =============================================================================================
// interpolation_segment_level (interpolation/interpolation_pcd.py:122-134): for every segment pair the rows
//   (1 - alpha) * xyz1[T21] + alpha * xyz2        (n2 rows, mesh2 pulled back onto its matches)
//   (1 - alpha) * xyz1      + alpha * xyz2[T12]   (n1 rows, mesh1 pushed towards its matches)
// stacked in that order.  numpy evaluates the two products and the sum separately (no contraction): so do we.
// A negative map entry counts from the end, as numpy's fancy indexing does; anything else out of range sets err.
// =============================================================================================
__global__ void __launch_bounds__(256) interpolate_kernel(LipArgs a, const long long* __restrict__ T21, const long long* __restrict__ T12,
                                                          double alpha, double* __restrict__ out, int* __restrict__ err) {
    const int s = blockIdx.x;
    const int u1 = a.mesh_ubase[s], n1 = a.mesh_ubase[s + 1] - u1;
    const int u2 = a.mesh_ubase[a.S + s] - a.n1_total, n2 = a.mesh_ubase[a.S + s + 1] - a.n1_total - u2;
    const double* __restrict__ x1 = a.xyz1 + 3 * (size_t)u1;
    const double* __restrict__ x2 = a.xyz2 + 3 * (size_t)u2;
    double* __restrict__ o = out + 3 * ((size_t)u1 + (size_t)u2);
    const double w1 = 1.0 - alpha;
    for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < n1 + n2; r += gridDim.y * blockDim.x) {
        long long i, j;
        if (r < n2) { j = r; i = T21[u2 + r]; if (i < 0) i += n1; }
        else { i = r - n2; j = T12[u1 + (r - n2)]; if (j < 0) j += n2; }
        if (i < 0 || i >= n1 || j < 0 || j >= n2) { atomicExch(err, 1); continue; }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            o[3 * (size_t)r + k] = __dadd_rn(__dmul_rn(w1, x1[3 * (size_t)i + k]), __dmul_rn(alpha, x2[3 * (size_t)j + k]));
    }
}

// maps handed to the refinement / the scores are used as indices on the device: every entry of T21 (off2[S]) must address
// a vertex of its segment pair's mesh1, every entry of T12 (optional) one of mesh2; *err = 1 otherwise
__global__ void __launch_bounds__(256) map_check_kernel(LipArgs a, const long long* __restrict__ T21, const long long* __restrict__ T12,
                                                        int* __restrict__ err) {
    const int s = blockIdx.x;
    const int u1 = a.mesh_ubase[s], n1 = a.mesh_ubase[s + 1] - u1;
    const int u2 = a.mesh_ubase[a.S + s] - a.n1_total, n2 = a.mesh_ubase[a.S + s + 1] - a.n1_total - u2;
    bool bad = false;
    for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < n1 + n2; r += gridDim.y * blockDim.x) {
        if (r < n2) { const long long v = T21[u2 + r]; bad |= v < 0 || v >= n1; }
        else if (T12) { const long long v = T12[u1 + (r - n2)]; bad |= v < 0 || v >= n2; }
    }
    if (bad) atomicExch(err, 1);
}
