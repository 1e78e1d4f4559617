// liblocalicp.so — C ABI (include/local_icp.h) over the sm_100a kernels in lip_kernels.cuh.
// Host side: batch planning (CSR -> mesh / bin / tile descriptors), workspace, launches.
#include "../../include/local_icp.h"
#include "lip_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>      // header-only; ranges cost a pointer test unless a profiler is attached

#ifndef LIP_ARCH
#define LIP_ARCH 100
#endif
#define LIP_PIPELINE_CHUNKS 8             // sub-batches of the pipelined host path
#define LIP_PIPELINE_RAMP 1.6               // sub-batch k ends where the running vertex count passes (k/chunks)^ramp of the total: the first upload is short, the late (GPU-bound) sub-batches large
#define LIP_DEVICE_CHUNKS 2               // sub-batches of the device-resident entry point
#define LIP_PIPELINE_MIN_VERTICES 400000   // smaller batches run as one piece

namespace {

// NVTX range on the calling thread (tracing hook): entry points and the stages they enqueue
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

constexpr size_t kGroupSmem = sizeof(IcpGroupSmem) * LIP_ICP_WARPS;     // dynamic SMEM of icp_group_kernel (> 48 KB: opt-in)
static_assert(LIP_ICP_CHUNK <= 256, "winners of a staged partition are stored as one byte");

std::string g_startup_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// pinned host buffer, grow-only (staging of pageable caller arrays)
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A few host threads that copy between the caller's pageable arrays and pinned staging buffers, so that the copy engines
// are fed at PCIe speed while the caller passes ordinary numpy arrays (a cudaMemcpyAsync from pageable memory is staged
// by the driver on the calling thread at a fraction of that).  Jobs are (dst, src, bytes) slices counted per group.
struct StagePool {
    struct Job { void* dst; const void* src; size_t bytes; std::atomic<int>* left; std::atomic<int>* bad; };
    // any NaN / infinity among n doubles (exponent bits all ones)
    static bool nonfinite(const void* src, size_t bytes) {
        const uint64_t* p = (const uint64_t*)src;
        uint64_t hit = 0;
        for (size_t i = 0; i < bytes / 8; ++i) hit |= (uint64_t)((p[i] & 0x7ff0000000000000ull) == 0x7ff0000000000000ull);
        return hit != 0;
    }
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    std::deque<Job> q;
    bool stop = false;
    void start(int n) {
        for (int i = (int)th.size(); i < n; ++i) th.emplace_back([this] { run(); });
    }
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv.wait(lk, [this] { return stop || !q.empty(); });
                if (q.empty()) return;
                j = q.front(); q.pop_front();
            }
            if (j.dst) memcpy(j.dst, j.src, j.bytes);
            if (j.bad && nonfinite(j.dst ? j.dst : j.src, j.bytes)) j.bad->store(1, std::memory_order_relaxed);
            if (j.left->fetch_sub(1, std::memory_order_acq_rel) == 1) { std::lock_guard<std::mutex> lk(m); cv_done.notify_all(); }
        }
    }
    // slices of at most `slice` bytes; *left must already count them (slices_of)
    static int slices_of(size_t bytes, size_t slice) { return (int)((bytes + slice - 1) / slice); }
    // dst == nullptr: no copy (check only); bad != nullptr: set to 1 if the slice (of doubles) holds a NaN or an infinity
    void submit(void* dst, const void* src, size_t bytes, size_t slice, std::atomic<int>* left, std::atomic<int>* bad = nullptr) {
        std::lock_guard<std::mutex> lk(m);
        for (size_t o = 0; o < bytes; o += slice)
            q.push_back({dst ? (char*)dst + o : nullptr, (const char*)src + o, std::min(slice, bytes - o), left, bad});
        cv.notify_all();
    }
    bool done(const std::atomic<int>& left) const { return left.load(std::memory_order_acquire) == 0; }
    void wait(const std::atomic<int>& left) {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return left.load(std::memory_order_acquire) == 0; });
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; q.clear(); }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};

enum Stage { ST_PARTITION = 1, ST_ICP = 2, ST_GRID = 4, ST_BLEND = 16, ST_QUERY = 32, ST_GRID_ALL = 64 };

}  // namespace

// one recorded step (run_stages): the arguments it was captured with and the executable graph
struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    LipArgs args;
    uint64_t plan_gen = 0;
    int stages = 0, seen = 0;
    cudaStream_t st = nullptr;
    int64_t launches = 0;
    void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; seen = 0; }
};

struct lip_ctx {
    int device = 0;
    int sm_count = 0;
    int group_blocks_per_sm = 1, big_blocks_per_sm = 1, map_blocks_per_sm = 1;
    cudaStream_t stream = nullptr;
    cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};   // concurrent branches: [0] CTA ICP kernel, [1] uploads of the host pipeline, [2] grid build, [3] cluster ICP kernel
    cudaEvent_t ev_begin = nullptr, ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    int cluster_size = LIP_CLUSTER_SIZE, cluster_cls = LIP_CLUSTER_CLASS, cluster_grid = 0;
    int64_t heavy_bound = 0;                                 // planned batch: upper bound on the partitions the cluster kernel can get
    bool rows_on_device = false;                             // d_rows holds h_rows (series_rows), uploaded on rows_stream
    cudaStream_t rows_stream = nullptr;
    StepGraph step_graph;                                    // the last step of this context, as a CUDA graph once it has repeated
    bool graphs_ok = true;                                   // false after a capture / instantiate / launch of a graph failed
    uint64_t plan_gen = 0;                                   // bumped by every plan_impl
    bool use_clusters = false;                               // planned batch: small, with whole-organ ICPs (see plan_impl)
    int64_t cluster_max_bins = LIP_CLUSTER_COOP_PARTS;
    bool fill_done = false;
    int map_mode = 0;        // final map: 0 = landmark-bin tiles, 1 = hashed grid (lip_set_map_mode)
    int map_bytes = 8;       // entry size of the dense maps of lip_batch_run_* / lip_series_run_host: 8 (int64, numpy intp) or 4 (lip_set_map_dtype)
    DevBuf in_frames, d_rows;   // lip_series_run_host: every frame of the series once; first row of each of this context's meshes in it
    std::vector<int> h_rows;
    cudaEvent_t ev_piece = nullptr;
    size_t max_map_tiles = 0;
    // host-path pipeline: the batch split into sub-batches, each on its own child context (own stream and
    // workspace), so that H2D of one, kernels of another and D2H of a third overlap (lip_batch_run_host)
    std::vector<lip_ctx*> kids, dkids;  // sub-batches of the host path (LIP_PIPELINE_CHUNKS) / of the device path (LIP_DEVICE_CHUNKS)
    std::vector<int64_t> kid_seg, dkid_seg;   // first segment pair of each kid, and S at the end
    int n_dkids = 0;
    bool last_run_dsplit = false;
    std::vector<int64_t> h_off1, h_off2, h_lm_off;   // the planned CSR (for slicing the host arrays)
    std::vector<int> h_qbase_eval, h_qbase_ref;      // first query of every mesh (self_knn_kernel), uploaded asynchronously
    int n_kids = 0;                     // kids in use by the current plan (0: no split)
    bool last_run_split = false;
    bool is_kid = false;
    lip_params params;
    std::string err;
    // plan.  planned: the plan arrays / workspace describe SOME batch (plan_impl); user_plan: that batch is the caller's
    // (lip_batch_plan) and not an internal one of lip_icp_host / lip_nn_host, which clear it before they do anything else
    bool planned = false, user_plan = false;
    int64_t S = 0, n1 = 0, n2 = 0, sumK = 0, maxK = 0, max_mesh = 0;
    int M = 0, B = 0, n_tiles = 0, tab_total = 0, directions = 3;
    std::vector<int> h_mesh_ubase, h_mesh_bin, h_mesh_tab, h_bin_mesh, h_tile_mesh, h_tile_first, h_plan;
    std::vector<int> h_ptile_mesh, h_ptile_first, h_ptile_cnt, h_mesh_ptile;   // tiles of the partition kernels
    int n_ptiles = 0;
    int64_t tile_cnt_total = 0;
    int *ptile_mesh = nullptr, *ptile_first = nullptr, *ptile_cnt = nullptr, *mesh_ptile = nullptr;
    DevBuf d_plan;      // all plan arrays in one allocation
    int *mesh_ubase = nullptr, *mesh_bin = nullptr, *mesh_tab = nullptr, *bin_mesh = nullptr, *tile_mesh = nullptr,
        *tile_first = nullptr;
    // workspace
    DevBuf ws;
    LipArgs args;
    // device copies of host-API inputs / outputs
    DevBuf in_xyz1, in_xyz2, in_lm1, in_lm2, out_T21, out_T12, misc, misc2;
    int64_t last_launches = 0;
    // optional per-stage CUDA-event timing (lip_set_profiling)
    bool profiling = false;
    cudaEvent_t ev[LIP_N_STAGES + 1] = {};
    bool ev_valid = false;
    cudaEvent_t tr[4] = {};       // LIP_TRACE_PIPELINE: H2D start / end, kernels end, D2H end of the last host run
    // staging of pageable caller arrays (host pipeline): per sub-batch pinned buffers, a small pool of copy threads
    HostBuf st_in, st_out;
    StagePool* pool = nullptr;
    bool validate_inputs = false;   // lip_set_input_validation
    int plan_chunks = 0, plan_dchunks = 0;   // sub-batch counts the current plan was made with
    std::vector<cudaEvent_t> stage_events;   // per-chunk download events of StageCall
    DevBuf d_err; HostBuf h_err; cudaEvent_t ev_chk = nullptr;   // map_check_kernel: flag, its pinned copy, "copy done"
    HostBuf h_counts;                        // score counts of this (sub-)batch, pinned: their download must not block the host
};

namespace {

int fail(lip_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_startup_error = msg;
    return code;
}
#define LIP_CUDA(c, call)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, LIP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

int next_pow2(int64_t x) { int p = 1; while (p < x) p <<= 1; return p; }

// carve the workspace and fill the LipArgs pointers that do not depend on the run
int carve_workspace(lip_ctx* c) {
    const size_t n = (size_t)(c->n1 + c->n2), B = (size_t)c->B, M = (size_t)c->M;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t o_nn0 = take(n * 4), o_nn1 = take(n * 4), o_d0 = take(n * 8), o_d1 = take(n * 8);
    const size_t o_binoff = take((B + 1) * 4), o_vid = take(n * 4), o_nnidx = take(n * 4);
    const size_t o_R = take(B * 72), o_t = take(B * 24), o_it = take(B * 4), o_stat = take(B * 4);
    const size_t o_fR = take(B * 72), o_ft = take(B * 24), o_bdesc = take(B * 128), o_frames = take(B * sizeof(IcpFrame));
    const size_t o_sched = take(32 * 4), o_warp = take(8 * B * 4), o_big = take(4 * B * 4), o_bsum = take(B * 32);
    const size_t o_far = take(n * 8), o_farq = take(n * 24), o_re = take(n * 8), o_req = take(n * 24);
    const size_t o_pts = take(n * 16), o_tab = take((size_t)c->tab_total * 4), o_hdr = take(M * sizeof(GridHdr));
    const size_t o_cnt = take(8 * 8);
    const size_t o_bpts = take(n * 16), o_bbox = take(B * 32), o_sbox = take(B * 32), o_msum = take(M * 32);
    c->max_map_tiles = n / LIP_MAP_TILE + B + 1;
    const size_t o_bxyz = take(n * 24), o_bq = take(n * 24), o_pb = take(n * 4), o_pn = take(n * 4), o_pd = take(n * 16);
    const size_t o_mt = take(c->max_map_tiles * sizeof(MapTile));
    const size_t o_tcnt = take((size_t)c->tile_cnt_total * 4), o_tsum = take((size_t)c->n_ptiles * 32);
    const size_t o_tbase = take((size_t)c->tile_cnt_total * 4), o_btmp = take(n * 16);
    LIP_CUDA(c, c->ws.ensure(off));
    char* base = c->ws.as<char>();
    LipArgs& a = c->args;
    a.nn0 = (int*)(base + o_nn0); a.nn1 = (int*)(base + o_nn1);
    a.d0 = (double*)(base + o_d0); a.d1 = (double*)(base + o_d1);
    a.bin_off = (int*)(base + o_binoff); a.part_vid = (int*)(base + o_vid); a.nn_idx = (int*)(base + o_nnidx);
    a.icp_R = (double*)(base + o_R); a.icp_t = (double*)(base + o_t);
    a.icp_iters = (int*)(base + o_it); a.icp_status = (int*)(base + o_stat);
    a.fill_R = (double*)(base + o_fR); a.fill_t = (double*)(base + o_ft); a.bin_desc = (double*)(base + o_bdesc); a.icp_frames = (IcpFrame*)(base + o_frames);
    a.sched = (int*)(base + o_sched); a.warp_list = (int*)(base + o_warp); a.bin_sum = (double*)(base + o_bsum);
    a.big_list = (int*)(base + o_big); a.far_list = (int*)(base + o_far); a.far_q = (double*)(base + o_farq);
    a.retry_list = (int*)(base + o_re); a.retry_q = (double*)(base + o_req);
    a.grid_pts = (float4*)(base + o_pts); a.grid_start = (int*)(base + o_tab); a.grid_hdr = (GridHdr*)(base + o_hdr);
    a.counters = (unsigned long long*)(base + o_cnt);
    a.bin_pts = (float4*)(base + o_bpts); a.bin_box = (float*)(base + o_bbox); a.sub_box = (float4*)(base + o_sbox); a.mesh_sum = (double*)(base + o_msum);
    a.map_tiles = (MapTile*)(base + o_mt);
    a.tile_cnt = (int*)(base + o_tcnt); a.tile_sum = (double*)(base + o_tsum);
    a.tile_base = (int*)(base + o_tbase); a.bin_tmp = (float4*)(base + o_btmp);
    a.bin_xyz = (double*)(base + o_bxyz); a.bin_q = (double*)(base + o_bq);
    a.pos_bin = (int*)(base + o_pb); a.pos_nn1 = (int*)(base + o_pn); a.pos_d = (double*)(base + o_pd);
    return LIP_OK;
}

void set_param_args(lip_ctx* c) {
    LipArgs& a = c->args;
    a.max_iter = c->params.max_iter;
    a.rise_tol = c->params.cost_rise_tol;
    a.flat_tol = c->params.cost_flat_tol;
    a.obtuse_cos = std::sin(c->params.obtuse_eps);          // C > pi/2 - eps  <=>  cos C < sin(eps)
    a.cos_den_eps = c->params.cos_den_eps;
    a.reject_dist = c->params.reject_dist;
    a.profile = c->profiling ? 1 : 0;
    a.map_mode = c->map_mode;
    a.map32 = c->map_bytes == 4 ? 1 : c->map_bytes == 2 ? 2 : 0;
    static const int tpw = getenv("LIP_ICP_TASKS_PER_WARP") ? std::max(1, atoi(getenv("LIP_ICP_TASKS_PER_WARP"))) : LIP_ICP_TASKS_PER_WARP;
    static const int csh = getenv("LIP_ICP_CLASS_SHIFT") ? atoi(getenv("LIP_ICP_CLASS_SHIFT")) : LIP_ICP_CLASS_SHIFT;
    a.icp_tasks_per_warp = tpw; a.icp_class_shift = csh;
    a.big_cluster_cls = c->use_clusters ? c->cluster_cls : 4;
    a.cluster_coop_parts = (int)std::min<int64_t>(c->cluster_max_bins, INT32_MAX);
    // warp / CTA split of the ICP partitions: large batches keep the (more efficient) warp kernel busy with up to 96 k pairs
    // per partition, small ones (the sub-batches of the host pipeline) hand more to the CTA kernel, whose tails are shorter
    static const long long ww = getenv("LIP_WARP_WORK") ? atoll(getenv("LIP_WARP_WORK")) : 0;
    a.warp_work = ww > 0 ? ww : (c->B >= LIP_WARP_WORK_BINS ? LIP_WARP_WORK_LARGE : LIP_WARP_WORK);
}

int plan_impl(lip_ctx* c, int64_t S, const int64_t* off1, const int64_t* off2, const int64_t* lm_off, int directions) {
    const int tile = LIP_TILE;                               // tiles of the grid-mode kernels (blend_nn_kernel, nn_query_kernel)
    if (!c) return LIP_EINVAL;
    c->planned = false;
    ++c->plan_gen;
    if (S <= 0 || !off1 || !off2 || !lm_off) return fail(c, LIP_EINVAL, "plan: S must be > 0 and offsets non-null");
    if ((directions & LIP_DIR_BOTH) == 0) return fail(c, LIP_EINVAL, "plan: directions must include T21 and/or T12");
    if (off1[0] != 0 || off2[0] != 0 || lm_off[0] != 0) return fail(c, LIP_EINVAL, "plan: offsets must start at 0");
    int64_t maxK = 0, max_mesh = 0, heavy_bound = 0;
    bool heavy_hint = false;
    static const int64_t kClassFloor[5] = {0, 150000, 400000, 1000000, INT64_MAX / 2};   // icp_schedule_kernel's classes
    const int64_t heavy_thr = kClassFloor[c->cluster_cls];
    for (int64_t s = 0; s < S; ++s) {
        if (off1[s + 1] <= off1[s] || off2[s + 1] <= off2[s])
            return fail(c, LIP_EINVAL, "plan: segment pair " + std::to_string(s) + " has an empty mesh");
        int64_t K = lm_off[s + 1] - lm_off[s];
        if (K <= 0) return fail(c, LIP_EINVAL, "plan: segment pair " + std::to_string(s) + " has no landmarks");
        if (K > LIP_MAX_LANDMARKS) return fail(c, LIP_EINVAL, "plan: more than LIP_MAX_LANDMARKS landmarks in a segment");
        maxK = std::max(maxK, K);
        max_mesh = std::max(max_mesh, std::max(off1[s + 1] - off1[s], off2[s + 1] - off2[s]));
        // partitions of this pair that can exceed the cluster kernel's work threshold: sum of n_src * n_tgt over its bins <= n1 * n2
        const int64_t pairs = (off1[s + 1] - off1[s]) * (off2[s + 1] - off2[s]);
        heavy_bound += 2 * std::min<int64_t>(K, pairs / (heavy_thr + 1));
        heavy_hint = heavy_hint || pairs / (K * K) > heavy_thr / 4;      // the AVERAGE bin pair of this segment pair is near the threshold
    }
    c->heavy_bound = heavy_bound;
    // The cluster kernel pays in batches whose ICP stage is the serial chain of a few whole-organ ICPs (one frame pair of small
    // organs, as the reference's driver submits them).  In large batches the GPU is busy anyway and a co-operating cluster costs
    // more SM time than a lone CTA; in batches without such partitions its launch is ~10 us of overhead.  Both keep the CTA kernel.
    c->use_clusters = heavy_hint && heavy_bound > 0 && !c->is_kid && c->cluster_cls < 4 && c->cluster_grid > 0 &&
                      2 * lm_off[S] <= c->cluster_max_bins;
    if (c->map_bytes == 2 && max_mesh > 65535 && !c->is_kid)
        return fail(c, LIP_EINVAL, "plan: 16-bit maps (lip_set_map_dtype) need every mesh to have fewer than 65,536 vertices");
    c->max_mesh = max_mesh;
    const int64_t n1 = off1[S], n2 = off2[S], sumK = lm_off[S];
    if (n1 + n2 >= (int64_t)1 << 30 || 2 * S >= (int64_t)1 << 28)
        return fail(c, LIP_EINVAL, "plan: batch too large for 32-bit indexing; split it");
    c->S = S; c->n1 = n1; c->n2 = n2; c->sumK = sumK; c->maxK = maxK;
    c->M = (int)(2 * S); c->B = (int)(2 * sumK); c->directions = directions & LIP_DIR_BOTH;
    const int M = c->M;
    c->h_mesh_ubase.assign(M + 1, 0); c->h_mesh_bin.assign(M + 1, 0); c->h_mesh_tab.assign(M + 1, 0);
    c->h_bin_mesh.assign(c->B, 0);
    for (int64_t s = 0; s <= S; ++s) {
        if (s < S) {
            c->h_mesh_ubase[s] = (int)off1[s]; c->h_mesh_ubase[S + s] = (int)(n1 + off2[s]);
            c->h_mesh_bin[s] = (int)lm_off[s]; c->h_mesh_bin[S + s] = (int)(sumK + lm_off[s]);
        }
    }
    c->h_mesh_ubase[M] = (int)(n1 + n2);
    c->h_mesh_bin[M] = (int)(2 * sumK);
    int64_t tab = 0;
    for (int m = 0; m < M; ++m) {
        c->h_mesh_tab[m] = (int)tab;
        int64_t n = c->h_mesh_ubase[m + 1] - c->h_mesh_ubase[m];
        tab += next_pow2(std::max<int64_t>(n, 32)) + 2;
        if (tab >= (int64_t)1 << 30) return fail(c, LIP_EINVAL, "plan: hash tables too large");
        for (int b = c->h_mesh_bin[m]; b < c->h_mesh_bin[m + 1]; ++b) c->h_bin_mesh[b] = m;
    }
    c->h_mesh_tab[M] = (int)tab;
    c->tab_total = (int)tab;
    // tiles of the per-vertex kernels: source meshes of the requested directions
    c->h_tile_mesh.clear(); c->h_tile_first.clear();
    for (int m = 0; m < M; ++m) {
        bool active = (m >= S) ? (c->directions & 1) : (c->directions & 2);
        if (!active) continue;
        int n = c->h_mesh_ubase[m + 1] - c->h_mesh_ubase[m];
        for (int f = 0; f < n; f += tile) { c->h_tile_mesh.push_back(m); c->h_tile_first.push_back(f); }
    }
    c->n_tiles = (int)c->h_tile_mesh.size();
    // tiles of the partition kernels: LIP_PART_CHUNK vertices of one mesh, every mesh; per tile one count per landmark of its mesh
    c->h_ptile_mesh.clear(); c->h_ptile_first.clear(); c->h_ptile_cnt.clear(); c->h_mesh_ptile.assign(M + 1, 0);
    int64_t cnt_total = 0;
    for (int m = 0; m < M; ++m) {
        c->h_mesh_ptile[m] = (int)c->h_ptile_mesh.size();
        const int n = c->h_mesh_ubase[m + 1] - c->h_mesh_ubase[m], K = c->h_mesh_bin[m + 1] - c->h_mesh_bin[m];
        for (int f = 0; f < n; f += LIP_PART_CHUNK) {
            c->h_ptile_mesh.push_back(m); c->h_ptile_first.push_back(f); c->h_ptile_cnt.push_back((int)cnt_total);
            cnt_total += K;
            if (cnt_total >= (int64_t)1 << 30) return fail(c, LIP_EINVAL, "plan: too many (tile, landmark) counters; split the batch");
        }
    }
    c->h_mesh_ptile[M] = (int)c->h_ptile_mesh.size();
    c->h_ptile_cnt.push_back((int)cnt_total);
    c->n_ptiles = (int)c->h_ptile_mesh.size();
    c->tile_cnt_total = cnt_total;
    // upload
    size_t off = 0;
    auto take = [&](size_t cnt) { size_t o = off; off = align_up(off + cnt * 4); return o; };
    const size_t o0 = take(M + 1), o1 = take(M + 1), o2 = take(M + 1), o3 = take(c->B), o4 = take(c->n_tiles), o5 = take(c->n_tiles);
    const size_t o6 = take(c->n_ptiles), o7 = take(c->n_ptiles), o8 = take(c->n_ptiles + 1), o9 = take(M + 1);
    LIP_CUDA(c, c->d_plan.ensure(off));
    char* base = c->d_plan.as<char>();
    c->mesh_ubase = (int*)(base + o0); c->mesh_bin = (int*)(base + o1); c->mesh_tab = (int*)(base + o2);
    c->bin_mesh = (int*)(base + o3); c->tile_mesh = (int*)(base + o4); c->tile_first = (int*)(base + o5);
    c->ptile_mesh = (int*)(base + o6); c->ptile_first = (int*)(base + o7); c->ptile_cnt = (int*)(base + o8); c->mesh_ptile = (int*)(base + o9);
    // one upload for all six arrays; complete on return, whatever stream the kernels will run on
    c->h_plan.assign(off / 4, 0);
    auto put = [&](size_t o, const std::vector<int>& h) { if (!h.empty()) memcpy(c->h_plan.data() + o / 4, h.data(), h.size() * 4); };
    put(o0, c->h_mesh_ubase); put(o1, c->h_mesh_bin); put(o2, c->h_mesh_tab); put(o3, c->h_bin_mesh); put(o4, c->h_tile_mesh); put(o5, c->h_tile_first);
    put(o6, c->h_ptile_mesh); put(o7, c->h_ptile_first); put(o8, c->h_ptile_cnt); put(o9, c->h_mesh_ptile);
    LIP_CUDA(c, cudaMemcpyAsync(base, c->h_plan.data(), off, cudaMemcpyHostToDevice, c->stream));
    LIP_CUDA(c, cudaStreamSynchronize(c->stream));
    int rc = carve_workspace(c);
    if (rc) return rc;
    LipArgs& a = c->args;
    a.S = (int)S; a.M = M; a.B = c->B; a.sumK = (int)sumK; a.n1_total = (int)n1; a.n_total = (int)(n1 + n2);
    a.dir_mask = c->directions; a.n_tiles = c->n_tiles;
    a.mesh_ubase = c->mesh_ubase; a.mesh_bin = c->mesh_bin; a.mesh_tab = c->mesh_tab; a.bin_mesh = c->bin_mesh;
    a.tile_mesh = c->tile_mesh; a.tile_first = c->tile_first;
    a.ptile_mesh = c->ptile_mesh; a.ptile_first = c->ptile_first; a.ptile_cnt = c->ptile_cnt; a.mesh_ptile = c->mesh_ptile;
    a.q_dist = nullptr;
    c->planned = true;
    return LIP_OK;
}

// the launches of one step (c->args filled in by run_stages)
int launch_stages(lip_ctx* c, int stages, cudaStream_t st) {
    LipArgs& a = c->args;
    int64_t launches = 0;
    int evi = 0;
    // While profiling, everything runs back to back on `st` so that the events bracket one kernel each;
    // otherwise the CTA ICP kernel (and the grid build of map mode 1) run as concurrent branches.
    const bool serial = c->profiling;
    auto mark = [&]() { if (c->profiling) cudaEventRecord(c->ev[evi], st); ++evi; };
    // LIP_DEBUG_SYNC=1: synchronise after every launch and name the kernel that faulted
    static const bool debug_sync = getenv("LIP_DEBUG_SYNC") != nullptr;
    auto check = [&](const char* what) -> int {
        if (!debug_sync) return LIP_OK;
        cudaError_t e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = cudaGetLastError();
        return e == cudaSuccess ? LIP_OK : fail(c, LIP_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    };
#define LIP_CHECK(what) do { int rc_ = check(what); if (rc_) return rc_; } while (0)
    cudaStream_t s_big = serial ? st : c->side[0], s_grid = serial ? st : c->side[2];
    c->ev_valid = false;
    c->fill_done = false;
    LIP_CUDA(c, cudaMemsetAsync(a.sched, 0, 32 * 4, st));
    LIP_CUDA(c, cudaMemsetAsync(a.counters, 0, 8 * 8, st));
    if (!serial) LIP_CUDA(c, cudaEventRecord(c->ev_begin, st));
    mark();                                                  // event 0
    if (stages & ST_PARTITION) {
        Range r_("lip:partition");
        landmark_nn_kernel<<<c->n_ptiles, LIP_PART_THREADS, (size_t)c->maxK * 44 + 32, st>>>(a);
        LIP_CHECK("landmark_nn_kernel");
        partition_scan_kernel<<<(c->M + 3) / 4, 128, 0, st>>>(a);
        LIP_CHECK("partition_scan_kernel");
        partition_place_kernel<<<c->n_ptiles, LIP_PART_THREADS, (size_t)c->maxK * 4 * (2 + LIP_PART_THREADS / 32) + 16, st>>>(a);
        LIP_CHECK("partition_place_kernel");
        bin_sort_kernel<<<(c->B + LIP_SORT_WARPS - 1) / LIP_SORT_WARPS, LIP_SORT_WARPS * 32, 0, st>>>(a);
        launches += 4;
        LIP_CHECK("bin_sort_kernel");
    }
    mark();                                                  // 1: after partition
    bool forked_icp = false, forked_grid = false;
    if (stages & ST_ICP) {
        Range r_("lip:icp");
        icp_schedule_kernel<<<(c->B + 255) / 256, 256, 0, st>>>(a);
        LIP_CHECK("icp_schedule_kernel");
        mark();                                              // 2: after schedule
        if (!serial) {
            LIP_CUDA(c, cudaEventRecord(c->ev_fork, st));
            LIP_CUDA(c, cudaStreamWaitEvent(s_big, c->ev_fork, 0));
            forked_icp = true;
        }
        // the CTA-per-partition kernel (few, long tasks) runs as a concurrent branch of the group kernel
        int64_t bound = std::min<int64_t>(c->B, 2 * (int64_t)a.n_total / 256 + 1);     // heavy partitions hold > 256 points
        int gb = (int)std::max<int64_t>(1, std::min<int64_t>(bound, (int64_t)c->sm_count * c->big_blocks_per_sm));
        int gs = c->sm_count * c->group_blocks_per_sm;
        gs = std::max(1, std::min(gs, (c->B + LIP_ICP_WARPS - 1) / LIP_ICP_WARPS));
        // ... and the heaviest of those on thread-block clusters (a partition of > 400 k pairs holds > 632 points on one side)
        const bool clusters = a.big_cluster_cls < 4;
        auto launch_clusters = [&](cudaStream_t s) -> cudaError_t {
            const int64_t ncl = std::max<int64_t>(1, std::min<int64_t>(c->heavy_bound, c->cluster_grid / c->cluster_size));
            cudaLaunchConfig_t q = {};
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)c->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.gridDim = dim3((unsigned)(ncl * c->cluster_size)); q.blockDim = dim3(LIP_CLUSTER_THREADS); q.stream = s;
            q.attrs = at; q.numAttrs = 1;
            return cudaLaunchKernelEx(&q, icp_cluster_kernel<LIP_CLUSTER_THREADS, LIP_CLUSTER_CHUNK>, a);
        };
        if (serial) {
            icp_group_kernel<<<gs, LIP_ICP_WARPS * 32, kGroupSmem, st>>>(a);
            mark();                                          // 3: after the group kernel
            if (clusters) LIP_CUDA(c, launch_clusters(st));
            icp_big_kernel<<<gb, LIP_BIG_THREADS, 0, st>>>(a);
        } else {
            if (clusters) {
                LIP_CUDA(c, cudaStreamWaitEvent(c->side[3], c->ev_fork, 0));
                LIP_CUDA(c, launch_clusters(c->side[3]));
                LIP_CUDA(c, cudaEventRecord(c->ev_join[3], c->side[3]));
            }
            icp_big_kernel<<<gb, LIP_BIG_THREADS, 0, s_big>>>(a);
            icp_group_kernel<<<gs, LIP_ICP_WARPS * 32, kGroupSmem, st>>>(a);
            mark();
            if (clusters) LIP_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[3], 0));
        }
        launches += clusters ? 4 : 3;
        LIP_CHECK("icp_group_kernel / icp_big_kernel");
    } else { mark(); mark(); }
    mark();                                                  // 4: after the CTA ICP kernel
    const bool tiles = (stages & ST_BLEND) && c->map_mode != 1;      // final map on landmark-bin tiles: no grid needed
    if ((stages & (ST_GRID | ST_GRID_ALL)) && !tiles) {
        if (!serial) { LIP_CUDA(c, cudaStreamWaitEvent(s_grid, c->ev_begin, 0)); forked_grid = true; }
        LIP_CUDA(c, cudaMemsetAsync(a.grid_start, 0, (size_t)c->tab_total * 4, s_grid));
        grid_build_kernel<<<c->M, 256, 0, s_grid>>>(a, (stages & ST_GRID_ALL) ? 1 : 0);
        ++launches;
    }
    mark();                                                  // 5: after grid build
    if (forked_icp) {
        LIP_CUDA(c, cudaEventRecord(c->ev_join[0], s_big)); LIP_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[0], 0));
    }
    if (forked_grid) { LIP_CUDA(c, cudaEventRecord(c->ev_join[2], s_grid)); LIP_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[2], 0)); }
    if ((stages & ST_BLEND) && c->n_tiles > 0) {
        Range r_("lip:blend+map");
        if (tiles) {
            bin_desc_kernel<<<(c->B + 127) / 128, 128, 0, st>>>(a);
            blend_kernel<<<(a.n_total + 255) / 256, 256, 0, st>>>(a);
            launches += 2;
            LIP_CHECK("blend_kernel");
            mark();                                          // 6: after the blend kernel
            if (c->map_mode == 2) {                          // scores on the tensor cores (tcgen05 + TMEM): two CTAs per SM share the 512 TMEM columns
                const size_t gm = std::min<size_t>(c->max_map_tiles, (size_t)c->sm_count * 2);
                blend_tile_tc_kernel<<<(unsigned)gm, LIP_TC_THREADS, 0, st>>>(a);
            } else {
                const size_t gm = std::min<size_t>(c->max_map_tiles, (size_t)c->sm_count * c->map_blocks_per_sm);
                blend_tile_kernel<<<(unsigned)gm, LIP_MAP_THREADS, 0, st>>>(a);
            }
        }
        else { mark(); blend_nn_kernel<<<c->n_tiles, LIP_TILE, 0, st>>>(a); }
        ++launches;
    } else mark();
    if ((stages & ST_QUERY) && c->n_tiles > 0) {
        nn_query_kernel<<<c->n_tiles, LIP_TILE, 0, st>>>(a);
        ++launches;
    }
    LIP_CHECK("final map kernel");
    mark();                                                  // 7: after the final-map search / query kernel
    if ((stages & ST_BLEND) && !tiles && c->n_tiles > 0) {
        int gr = std::max(1, std::min(c->sm_count * 8, (a.n_total + 127) / 128));
        retry_nn_kernel<<<gr, 128, 0, st>>>(a);
        ++launches;
    }
    mark();                                                  // 8: after retry
    if ((stages & (ST_QUERY | ST_BLEND)) && !tiles && c->n_tiles > 0) {
        int gf = std::max(1, std::min(c->sm_count * 8, (a.n_total + 3) / 4));
        far_nn_kernel<<<gf, 128, 0, st>>>(a);
        ++launches;
    }
    mark();                                                  // 9: after far queries
    LIP_CUDA(c, cudaGetLastError());
    c->last_launches = launches;
    c->ev_valid = c->profiling;
    return LIP_OK;
}

// One step of the planned batch on `st`.  A caller that repeats the SAME step — the sub-batches of the host pipeline do (their
// device buffers are their own), and so does a device-resident caller that reuses its buffers — gets it replayed as a CUDA
// graph from the third identical call on: the second one captures it (memsets, launches, the fork / join of the side streams).
// What a graph saves is host time, 47-93 -> 5-8 us per (sub-)batch (profiles/README.md): the pipelined host paths of
// medium-sized batches are bound by exactly that.  LIP_GRAPHS=0 turns it off; profiling, LIP_DEBUG_SYNC and a caller that is
// capturing `st` itself get plain launches.
int run_stages(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
               long long* T21, long long* T12, double* q_dist, int stages, cudaStream_t st, const int* mesh_row = nullptr) {
    if (!c->planned) return fail(c, LIP_ENOPLAN, "run: no batch planned");
    LipArgs& a = c->args;
    a.xyz1 = xyz1; a.xyz2 = xyz2; a.lm1 = lm1; a.lm2 = lm2; a.T21 = T21; a.T12 = T12; a.q_dist = q_dist;
    a.mesh_row = mesh_row;                                   // (reset below: later calls copy c->args)
    struct RowReset { LipArgs& a; ~RowReset() { a.mesh_row = nullptr; } } row_reset{a};
    set_param_args(c);
    static const bool graphs_env = !(getenv("LIP_GRAPHS") && atoi(getenv("LIP_GRAPHS")) == 0);
    static const bool debug_sync = getenv("LIP_DEBUG_SYNC") != nullptr;
    StepGraph& g = c->step_graph;
    bool may = graphs_env && c->graphs_ok && !c->profiling && !debug_sync;
    if (may) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); may = false; }
    }
    if (!may) return launch_stages(c, stages, st);
    const bool same = g.seen > 0 && g.plan_gen == c->plan_gen && g.stages == stages && g.st == st && !memcmp(&g.args, &a, sizeof(LipArgs));
    if (same && g.exec) {
        if (cudaGraphLaunch(g.exec, st) == cudaSuccess) {
            c->last_launches = g.launches; c->ev_valid = false; c->fill_done = false;
            return LIP_OK;
        }
        cudaGetLastError();
        c->graphs_ok = false;                                // never again on this context; this step goes out as plain launches
        return launch_stages(c, stages, st);
    }
    if (!same) {
        g.reset();
        memcpy(&g.args, &a, sizeof(LipArgs)); g.plan_gen = c->plan_gen; g.stages = stages; g.st = st; g.seen = 1;
        return launch_stages(c, stages, st);
    }
    ++g.seen;                                                // the second identical step: capture it, instantiate, launch
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError(); c->graphs_ok = false;
        return launch_stages(c, stages, st);
    }
    const int rc = launch_stages(c, stages, st);
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
    if (rc == LIP_OK && e_end == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        g.launches = c->last_launches;
        if (cudaGraphLaunch(g.exec, st) == cudaSuccess) return LIP_OK;
    } else if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    g.reset(); c->graphs_ok = false;
    if (rc != LIP_OK) return rc;
    return launch_stages(c, stages, st);                     // nothing ran during the capture
}

const int kFullPipeline = ST_PARTITION | ST_ICP | ST_GRID_ALL | ST_BLEND;   // (the grid is only built in map mode 1; the fill is resolved on the fly)

// the sub-batch set the last run used (host pipeline or device split)
std::vector<lip_ctx*>& split_kids(lip_ctx* c) { return c->last_run_dsplit ? c->dkids : c->kids; }
std::vector<int64_t>& split_seg(lip_ctx* c) { return c->last_run_dsplit ? c->dkid_seg : c->kid_seg; }
int split_n(lip_ctx* c) { return c->last_run_dsplit ? c->n_dkids : c->n_kids; }

// what lip_batch_match_host runs after the maps of a (sub-)batch: refinement (optional) and the score counts
struct MatchOpts { int refine; double voxel_size, sample_proportion; };
}  // namespace
extern "C" {
static int refine_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, long long* d_T21, const long long* d_T12, cudaStream_t st);
static int evaluate_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, const double* d_lm1, const double* d_lm2,
                           const long long* d_T21, const long long* d_T12, double voxel_size, double sample_proportion,
                           cudaStream_t st, long long* d_counts_out = nullptr, const int* mesh_row = nullptr);
}
namespace {

// H2D of one (sub-)batch, enqueued on `up` (the context's own stream, or the parent's upload stream: then the context's
// stream waits for it); no synchronisation
int host_upload_async(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2, cudaStream_t up = nullptr) {
    cudaStream_t st = up ? up : c->stream;
    const size_t b1 = (size_t)c->n1 * 24, b2 = (size_t)c->n2 * 24, bl = (size_t)c->sumK * 24;
    LIP_CUDA(c, c->in_xyz1.ensure(b1)); LIP_CUDA(c, c->in_xyz2.ensure(b2));
    LIP_CUDA(c, c->in_lm1.ensure(bl)); LIP_CUDA(c, c->in_lm2.ensure(bl));
    LIP_CUDA(c, c->out_T21.ensure((size_t)c->n2 * 8)); LIP_CUDA(c, c->out_T12.ensure((size_t)c->n1 * 8));
    static const bool trace = getenv("LIP_TRACE_PIPELINE") != nullptr;     // per-sub-batch timeline (lip_batch_run_host prints it)
    if (trace && !c->tr[0]) for (auto& e : c->tr) cudaEventCreate(&e);
    if (trace) cudaEventRecord(c->tr[0], st);
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz1.p, xyz1, b1, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz2.p, xyz2, b2, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm1.p, lm1, bl, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm2.p, lm2, bl, cudaMemcpyHostToDevice, st));
    if (trace) cudaEventRecord(c->tr[1], st);
    if (up) {
        LIP_CUDA(c, cudaEventRecord(c->ev_fork, up));        // (ev_fork is re-recorded later by run_stages, after this wait is enqueued)
        LIP_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
    }
    return LIP_OK;
}
// the whole pipeline on the uploaded (sub-)batch and D2H of the maps, enqueued on the context's stream
struct DevIn { const double *xyz1, *xyz2, *lm1, *lm2; const int* rows; };     // inputs that are not the context's own upload buffers (series)
int host_compute_async(lip_ctx* c, void* T21, void* T12, const MatchOpts* mo = nullptr, const DevIn* in = nullptr) {
    cudaStream_t st = c->stream;
    static const bool trace = getenv("LIP_TRACE_PIPELINE") != nullptr;
    const size_t msz = (size_t)c->map_bytes;
    if (in) { LIP_CUDA(c, c->out_T21.ensure((size_t)c->n2 * 8)); LIP_CUDA(c, c->out_T12.ensure((size_t)c->n1 * 8)); }
    int rc = in ? run_stages(c, in->xyz1, in->xyz2, in->lm1, in->lm2, c->out_T21.as<long long>(), c->out_T12.as<long long>(), nullptr,
                             kFullPipeline, st, in->rows)
                : run_stages(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->in_lm1.as<double>(), c->in_lm2.as<double>(),
                             c->out_T21.as<long long>(), c->out_T12.as<long long>(), nullptr, kFullPipeline, st);
    if (rc) return rc;
    if (mo) {                                                // the rest of the per-segment body, maps still on the device
        const int64_t hot = c->last_launches;
        if (mo->refine && (rc = refine_device(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->out_T21.as<long long>(),
                                              c->out_T12.as<long long>(), st)) != LIP_OK) return rc;
        if ((rc = evaluate_device(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->in_lm1.as<double>(), c->in_lm2.as<double>(),
                                  c->out_T21.as<long long>(), c->out_T12.as<long long>(), mo->voxel_size, mo->sample_proportion,
                                  st)) != LIP_OK) return rc;
        c->last_launches = hot + (mo->refine ? 2 : 0) + 2;
    }
    if (trace) cudaEventRecord(c->tr[2], st);
    if ((c->directions & 1) && T21) LIP_CUDA(c, cudaMemcpyAsync(T21, c->out_T21.p, (size_t)c->n2 * msz, cudaMemcpyDeviceToHost, st));
    if ((c->directions & 2) && T12) LIP_CUDA(c, cudaMemcpyAsync(T12, c->out_T12.p, (size_t)c->n1 * msz, cudaMemcpyDeviceToHost, st));
    if (trace) cudaEventRecord(c->tr[3], st);
    return LIP_OK;
}

// ---- pageable caller arrays: the pipeline fed through pinned staging buffers --------------------------------------
bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}
int staging_threads() {
    static const int n = [] {
        if (const char* e = getenv("LIP_STAGING_THREADS")) return std::max(0, atoi(e));
        const unsigned hw = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(6u, hw / 2));
    }();
    return n;
}
// lip_set_input_validation: the check sklearn's check_array makes at every KDTree construction of the reference
// (registration_icp.py:51), here once per call and on all copy threads
int check_finite(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2) {
    const size_t slice = (size_t)1 << 20;
    const size_t b[4] = {(size_t)c->n1 * 24, (size_t)c->n2 * 24, (size_t)c->sumK * 24, (size_t)c->sumK * 24};
    const double* src[4] = {xyz1, xyz2, lm1, lm2};
    std::atomic<int> bad(0), left(0);
    if (b[0] + b[1] < ((size_t)1 << 21) || staging_threads() == 0) {
        for (int i = 0; i < 4; ++i) if (StagePool::nonfinite(src[i], b[i])) bad.store(1);
    } else {
        if (!c->pool) c->pool = new StagePool();
        c->pool->start(staging_threads());
        int n = 0;
        for (int i = 0; i < 4; ++i) n += StagePool::slices_of(b[i], slice);
        left.store(n);
        for (int i = 0; i < 4; ++i) c->pool->submit(nullptr, src[i], b[i], slice, &left, &bad);
        c->pool->wait(left);
    }
    return bad.load() ? fail(c, LIP_EINVAL, "Input contains NaN or infinity.") : LIP_OK;
}

// Same pipeline as lip_batch_run_host's, with a host stage in front and behind: copy threads fill the pinned input buffer
// of sub-batch k (in order), the calling thread enqueues its upload as soon as it is full and the kernels + download of the
// sub-batches already uploaded in between; finished maps go from the pinned output buffers to the caller's arrays.
int run_host_staged(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                    void* T21v, void* T12v, const MatchOpts* mo) {
    const int nk = c->n_kids;
    const size_t msz = (size_t)c->map_bytes;
    char *T21 = (char*)T21v, *T12 = (char*)T12v;
    if (!c->pool) c->pool = new StagePool();
    c->pool->start(staging_threads());
    StagePool& pool = *c->pool;
    const size_t slice = (size_t)1 << 20;
    std::vector<std::atomic<int>> in_left(nk), out_left(nk);
    std::atomic<int> bad(0);
    std::atomic<int>* chk = c->validate_inputs ? &bad : nullptr;
    struct Part { size_t b1, b2, bl; char* in; char *o21, *o12; };
    std::vector<Part> part(nk);
    for (int k = 0; k < nk; ++k) {
        lip_ctx* kc = c->kids[k];
        Part& p = part[k];
        p.b1 = (size_t)kc->n1 * 24; p.b2 = (size_t)kc->n2 * 24; p.bl = (size_t)kc->sumK * 24;
        LIP_CUDA(c, kc->st_in.ensure(p.b1 + p.b2 + 2 * p.bl));
        LIP_CUDA(c, kc->st_out.ensure(((size_t)kc->n1 + (size_t)kc->n2) * 8));
        p.in = (char*)kc->st_in.p;
        p.o21 = (char*)kc->st_out.p; p.o12 = p.o21 + (size_t)kc->n2 * msz;
        const int64_t s0 = c->kid_seg[k];
        in_left[k].store(StagePool::slices_of(p.b1, slice) + StagePool::slices_of(p.b2, slice) + 2 * StagePool::slices_of(p.bl, slice));
        pool.submit(p.in, xyz1 + 3 * c->h_off1[s0], p.b1, slice, &in_left[k], chk);
        pool.submit(p.in + p.b1, xyz2 + 3 * c->h_off2[s0], p.b2, slice, &in_left[k], chk);
        pool.submit(p.in + p.b1 + p.b2, lm1 + 3 * c->h_lm_off[s0], p.bl, slice, &in_left[k], chk);
        pool.submit(p.in + p.b1 + p.b2 + p.bl, lm2 + 3 * c->h_lm_off[s0], p.bl, slice, &in_left[k], chk);
    }
    int rc = LIP_OK, up = 0, cp = 0;
    while (cp < nk && rc == LIP_OK) {
        if (up < nk && (pool.done(in_left[up]) || cp == up)) {
            pool.wait(in_left[up]);
            if (bad.load()) { rc = fail(c, LIP_EINVAL, "Input contains NaN or infinity."); break; }   // nothing of it was uploaded
            lip_ctx* kc = c->kids[up];
            const Part& p = part[up];
            kc->params = c->params; kc->map_bytes = c->map_bytes;
            rc = host_upload_async(kc, (const double*)p.in, (const double*)(p.in + p.b1), (const double*)(p.in + p.b1 + p.b2),
                                   (const double*)(p.in + p.b1 + p.b2 + p.bl), c->side[1]);
            if (rc) c->err = "pipeline upload: " + kc->err;
            ++up;
        } else {
            lip_ctx* kc = c->kids[cp];
            rc = host_compute_async(kc, T21 ? part[cp].o21 : nullptr, T12 ? part[cp].o12 : nullptr, mo);
            if (rc) c->err = "pipeline run: " + kc->err;
            ++cp;
        }
    }
    for (int k = 0; k < nk; ++k) {
        cudaError_t e = cudaStreamSynchronize(c->kids[k]->stream);
        if (e != cudaSuccess && rc == LIP_OK) rc = fail(c, LIP_ECUDA, std::string("pipeline sync: ") + cudaGetErrorString(e));
        out_left[k].store(0);
        if (rc != LIP_OK || k >= cp) continue;
        const int64_t s0 = c->kid_seg[k];
        const size_t o1 = T12 ? (size_t)c->kids[k]->n1 * msz : 0, o2 = T21 ? (size_t)c->kids[k]->n2 * msz : 0;
        out_left[k].store(StagePool::slices_of(o1, slice) + StagePool::slices_of(o2, slice));
        if (o2) pool.submit(T21 + (size_t)c->h_off2[s0] * msz, part[k].o21, o2, slice, &out_left[k]);
        if (o1) pool.submit(T12 + (size_t)c->h_off1[s0] * msz, part[k].o12, o1, slice, &out_left[k]);
    }
    for (int k = 0; k < nk; ++k) { pool.wait(in_left[k]); pool.wait(out_left[k]); }   // nothing of this call left in the queue
    if (rc != LIP_OK) return rc;
    c->last_run_split = true; c->last_run_dsplit = false;
    return LIP_OK;
}

// The planned batch as a pipeline of its sub-batches (n_kids >= 2), every sub-batch on its own stream: copies in, kernels
// and copies out of different sub-batches overlap.  mo != nullptr: the whole per-segment body (lip_batch_match_host), the
// score counts of sub-batch k are left in kids[k]->h_counts.
int run_host_split(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                   void* T21v, void* T12v, const MatchOpts* mo) {
    const size_t msz = (size_t)c->map_bytes;
    char *T21 = (char*)T21v, *T12 = (char*)T12v;
    if (staging_threads() > 0 &&
        (is_pageable(xyz1) || is_pageable(xyz2) || (T21 && is_pageable(T21)) || (T12 && is_pageable(T12)))) {
        bool ready = true;                                   // pinned staging buffers for every sub-batch; if the host cannot pin
        for (int k = 0; k < c->n_kids && ready; ++k) {       // that much memory the arrays go to the driver as they are
            lip_ctx* kc = c->kids[k];
            ready = kc->st_in.ensure(((size_t)kc->n1 + (size_t)kc->n2 + 2 * (size_t)kc->sumK) * 24) == cudaSuccess &&
                    kc->st_out.ensure(((size_t)kc->n1 + (size_t)kc->n2) * 8) == cudaSuccess;
        }
        if (ready) return run_host_staged(c, xyz1, xyz2, lm1, lm2, T21, T12, mo);
        cudaGetLastError();
    }
    if (c->validate_inputs) {
        int rc = check_finite(c, xyz1, xyz2, lm1, lm2);
        if (rc) return rc;
    }
    // nothing enqueued by this call may still touch the caller's arrays when an error is returned
    auto drain = [&] { cudaStreamSynchronize(c->side[1]); for (int k = 0; k < c->n_kids; ++k) cudaStreamSynchronize(c->kids[k]->stream); };
    // All uploads are enqueued first, on ONE stream: issuing the kernels of a sub-batch takes the host longer than its
    // upload takes the copy engine (which must never wait for the host), and in-order uploads give every sub-batch its
    // data as early as possible.
    for (int k = 0; k < c->n_kids; ++k) {
        lip_ctx* kc = c->kids[k];
        const int64_t s0 = c->kid_seg[k];
        kc->params = c->params; kc->map_bytes = c->map_bytes;
        int rc = host_upload_async(kc, xyz1 + 3 * c->h_off1[s0], xyz2 + 3 * c->h_off2[s0], lm1 + 3 * c->h_lm_off[s0], lm2 + 3 * c->h_lm_off[s0],
                                   c->side[1]);             // one upload stream: the sub-batches arrive in order, each as early as possible
        if (rc) { drain(); return fail(c, rc, "pipeline upload: " + kc->err); }
    }
    for (int k = 0; k < c->n_kids; ++k) {
        lip_ctx* kc = c->kids[k];
        const int64_t s0 = c->kid_seg[k];
        int rc = host_compute_async(kc, T21 ? T21 + (size_t)c->h_off2[s0] * msz : nullptr, T12 ? T12 + (size_t)c->h_off1[s0] * msz : nullptr, mo);
        if (rc) { drain(); return fail(c, rc, "pipeline run: " + kc->err); }
    }
    for (int k = 0; k < c->n_kids; ++k) LIP_CUDA(c, cudaStreamSynchronize(c->kids[k]->stream));
    if (getenv("LIP_TRACE_PIPELINE") && c->kids[0]->tr[0]) {
        for (int k = 0; k < c->n_kids; ++k) {
            float t[4];
            for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], c->kids[0]->tr[0], c->kids[k]->tr[i]);
            fprintf(stderr, "[lip pipeline] sub-batch %d: H2D %.3f-%.3f ms, kernels end %.3f, D2H end %.3f\n", k, t[0], t[1], t[2], t[3]);
        }
    }
    c->last_run_split = true; c->last_run_dsplit = false;
    return LIP_OK;
}

// The staged transfers of ONE call of a *_host entry point that is not pipelined over sub-batches (evaluate, refine, match,
// interpolate): large pageable arrays go through the context's pinned buffers in 8 MB chunks — the copy threads fill chunk
// i + 1 while chunk i is on the bus; downloads land in the pinned output buffer chunk by chunk (an event each) and are
// copied out by the threads in finish().  Small or pinned arrays take plain cudaMemcpyAsync.
struct StageCall {
    lip_ctx* c;
    size_t in_off = 0, out_off = 0;
    std::deque<std::atomic<int>> counters;                   // stable addresses
    struct Out { void* h_dst; const void* pinned; size_t bytes; cudaEvent_t ev; };
    std::vector<Out> outs;
    size_t n_ev = 0;
    static constexpr size_t kChunk = (size_t)8 << 20, kSlice = (size_t)1 << 20;
    explicit StageCall(lip_ctx* ctx) : c(ctx) {}
    static bool wanted(const void* h, size_t bytes) { return bytes >= ((size_t)1 << 20) && staging_threads() > 0 && is_pageable(h); }
    // capacity for everything this call will stage (the buffers must not move once in use)
    int begin(size_t in_bytes, size_t out_bytes) {
        // (if the host cannot pin that much memory the transfers below fall back to plain copies: capacity stays short)
        if (in_bytes && c->st_in.ensure(in_bytes + 4096) != cudaSuccess) cudaGetLastError();
        if (out_bytes && c->st_out.ensure(out_bytes + 4096) != cudaSuccess) cudaGetLastError();
        if ((in_bytes || out_bytes) && !c->pool) c->pool = new StagePool();
        if (c->pool) c->pool->start(staging_threads());
        return LIP_OK;
    }
    int h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
        if (!bytes) return LIP_OK;
        if (!wanted(h_src, bytes) || in_off + bytes > c->st_in.cap) {
            LIP_CUDA(c, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
            return LIP_OK;
        }
        char* pin = (char*)c->st_in.p + in_off;
        in_off += (bytes + 255) & ~(size_t)255;
        const size_t first = counters.size();
        for (size_t o = 0; o < bytes; o += kChunk) {
            const size_t b = std::min(kChunk, bytes - o);
            counters.emplace_back(StagePool::slices_of(b, kSlice));
            c->pool->submit(pin + o, (const char*)h_src + o, b, kSlice, &counters.back());
        }
        size_t k = first;
        for (size_t o = 0; o < bytes; o += kChunk, ++k) {
            c->pool->wait(counters[k]);
            LIP_CUDA(c, cudaMemcpyAsync((char*)d_dst + o, pin + o, std::min(kChunk, bytes - o), cudaMemcpyHostToDevice, st));
        }
        return LIP_OK;
    }
    int d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
        if (!bytes) return LIP_OK;
        if (!wanted(h_dst, bytes) || out_off + bytes > c->st_out.cap) {
            LIP_CUDA(c, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
            return LIP_OK;
        }
        char* pin = (char*)c->st_out.p + out_off;
        out_off += (bytes + 255) & ~(size_t)255;
        for (size_t o = 0; o < bytes; o += kChunk) {
            const size_t b = std::min(kChunk, bytes - o);
            if (n_ev == c->stage_events.size()) {
                cudaEvent_t e;
                LIP_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                c->stage_events.push_back(e);
            }
            LIP_CUDA(c, cudaMemcpyAsync(pin + o, (const char*)d_src + o, b, cudaMemcpyDeviceToHost, st));
            LIP_CUDA(c, cudaEventRecord(c->stage_events[n_ev], st));
            outs.push_back({(char*)h_dst + o, pin + o, b, c->stage_events[n_ev++]});
        }
        return LIP_OK;
    }
    // everything enqueued on st is complete and every staged download has reached the caller's arrays on return
    int finish(cudaStream_t st) {
        cudaError_t err = cudaSuccess;
        for (const Out& o : outs) {
            cudaError_t e = cudaEventSynchronize(o.ev);
            if (e != cudaSuccess) { err = e; break; }
            counters.emplace_back(StagePool::slices_of(o.bytes, kSlice));
            c->pool->submit(o.h_dst, o.pinned, o.bytes, kSlice, &counters.back());
        }
        for (auto& k : counters) c->pool->wait(k);
        if (err == cudaSuccess) err = cudaStreamSynchronize(st);
        if (err != cudaSuccess) return fail(c, LIP_ECUDA, std::string("staged transfer: ") + cudaGetErrorString(err));
        return LIP_OK;
    }
    ~StageCall() { if (c->pool) for (auto& k : counters) c->pool->wait(k); }     // no job may outlive the counters
};

}  // namespace

extern "C" {

int lip_version(void) { return LIP_ABI_VERSION; }
int lip_compiled_arch(void) { return LIP_ARCH; }

void lip_default_params(lip_params* p) {
    if (!p) return;
    p->max_iter = 500;
    p->cost_rise_tol = 1e-4;
    p->cost_flat_tol = 0.01;
    p->obtuse_eps = 1e-5;
    p->cos_den_eps = 1e-5;
    p->reject_dist = std::numeric_limits<double>::infinity();
}

const char* lip_last_error(const lip_ctx* ctx) { return ctx ? ctx->err.c_str() : g_startup_error.c_str(); }

int lip_create(lip_ctx** out, int device) {
    if (!out) return LIP_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, LIP_ENODEV, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
    if (device < 0 || device >= count) return fail(nullptr, LIP_EINVAL, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, LIP_ENODEV, cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, LIP_ENODEV, cudaGetErrorString(e));
    if (prop.major * 10 + prop.minor != LIP_ARCH)
        return fail(nullptr, LIP_ENODEV, "liblocalicp was built for sm_" + std::to_string(LIP_ARCH) + "a only; device is sm_" +
                                             std::to_string(prop.major * 10 + prop.minor));
    lip_ctx* c = new lip_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    lip_default_params(&c->params);
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fail(nullptr, LIP_ENODEV, cudaGetErrorString(e));
    }
    for (int i = 0; i < 4; ++i) {
        cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&c->ev_begin, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    int nb = 0;
    cudaFuncSetAttribute(icp_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGroupSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, icp_group_kernel, LIP_ICP_WARPS * 32, kGroupSmem) == cudaSuccess && nb > 0)
        c->group_blocks_per_sm = nb;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, blend_tile_kernel, LIP_MAP_THREADS, 0) == cudaSuccess && nb > 0)
        c->map_blocks_per_sm = nb;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, icp_big_kernel, LIP_BIG_THREADS, 0) == cudaSuccess && nb > 0)
        c->big_blocks_per_sm = nb;
    // thread-block clusters for the heaviest partitions: size and first work class (LIP_CLUSTER_SIZE=1 or LIP_CLUSTER_CLASS=4: off)
    if (const char* e = getenv("LIP_CLUSTER_SIZE")) c->cluster_size = std::max(1, std::min(8, atoi(e)));
    if (const char* e = getenv("LIP_CLUSTER_CLASS")) c->cluster_cls = std::max(0, std::min(4, atoi(e)));
    if (c->cluster_size < 2) c->cluster_cls = 4;
    if (const char* e = getenv("LIP_CLUSTER_COOP_PARTS")) c->cluster_max_bins = std::max<int64_t>(0, atoll(e));
    if (c->cluster_cls < 4) {
        cudaLaunchConfig_t q = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)c->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.gridDim = dim3((unsigned)c->cluster_size); q.blockDim = dim3(LIP_CLUSTER_THREADS); q.attrs = at; q.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, icp_cluster_kernel<LIP_CLUSTER_THREADS, LIP_CLUSTER_CHUNK>, &q) == cudaSuccess && ncl > 0)
            c->cluster_grid = ncl * c->cluster_size;
        else { cudaGetLastError(); c->cluster_cls = 4; }       // no room for such clusters on this device: everything stays with icp_big_kernel
    }
    cudaFuncSetAttribute(landmark_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LIP_MAX_LANDMARKS * 44 + 32);
    cudaFuncSetAttribute(partition_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         LIP_MAX_LANDMARKS * 4 * (2 + LIP_PART_THREADS / 32) + 16);
    memset(&c->args, 0, sizeof(c->args));
    *out = c;
    return LIP_OK;
}

void lip_destroy(lip_ctx* c) {
    if (!c) return;
    for (lip_ctx* k : c->kids) lip_destroy(k);
    c->kids.clear();
    for (lip_ctx* k : c->dkids) lip_destroy(k);
    c->dkids.clear();
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->step_graph.reset();
    for (int i = 0; i <= LIP_N_STAGES; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 4; ++i) { if (c->side[i]) cudaStreamDestroy(c->side[i]); if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]); }
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    c->in_frames.release(); c->d_rows.release();
    if (c->ev_piece) cudaEventDestroy(c->ev_piece);
    for (DevBuf* b : {&c->d_plan, &c->ws, &c->in_xyz1, &c->in_xyz2, &c->in_lm1, &c->in_lm2, &c->out_T21, &c->out_T12, &c->misc, &c->misc2})
        b->release();
    delete c->pool;
    c->st_in.release(); c->st_out.release(); c->h_counts.release(); c->d_err.release(); c->h_err.release();
    if (c->ev_chk) cudaEventDestroy(c->ev_chk);
    for (cudaEvent_t e : c->stage_events) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

int lip_set_params(lip_ctx* c, const lip_params* p) {
    if (!c || !p) return LIP_EINVAL;
    if (p->max_iter < 1) return fail(c, LIP_EINVAL, "max_iter must be >= 1");
    c->params = *p;
    return LIP_OK;
}

int lip_set_input_validation(lip_ctx* c, int on) {
    if (!c) return LIP_EINVAL;
    c->validate_inputs = on != 0;
    return LIP_OK;
}

int lip_set_map_mode(lip_ctx* c, int mode) {
    if (!c) return LIP_EINVAL;
    if (mode < 0 || mode > 2) return fail(c, LIP_EINVAL, "map mode must be 0 (landmark-bin tiles), 1 (hashed grid) or 2 (tiles on the tensor cores)");
    if (mode != c->map_mode) {                               // the tile list depends on the mode: plan again (sub-batches too)
        c->planned = false; c->user_plan = false; c->n_kids = 0; c->n_dkids = 0;
    }
    c->map_mode = mode;
    return LIP_OK;
}

int lip_set_map_dtype(lip_ctx* c, int bits) {
    if (!c) return LIP_EINVAL;
    if (bits != 16 && bits != 32 && bits != 64) return fail(c, LIP_EINVAL, "map dtype must be 16, 32 or 64 bits");
    if (bits == 16 && c->planned && c->user_plan && c->max_mesh > 65535)
        return fail(c, LIP_EINVAL, "16-bit maps need every mesh of the planned batch to have fewer than 65,536 vertices");
    c->map_bytes = bits / 8;
    return LIP_OK;
}

int lip_host_alloc(void** p, uint64_t bytes) {
    if (!p) return LIP_EINVAL;
    return cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? LIP_OK : LIP_ECUDA;
}
int lip_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? LIP_OK : LIP_ECUDA; }

// split the planned batch into `chunks` sub-batches of about equal vertex count, each planned on its own child context
struct KidSet { std::vector<lip_ctx*>* kids; std::vector<int64_t>* seg; int* n; };
static int plan_kids(lip_ctx* c, KidSet ks, int chunks, int64_t S, const int64_t* off1, const int64_t* off2, const int64_t* lm_off,
                     int directions) {
    *ks.n = 0;
    const int64_t n = off1[S] + off2[S];
    double ramp = LIP_PIPELINE_RAMP;
    if (const char* e = getenv("LIP_PIPELINE_RAMP")) ramp = std::max(1.0, std::min(4.0, atof(e)));
    int want = (int)std::min<int64_t>(chunks, S);
    if (n < LIP_PIPELINE_MIN_VERTICES || want < 2) return LIP_OK;
    std::vector<int64_t>& seg = *ks.seg;
    seg.assign(1, 0);
    for (int k = 1; k < want; ++k) {                         // cut where the running vertex count passes (k/want)^ramp of the total
        const int64_t target = (int64_t)((double)n * std::pow((double)k / want, ramp));
        int64_t s = seg.back() + 1;
        while (s < S && off1[s] + off2[s] < target) ++s;
        if (s >= S) break;
        seg.push_back(s);
    }
    seg.push_back(S);
    const int nk = (int)seg.size() - 1;
    if (nk < 2) return LIP_OK;
    std::vector<lip_ctx*>& kids = *ks.kids;
    while ((int)kids.size() < nk) {
        lip_ctx* k = nullptr;
        int rc = lip_create(&k, c->device);
        if (rc) return fail(c, rc, std::string("pipeline context: ") + lip_last_error(nullptr));
        k->is_kid = true;
        kids.push_back(k);
    }
    std::vector<int64_t> o1, o2, ol;
    for (int k = 0; k < nk; ++k) {
        const int64_t s0 = seg[k], s1 = seg[k + 1];
        o1.assign(off1 + s0, off1 + s1 + 1); o2.assign(off2 + s0, off2 + s1 + 1); ol.assign(lm_off + s0, lm_off + s1 + 1);
        for (auto& x : o1) x -= off1[s0];
        for (auto& x : o2) x -= off2[s0];
        for (auto& x : ol) x -= lm_off[s0];
        lip_ctx* kc = kids[k];
        kc->params = c->params; kc->map_mode = c->map_mode;
        int rc = plan_impl(kc, s1 - s0, o1.data(), o2.data(), ol.data(), directions);
        if (rc) return fail(c, rc, "pipeline plan: " + kc->err);
        kc->h_off1 = o1; kc->h_off2 = o2; kc->h_lm_off = ol;   // the refinement and the scores of a sub-batch slice by segment pair
        kc->user_plan = true;
    }
    *ks.n = nk;
    return LIP_OK;
}

int lip_batch_plan(lip_ctx* c, int64_t S, const int64_t* off1, const int64_t* off2, const int64_t* lm_off, int directions) {
    Range nvtx_("lip_batch_plan");
    if (!c) return LIP_EINVAL;
    cudaSetDevice(c->device);
    int chunks = LIP_PIPELINE_CHUNKS;
    if (const char* e = getenv("LIP_PIPELINE_CHUNKS")) chunks = std::max(1, std::min(64, atoi(e)));   // tuning knob
    int dchunks = LIP_DEVICE_CHUNKS;
    if (const char* e = getenv("LIP_DEVICE_CHUNKS")) dchunks = std::max(1, std::min(16, atoi(e)));
    // the same batch shape as the one already planned (a caller registering frame after frame of one layout): nothing to do
    if (c->planned && c->user_plan && S > 0 && off1 && off2 && lm_off && c->S == S && c->directions == directions && (int64_t)c->h_off1.size() == S + 1 &&
        c->plan_chunks == chunks && c->plan_dchunks == dchunks &&
        !memcmp(c->h_off1.data(), off1, (S + 1) * 8) && !memcmp(c->h_off2.data(), off2, (S + 1) * 8) &&
        !memcmp(c->h_lm_off.data(), lm_off, (S + 1) * 8))
        return LIP_OK;
    c->user_plan = false; c->n_kids = 0; c->n_dkids = 0;
    int rc = plan_impl(c, S, off1, off2, lm_off, directions);
    if (rc) return rc;
    c->h_off1.assign(off1, off1 + S + 1); c->h_off2.assign(off2, off2 + S + 1); c->h_lm_off.assign(lm_off, lm_off + S + 1);
    c->last_run_split = false;
    c->plan_chunks = chunks; c->plan_dchunks = dchunks;
    rc = plan_kids(c, KidSet{&c->kids, &c->kid_seg, &c->n_kids}, chunks, S, off1, off2, lm_off, directions);
    if (rc) { c->n_kids = 0; c->planned = false; return rc; }
    // the device-resident entry point runs the batch as LIP_DEVICE_CHUNKS sub-batches on their own streams: the
    // latency-bound kernels of one (partition, the tails of the ICP kernels) overlap the scans of the other
    rc = plan_kids(c, KidSet{&c->dkids, &c->dkid_seg, &c->n_dkids}, dchunks, S, off1, off2, lm_off, directions);
    if (rc) { c->n_kids = 0; c->n_dkids = 0; c->planned = false; return rc; }
    c->user_plan = true;
    return LIP_OK;
}

int lip_batch_run_device(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                         void* T21, void* T12, void* stream) {
    Range nvtx_("lip_batch_run_device");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan) return fail(c, LIP_ENOPLAN, "run: no batch planned");
    if (!xyz1 || !xyz2 || !lm1 || !lm2) return fail(c, LIP_EINVAL, "run: null input pointer");
    if (((c->directions & 1) && !T21) || ((c->directions & 2) && !T12)) return fail(c, LIP_EINVAL, "run: null output pointer");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (c->n_dkids >= 2 && !c->profiling) {
        // sub-batches on the child contexts' streams, fork / join on the caller's stream
        LIP_CUDA(c, cudaEventRecord(c->ev_fork, st));
        for (int k = 0; k < c->n_dkids; ++k) {
            lip_ctx* kc = c->dkids[k];
            const int64_t s0 = c->dkid_seg[k];
            kc->params = c->params; kc->map_bytes = c->map_bytes;
            LIP_CUDA(c, cudaStreamWaitEvent(kc->stream, c->ev_fork, 0));
            const size_t msz = (size_t)c->map_bytes;
            int rc = run_stages(kc, xyz1 + 3 * c->h_off1[s0], xyz2 + 3 * c->h_off2[s0], lm1 + 3 * c->h_lm_off[s0], lm2 + 3 * c->h_lm_off[s0],
                                T21 ? (long long*)((char*)T21 + (size_t)c->h_off2[s0] * msz) : nullptr,
                                T12 ? (long long*)((char*)T12 + (size_t)c->h_off1[s0] * msz) : nullptr,
                                nullptr, kFullPipeline, kc->stream);
            if (rc) return fail(c, rc, "device split: " + kc->err);
            LIP_CUDA(c, cudaEventRecord(kc->ev_begin, kc->stream));      // (ev_begin is free again once run_stages has forked)
            LIP_CUDA(c, cudaStreamWaitEvent(st, kc->ev_begin, 0));
        }
        c->last_run_split = true; c->last_run_dsplit = true;
        return LIP_OK;
    }
    c->last_run_split = false; c->last_run_dsplit = false;
    return run_stages(c, xyz1, xyz2, lm1, lm2, (long long*)T21, (long long*)T12, nullptr, kFullPipeline, st);
}

int lip_batch_run_host(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                       void* T21, void* T12) {
    Range nvtx_("lip_batch_run_host");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan) return fail(c, LIP_ENOPLAN, "run: no batch planned");
    if (!xyz1 || !xyz2 || !lm1 || !lm2) return fail(c, LIP_EINVAL, "run: null input pointer");
    if (((c->directions & 1) && !T21) || ((c->directions & 2) && !T12)) return fail(c, LIP_EINVAL, "run: null output pointer");
    cudaSetDevice(c->device);
    if (c->n_kids >= 2) return run_host_split(c, xyz1, xyz2, lm1, lm2, T21, T12, nullptr);
    if (c->validate_inputs) {
        int rc = check_finite(c, xyz1, xyz2, lm1, lm2);
        if (rc) return rc;
    }
    c->last_run_split = false; c->last_run_dsplit = false;
    int rc = host_upload_async(c, xyz1, xyz2, lm1, lm2);
    if (rc == LIP_OK) rc = host_compute_async(c, T21, T12);
    if (rc) return rc;
    LIP_CUDA(c, cudaStreamSynchronize(c->stream));
    return LIP_OK;
}

// Time series (run_registration_pipeline.py:64,100-104: frame k is day2 of frame pair k-1 and day1 of frame pair k).
// `frames` holds every mesh ONCE; row1[s] / row2[s] = first row of mesh1 / mesh2 of segment pair s in it, so meshes can be shared
// between segment pairs.  Rows of this context's meshes as the kernels want them ([mesh1 of every pair | mesh2 of every pair]):
static int series_rows(lip_ctx* kc, const int64_t* row1, const int64_t* row2, int64_t s0, int64_t n_rows, const int64_t* off1,
                       const int64_t* off2, cudaStream_t st) {
    const int64_t S = kc->S;
    bool same = kc->rows_on_device && kc->h_rows.size() == 2 * (size_t)S;      // a series step that repeats: the rows are already there
    kc->h_rows.resize(2 * (size_t)S);
    for (int64_t s = 0; s < S; ++s) {
        const int64_t r1 = row1[s0 + s], r2 = row2[s0 + s];
        if (r1 < 0 || r2 < 0 || r1 + (off1[s0 + s + 1] - off1[s0 + s]) > n_rows || r2 + (off2[s0 + s + 1] - off2[s0 + s]) > n_rows) {
            kc->rows_on_device = false;
            return fail(kc, LIP_EINVAL, "series: a mesh of segment pair " + std::to_string(s0 + s) + " does not lie inside `frames`");
        }
        same = same && kc->h_rows[s] == (int)r1 && kc->h_rows[S + s] == (int)r2;
        kc->h_rows[s] = (int)r1; kc->h_rows[S + s] = (int)r2;
    }
    const void* before = kc->d_rows.p;
    LIP_CUDA(kc, kc->d_rows.ensure(2 * (size_t)S * 4));
    if (same && before == kc->d_rows.p && st == kc->rows_stream) return LIP_OK;      // (same stream: ordered after that upload)
    kc->rows_on_device = false;
    LIP_CUDA(kc, cudaMemcpyAsync(kc->d_rows.p, kc->h_rows.data(), 2 * (size_t)S * 4, cudaMemcpyHostToDevice, st));
    kc->rows_on_device = true; kc->rows_stream = st;
    return LIP_OK;
}

int lip_series_run_device(lip_ctx* c, const double* d_frames, int64_t n_rows, const int64_t* row1, const int64_t* row2,
                          const double* d_lm1, const double* d_lm2, void* T21, void* T12, void* stream) {
    Range nvtx_("lip_series_run_device");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan) return fail(c, LIP_ENOPLAN, "series: no batch planned");
    if (!d_frames || !row1 || !row2 || !d_lm1 || !d_lm2) return fail(c, LIP_EINVAL, "series: null input pointer");
    if (((c->directions & 1) && !T21) || ((c->directions & 2) && !T12)) return fail(c, LIP_EINVAL, "series: null output pointer");
    if (n_rows <= 0 || n_rows >= ((int64_t)1 << 31)) return fail(c, LIP_EINVAL, "series: n_rows out of range");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t msz = (size_t)c->map_bytes;
    if (c->n_dkids >= 2 && !c->profiling) {
        LIP_CUDA(c, cudaEventRecord(c->ev_fork, st));
        for (int k = 0; k < c->n_dkids; ++k) {
            lip_ctx* kc = c->dkids[k];
            const int64_t s0 = c->dkid_seg[k];
            kc->params = c->params; kc->map_bytes = c->map_bytes;
            LIP_CUDA(c, cudaStreamWaitEvent(kc->stream, c->ev_fork, 0));
            int rc = series_rows(kc, row1, row2, s0, n_rows, c->h_off1.data(), c->h_off2.data(), kc->stream);
            if (rc == LIP_OK)
                rc = run_stages(kc, d_frames, d_frames, d_lm1 + 3 * c->h_lm_off[s0], d_lm2 + 3 * c->h_lm_off[s0],
                                T21 ? (long long*)((char*)T21 + (size_t)c->h_off2[s0] * msz) : nullptr,
                                T12 ? (long long*)((char*)T12 + (size_t)c->h_off1[s0] * msz) : nullptr,
                                nullptr, kFullPipeline, kc->stream, kc->d_rows.as<int>());
            if (rc) return fail(c, rc, "series device split: " + kc->err);
            LIP_CUDA(c, cudaEventRecord(kc->ev_begin, kc->stream));
            LIP_CUDA(c, cudaStreamWaitEvent(st, kc->ev_begin, 0));
        }
        c->last_run_split = true; c->last_run_dsplit = true;
        return LIP_OK;
    }
    c->last_run_split = false; c->last_run_dsplit = false;
    int rc = series_rows(c, row1, row2, 0, n_rows, c->h_off1.data(), c->h_off2.data(), st);
    if (rc) return rc;
    return run_stages(c, d_frames, d_frames, d_lm1, d_lm2, (long long*)T21, (long long*)T12, nullptr, kFullPipeline, st, c->d_rows.as<int>());
}

// One upload of `frames` (in pieces, so that the sub-batches of the pipeline start as soon as their rows have arrived) instead
// of the two uploads of lip_batch_run_host.
int lip_series_run_host(lip_ctx* c, const double* frames, int64_t n_rows, const int64_t* row1, const int64_t* row2,
                        const double* lm1, const double* lm2, void* T21v, void* T12v) {
    Range nvtx_("lip_series_run_host");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan) return fail(c, LIP_ENOPLAN, "series: no batch planned");
    if (!frames || !row1 || !row2 || !lm1 || !lm2) return fail(c, LIP_EINVAL, "series: null input pointer");
    if (((c->directions & 1) && !T21v) || ((c->directions & 2) && !T12v)) return fail(c, LIP_EINVAL, "series: null output pointer");
    if (n_rows <= 0 || n_rows >= ((int64_t)1 << 31)) return fail(c, LIP_EINVAL, "series: n_rows out of range");
    cudaSetDevice(c->device);
    const size_t msz = (size_t)c->map_bytes;
    char *T21 = (char*)T21v, *T12 = (char*)T12v;
    const int nk = c->n_kids >= 2 ? c->n_kids : 1;
    auto kid = [&](int k) { return c->n_kids >= 2 ? c->kids[k] : c; };
    auto seg0 = [&](int k) { return c->n_kids >= 2 ? c->kid_seg[k] : (int64_t)0; };
    auto seg1 = [&](int k) { return c->n_kids >= 2 ? c->kid_seg[k + 1] : c->S; };
    if (!c->ev_piece) LIP_CUDA(c, cudaEventCreateWithFlags(&c->ev_piece, cudaEventDisableTiming));
    const size_t bl = (size_t)c->sumK * 24;
    LIP_CUDA(c, c->in_frames.ensure((size_t)n_rows * 24));
    LIP_CUDA(c, c->in_lm1.ensure(bl)); LIP_CUDA(c, c->in_lm2.ensure(bl));
    // rows sub-batch k needs: up to the end of its last mesh (running maximum: the pieces are uploaded in row order)
    std::vector<int64_t> need(nk);
    int64_t run_max = 0;
    for (int k = 0; k < nk; ++k) {
        for (int64_t s = seg0(k); s < seg1(k); ++s) {
            if (row1[s] < 0 || row2[s] < 0) return fail(c, LIP_EINVAL, "series: negative row");
            run_max = std::max(run_max, std::max(row1[s] + (c->h_off1[s + 1] - c->h_off1[s]), row2[s] + (c->h_off2[s + 1] - c->h_off2[s])));
        }
        if (run_max > n_rows) return fail(c, LIP_EINVAL, "series: a mesh does not lie inside `frames`");
        need[k] = run_max;
    }
    const bool stage = staging_threads() > 0 && (size_t)n_rows * 24 >= ((size_t)1 << 21) &&
                       (is_pageable(frames) || (T21 && is_pageable(T21)) || (T12 && is_pageable(T12)));
    const size_t slice = (size_t)1 << 20;
    std::atomic<int> bad(0);
    std::atomic<int>* chk = c->validate_inputs ? &bad : nullptr;
    std::vector<std::atomic<int>> in_left(nk), out_left(nk);
    for (auto& x : in_left) x.store(0);
    for (auto& x : out_left) x.store(0);
    const double* src = frames;
    if (stage) {                                             // pageable caller arrays: the pieces go through one pinned buffer
        if (c->st_in.ensure((size_t)n_rows * 24 + 256) != cudaSuccess) { cudaGetLastError(); return fail(c, LIP_ECUDA, "series: cannot pin the staging buffer"); }
        for (int k = 0; k < nk; ++k) {
            lip_ctx* kc = kid(k);
            if (kc->st_out.ensure(((size_t)kc->n1 + (size_t)kc->n2) * msz + 256) != cudaSuccess) { cudaGetLastError(); return fail(c, LIP_ECUDA, "series: cannot pin the output buffers"); }
        }
        if (!c->pool) c->pool = new StagePool();
        c->pool->start(staging_threads());
        int64_t r0 = 0;
        for (int k = 0; k < nk; ++k) {
            const size_t bytes = (size_t)(need[k] - r0) * 24;
            in_left[k].store(StagePool::slices_of(bytes, slice));
            if (bytes) c->pool->submit((char*)c->st_in.p + (size_t)r0 * 24, (const char*)frames + (size_t)r0 * 24, bytes, slice, &in_left[k], chk);
            r0 = need[k];
        }
        src = (const double*)c->st_in.p;
    } else if (c->validate_inputs && StagePool::nonfinite(frames, (size_t)need[nk - 1] * 24)) {
        return fail(c, LIP_EINVAL, "Input contains NaN or infinity.");
    }
    if (c->validate_inputs && (StagePool::nonfinite(lm1, bl) || StagePool::nonfinite(lm2, bl))) {
        if (c->pool) for (int k = 0; k < nk; ++k) c->pool->wait(in_left[k]);
        return fail(c, LIP_EINVAL, "Input contains NaN or infinity.");
    }
    cudaStream_t up = c->side[1];
    auto wait_all = [&] {
        cudaStreamSynchronize(up);
        for (int k = 0; k < nk; ++k) cudaStreamSynchronize(kid(k)->stream);
        if (c->pool) for (int k = 0; k < nk; ++k) { c->pool->wait(in_left[k]); c->pool->wait(out_left[k]); }
    };
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm1.p, lm1, bl, cudaMemcpyHostToDevice, up));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm2.p, lm2, bl, cudaMemcpyHostToDevice, up));
    const double* D = c->in_frames.as<double>();
    int rc = LIP_OK;
    int64_t r0 = 0;
    for (int k = 0; k < nk && rc == LIP_OK; ++k) {
        lip_ctx* kc = kid(k);
        if (stage) {
            c->pool->wait(in_left[k]);
            if (bad.load()) { rc = fail(c, LIP_EINVAL, "Input contains NaN or infinity."); break; }
        }
        static const bool trace = getenv("LIP_TRACE_PIPELINE") != nullptr;
        if (trace && !kc->tr[0]) for (auto& e : kc->tr) cudaEventCreate(&e);
        if (trace) cudaEventRecord(kc->tr[0], up);
        if (need[k] > r0) {
            cudaError_t e = cudaMemcpyAsync((char*)c->in_frames.p + (size_t)r0 * 24, (const char*)src + (size_t)r0 * 24,
                                            (size_t)(need[k] - r0) * 24, cudaMemcpyHostToDevice, up);
            if (e != cudaSuccess) { rc = fail(c, LIP_ECUDA, std::string("series upload: ") + cudaGetErrorString(e)); break; }
            r0 = need[k];
        }
        if (trace) cudaEventRecord(kc->tr[1], up);
        cudaEventRecord(c->ev_piece, up);
        cudaStreamWaitEvent(kc->stream, c->ev_piece, 0);
        kc->params = c->params; kc->map_bytes = c->map_bytes;
        const int64_t s0 = seg0(k);
        rc = series_rows(kc, row1, row2, s0, n_rows, c->h_off1.data(), c->h_off2.data(), kc->stream);
        if (rc == LIP_OK) {
            const DevIn in{D, D, c->in_lm1.as<double>() + 3 * c->h_lm_off[s0], c->in_lm2.as<double>() + 3 * c->h_lm_off[s0], kc->d_rows.as<int>()};
            char* o21 = T21 ? (stage ? (char*)kc->st_out.p : T21 + (size_t)c->h_off2[s0] * msz) : nullptr;
            char* o12 = T12 ? (stage ? (char*)kc->st_out.p + (size_t)kc->n2 * msz : T12 + (size_t)c->h_off1[s0] * msz) : nullptr;
            rc = host_compute_async(kc, o21, o12, nullptr, &in);
        }
        if (rc && kc != c) c->err = "series run: " + kc->err;
    }
    if (rc != LIP_OK) { wait_all(); return rc; }
    for (int k = 0; k < nk; ++k) {
        lip_ctx* kc = kid(k);
        cudaError_t e = cudaStreamSynchronize(kc->stream);
        if (e != cudaSuccess) { wait_all(); return fail(c, LIP_ECUDA, std::string("series sync: ") + cudaGetErrorString(e)); }
        if (!stage) continue;
        const int64_t s0 = seg0(k);
        const size_t o2 = T21 ? (size_t)kc->n2 * msz : 0, o1 = T12 ? (size_t)kc->n1 * msz : 0;
        out_left[k].store(StagePool::slices_of(o1, slice) + StagePool::slices_of(o2, slice));
        if (o2) c->pool->submit(T21 + (size_t)c->h_off2[s0] * msz, kc->st_out.p, o2, slice, &out_left[k]);
        if (o1) c->pool->submit(T12 + (size_t)c->h_off1[s0] * msz, (char*)kc->st_out.p + (size_t)kc->n2 * msz, o1, slice, &out_left[k]);
    }
    wait_all();
    if (getenv("LIP_TRACE_PIPELINE") && kid(0)->tr[0]) {
        for (int k = 0; k < nk; ++k) {
            float t[4];
            for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], kid(0)->tr[0], kid(k)->tr[i]);
            fprintf(stderr, "[lip series] sub-batch %d: H2D %.3f-%.3f ms, kernels end %.3f, D2H end %.3f\n", k, t[0], t[1], t[2], t[3]);
        }
    }
    c->last_run_split = c->n_kids >= 2; c->last_run_dsplit = false;
    return LIP_OK;
}

int lip_batch_get_transforms(lip_ctx* c, int filled, double* R21, double* t21, int32_t* iters21, int32_t* status21,
                             double* R12, double* t12, int32_t* iters12, int32_t* status12) {
    if (!c || !c->planned || !c->user_plan) return c ? fail(c, LIP_ENOPLAN, "no batch planned") : LIP_EINVAL;
    cudaSetDevice(c->device);
    if (c->last_run_split) {                                 // the last run was pipelined: collect from the sub-batches
        for (int k = 0; k < split_n(c); ++k) {
            const size_t o = (size_t)c->h_lm_off[split_seg(c)[k]];
            int rc = lip_batch_get_transforms(split_kids(c)[k], filled, R21 ? R21 + 9 * o : nullptr, t21 ? t21 + 3 * o : nullptr,
                                              iters21 ? iters21 + o : nullptr, status21 ? status21 + o : nullptr,
                                              R12 ? R12 + 9 * o : nullptr, t12 ? t12 + 3 * o : nullptr,
                                              iters12 ? iters12 + o : nullptr, status12 ? status12 + o : nullptr);
            if (rc) return fail(c, rc, split_kids(c)[k]->err);
        }
        return LIP_OK;
    }
    LIP_CUDA(c, cudaDeviceSynchronize());
    const LipArgs& a = c->args;
    const size_t K = (size_t)c->sumK;
    if (filled && !c->fill_done) {                           // the hot path resolves the fill on the fly; materialise it here
        fill_kernel<<<(c->M + 127) / 128, 128, 0, c->stream>>>(a);
        LIP_CUDA(c, cudaGetLastError());
        LIP_CUDA(c, cudaStreamSynchronize(c->stream));
        c->fill_done = true;
    }
    const double* R = filled ? a.fill_R : a.icp_R;
    const double* t = filled ? a.fill_t : a.icp_t;
    // bins [0, sumK): src = mesh1 -> T12 ICPs ; bins [sumK, 2 sumK): src = mesh2 -> T21 ICPs
    if (R12) LIP_CUDA(c, cudaMemcpy(R12, R, K * 72, cudaMemcpyDeviceToHost));
    if (t12) LIP_CUDA(c, cudaMemcpy(t12, t, K * 24, cudaMemcpyDeviceToHost));
    if (iters12) LIP_CUDA(c, cudaMemcpy(iters12, a.icp_iters, K * 4, cudaMemcpyDeviceToHost));
    if (status12) LIP_CUDA(c, cudaMemcpy(status12, a.icp_status, K * 4, cudaMemcpyDeviceToHost));
    if (R21) LIP_CUDA(c, cudaMemcpy(R21, R + 9 * K, K * 72, cudaMemcpyDeviceToHost));
    if (t21) LIP_CUDA(c, cudaMemcpy(t21, t + 3 * K, K * 24, cudaMemcpyDeviceToHost));
    if (iters21) LIP_CUDA(c, cudaMemcpy(iters21, a.icp_iters + K, K * 4, cudaMemcpyDeviceToHost));
    if (status21) LIP_CUDA(c, cudaMemcpy(status21, a.icp_status + K, K * 4, cudaMemcpyDeviceToHost));
    return LIP_OK;
}

int lip_batch_get_landmark_nn(lip_ctx* c, int side, int32_t* nn, double* dist) {
    if (!c || !c->planned || !c->user_plan) return c ? fail(c, LIP_ENOPLAN, "no batch planned") : LIP_EINVAL;
    if (side != 1 && side != 2) return fail(c, LIP_EINVAL, "side must be 1 or 2");
    cudaSetDevice(c->device);
    if (c->last_run_split) {
        for (int k = 0; k < split_n(c); ++k) {
            const size_t o = (size_t)(side == 1 ? c->h_off1 : c->h_off2)[split_seg(c)[k]];
            int rc = lip_batch_get_landmark_nn(split_kids(c)[k], side, nn ? nn + 2 * o : nullptr, dist ? dist + 2 * o : nullptr);
            if (rc) return fail(c, rc, split_kids(c)[k]->err);
        }
        return LIP_OK;
    }
    const LipArgs& a = c->args;
    const size_t first = side == 1 ? 0 : (size_t)c->n1, n = side == 1 ? (size_t)c->n1 : (size_t)c->n2;
    // only what the caller asked for travels (the drop-in's vertex colours need the nearest landmark alone); one synchronisation
    std::vector<int> i0(nn ? n : 0), i1(nn ? n : 0);
    std::vector<double> e0(dist ? n : 0), e1(dist ? n : 0);
    cudaStream_t st = c->stream;
    LIP_CUDA(c, cudaDeviceSynchronize());
    if (nn) {
        LIP_CUDA(c, cudaMemcpyAsync(i0.data(), a.nn0 + first, n * 4, cudaMemcpyDeviceToHost, st));
        LIP_CUDA(c, cudaMemcpyAsync(i1.data(), a.nn1 + first, n * 4, cudaMemcpyDeviceToHost, st));
    }
    if (dist) {
        LIP_CUDA(c, cudaMemcpyAsync(e0.data(), a.d0 + first, n * 8, cudaMemcpyDeviceToHost, st));
        LIP_CUDA(c, cudaMemcpyAsync(e1.data(), a.d1 + first, n * 8, cudaMemcpyDeviceToHost, st));
    }
    LIP_CUDA(c, cudaStreamSynchronize(st));
    for (size_t i = 0; i < n; ++i) {
        if (nn) { nn[2 * i] = i0[i]; nn[2 * i + 1] = i1[i]; }
        if (dist) { dist[2 * i] = e0[i]; dist[2 * i + 1] = e1[i]; }
    }
    return LIP_OK;
}

int lip_batch_get_counters(lip_ctx* c, int64_t counters[8]) {
    if (!c || !c->planned || !c->user_plan || !counters) return c ? fail(c, LIP_ENOPLAN, "no batch planned") : LIP_EINVAL;
    cudaSetDevice(c->device);
    if (c->last_run_split) {
        for (int i = 0; i < 8; ++i) counters[i] = 0;
        for (int k = 0; k < split_n(c); ++k) {
            int64_t part[8];
            int rc = lip_batch_get_counters(split_kids(c)[k], part);
            if (rc) return fail(c, rc, split_kids(c)[k]->err);
            for (int i = 0; i < 8; ++i) counters[i] += part[i];
        }
        return LIP_OK;
    }
    LIP_CUDA(c, cudaDeviceSynchronize());
    unsigned long long h[8];
    int sched[32];
    LIP_CUDA(c, cudaMemcpy(h, c->args.counters, sizeof(h), cudaMemcpyDeviceToHost));
    LIP_CUDA(c, cudaMemcpy(sched, c->args.sched, sizeof(sched), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) counters[i] = (int64_t)h[i];
    counters[0] = sched[0]; counters[1] = sched[1]; counters[5] = sched[3]; counters[4] = sched[7];
    counters[6] = c->last_launches;
    return LIP_OK;
}

int lip_batch_get_pair_counts(lip_ctx* c, int64_t pairs[3]) {
    if (!c || !c->planned || !c->user_plan || !pairs) return c ? fail(c, LIP_ENOPLAN, "no batch planned") : LIP_EINVAL;
    cudaSetDevice(c->device);
    pairs[0] = pairs[1] = pairs[2] = 0;
    if (c->last_run_split) {
        for (int k = 0; k < split_n(c); ++k) {
            int64_t part[3];
            int rc = lip_batch_get_pair_counts(split_kids(c)[k], part);
            if (rc) return fail(c, rc, split_kids(c)[k]->err);
            for (int i = 0; i < 3; ++i) pairs[i] += part[i];
        }
        return LIP_OK;
    }
    LIP_CUDA(c, cudaDeviceSynchronize());
    unsigned long long h[8];
    LIP_CUDA(c, cudaMemcpy(h, c->args.counters, sizeof(h), cudaMemcpyDeviceToHost));
    pairs[0] = (int64_t)h[4]; pairs[1] = (int64_t)h[5]; pairs[2] = (int64_t)h[7];
    return LIP_OK;
}

int lip_set_profiling(lip_ctx* c, int on) {
    if (!c) return LIP_EINVAL;
    cudaSetDevice(c->device);
    if (on && !c->ev[0]) {
        for (int i = 0; i <= LIP_N_STAGES; ++i) LIP_CUDA(c, cudaEventCreate(&c->ev[i]));
    }
    c->profiling = on != 0;
    c->ev_valid = false;
    return LIP_OK;
}

int lip_batch_get_stage_ms(lip_ctx* c, float ms[LIP_N_STAGES]) {
    if (!c || !ms) return LIP_EINVAL;
    if (!c->ev_valid) return fail(c, LIP_ENOPLAN, "stage times: profiling was not enabled for the last run");
    cudaSetDevice(c->device);
    LIP_CUDA(c, cudaEventSynchronize(c->ev[LIP_N_STAGES]));
    for (int i = 0; i < LIP_N_STAGES; ++i) LIP_CUDA(c, cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    return LIP_OK;
}

int lip_batch_get_bin_offsets(lip_ctx* c, int32_t* bin_off) {
    if (!c || !c->planned || !c->user_plan || !bin_off) return c ? fail(c, LIP_ENOPLAN, "no batch planned") : LIP_EINVAL;
    if (c->last_run_split) return fail(c, LIP_ENOPLAN, "bin offsets: the last run was a pipelined host run; use lip_batch_run_device");
    cudaSetDevice(c->device);
    LIP_CUDA(c, cudaDeviceSynchronize());
    LIP_CUDA(c, cudaMemcpy(bin_off, c->args.bin_off, ((size_t)c->B + 1) * 4, cudaMemcpyDeviceToHost));
    return LIP_OK;
}

// launch self_knn_kernel<KNN> for side (0: mesh1, 1: mesh2), every step-th vertex; qbase (device) = first query of each mesh
extern "C++" {
template <int KNN>
static int launch_self_knn(lip_ctx* c, const LipArgs& a, int side, int step, const int* d_qbase, int* d_out, int64_t max_n, cudaStream_t st) {
    int smem_cap = 0;
    LIP_CUDA(c, cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const size_t bucket_bytes = (LIP_KNN_BUCKETS + 1) * 4;
    int smem_pts = (int)std::min<int64_t>(max_n, (smem_cap - 2048 - (int)bucket_bytes) / 16);   // fp32 tile of the whole mesh
    if (getenv("LIP_KNN_BRUTE")) smem_pts = 0;               // debugging aid: the float64 brute-force path for every mesh
    const size_t smem = (size_t)smem_pts * 16 + bucket_bytes;
    LIP_CUDA(c, cudaFuncSetAttribute(self_knn_kernel<KNN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // every CTA of a mesh sorts it again: a few CTAs per mesh only (about 1,024 sorted positions each)
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(65535, (max_n + 1023) / 1024));
    self_knn_kernel<KNN><<<dim3((unsigned)c->S, gy), 256, smem, st>>>(a, side, step, d_qbase, d_out, smem_pts);
    LIP_CUDA(c, cudaGetLastError());
    return LIP_OK;
}
}  // extern "C++"

// range check of uploaded maps (they index device arrays): enqueue, then wait for the verdict only
static int check_maps_enqueue(lip_ctx* c, const long long* d_T21, const long long* d_T12, cudaStream_t st) {
    LIP_CUDA(c, c->d_err.ensure(16)); LIP_CUDA(c, c->h_err.ensure(16));
    if (!c->ev_chk) LIP_CUDA(c, cudaEventCreateWithFlags(&c->ev_chk, cudaEventDisableTiming));
    LIP_CUDA(c, cudaMemsetAsync(c->d_err.p, 0, 4, st));
    int64_t max_rows = 1;
    for (size_t s = 0; s < (size_t)c->S; ++s)
        max_rows = std::max(max_rows, (c->h_off1[s + 1] - c->h_off1[s]) + (c->h_off2[s + 1] - c->h_off2[s]));
    const int64_t want = (4 * (int64_t)c->sm_count + c->S - 1) / c->S;
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>((max_rows + 255) / 256, want));
    LipArgs a = c->args;
    map_check_kernel<<<dim3((unsigned)c->S, gy), 256, 0, st>>>(a, d_T21, d_T12, c->d_err.as<int>());
    LIP_CUDA(c, cudaGetLastError());
    LIP_CUDA(c, cudaMemcpyAsync(c->h_err.p, c->d_err.p, 4, cudaMemcpyDeviceToHost, st));
    LIP_CUDA(c, cudaEventRecord(c->ev_chk, st));
    return LIP_OK;
}
static int check_maps_verdict(lip_ctx* c, cudaStream_t st, const char* who) {
    LIP_CUDA(c, cudaEventSynchronize(c->ev_chk));
    if (*(const int*)c->h_err.p) {
        cudaStreamSynchronize(st);                           // nothing of this call may still be reading the caller's arrays
        return fail(c, LIP_EINVAL, std::string(who) + ": map entry out of range for its segment pair");
    }
    return LIP_OK;
}

// evaluate_p2p_matching on device-resident inputs and maps: enqueues the kernels and the copy of the counts
static int evaluate_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, const double* d_lm1, const double* d_lm2,
                           const long long* d_T21, const long long* d_T12, double voxel_size, double sample_proportion,
                           cudaStream_t st, long long* d_counts_out, const int* mesh_row) {
    const size_t S = (size_t)c->S;
    const int step = std::max((int)(1.0 / sample_proportion), 1);              // p2p_matching_evaluation.py:40
    std::vector<int>& qbase = c->h_qbase_eval;                 // kept in the context: the upload below is asynchronous
    qbase.assign(S + 1, 0);
    int64_t max_n2 = 0;
    for (size_t s = 0; s < S; ++s) {
        const int64_t n2s = c->h_off2[s + 1] - c->h_off2[s];
        qbase[s + 1] = qbase[s] + (int)((n2s + step - 1) / step);
        max_n2 = std::max(max_n2, n2s);
    }
    const size_t o_cnt = align_up((size_t)c->B * 4), o_qb = o_cnt + align_up(S * 48), o_nn = o_qb + align_up((S + 1) * 4);
    LIP_CUDA(c, c->misc2.ensure(o_nn + (size_t)qbase[S] * 8 + 16));
    LipArgs a = c->args;
    a.xyz1 = d_xyz1; a.xyz2 = d_xyz2; a.lm1 = d_lm1; a.lm2 = d_lm2; a.mesh_row = mesh_row;
    int* lm_idx = c->misc2.as<int>();
    long long* cnt = (long long*)(c->misc2.as<char>() + o_cnt);
    int* d_qbase = (int*)(c->misc2.as<char>() + o_qb);
    int* d_nn2 = (int*)(c->misc2.as<char>() + o_nn);
    LIP_CUDA(c, cudaMemcpyAsync(d_qbase, qbase.data(), (S + 1) * 4, cudaMemcpyHostToDevice, st));
    eval_landmark_nn_kernel<<<(c->B + 3) / 4, 128, 0, st>>>(a, lm_idx);
    int rc = launch_self_knn<2>(c, a, 1, step, d_qbase, d_nn2, max_n2, st);
    if (rc) return rc;
    eval_scores_kernel<<<(unsigned)S, 256, 0, st>>>(a, lm_idx, d_nn2, d_qbase, d_T21, d_T12, 2.0 * voxel_size, step, cnt);
    LIP_CUDA(c, cudaGetLastError());
    if (d_counts_out) {                                      // device-resident caller: the counts stay on the device
        LIP_CUDA(c, cudaMemcpyAsync(d_counts_out, cnt, S * 48, cudaMemcpyDeviceToDevice, st));
        return LIP_OK;
    }
    LIP_CUDA(c, c->h_counts.ensure(S * 48));                 // (valid after the stream has been synchronised)
    LIP_CUDA(c, cudaMemcpyAsync(c->h_counts.p, cnt, S * 48, cudaMemcpyDeviceToHost, st));
    return LIP_OK;
}
static void scores_from_counts(const long long* h, size_t S, bool with_T12, double* scores, int64_t* counts) {
    for (size_t s = 0; s < S; ++s) {
        const long long* k = &h[6 * s];
        scores[3 * s] = (double)k[2] / (double)k[3];                            // continuity
        scores[3 * s + 1] = (double)k[0] / (double)k[1];                        // landmark precision
        scores[3 * s + 2] = with_T12 ? (double)k[4] / (double)k[5] : std::numeric_limits<double>::quiet_NaN();   // cycle
        if (counts) for (int i = 0; i < 6; ++i) counts[6 * s + i] = k[i];
    }
}

int lip_batch_evaluate_host(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                            const int64_t* T21, const int64_t* T12, double voxel_size, double sample_proportion,
                            double* scores, int64_t* counts) {
    Range nvtx_("lip_batch_evaluate_host");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "evaluate: no batch planned");
    if (!xyz1 || !xyz2 || !lm1 || !lm2 || !T21 || !scores) return fail(c, LIP_EINVAL, "evaluate: null pointer");
    if (!(sample_proportion > 0.0)) return fail(c, LIP_EINVAL, "evaluate: sample_proportion must be > 0");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const size_t b1 = (size_t)c->n1 * 24, b2 = (size_t)c->n2 * 24, bl = (size_t)c->sumK * 24;
    LIP_CUDA(c, c->in_xyz1.ensure(b1)); LIP_CUDA(c, c->in_xyz2.ensure(b2));
    LIP_CUDA(c, c->in_lm1.ensure(bl)); LIP_CUDA(c, c->in_lm2.ensure(bl));
    LIP_CUDA(c, c->out_T21.ensure((size_t)c->n2 * 8)); LIP_CUDA(c, c->out_T12.ensure((size_t)c->n1 * 8));
    StageCall sc(c);
    int src = sc.begin(b1 + b2 + ((size_t)c->n1 + (size_t)c->n2) * 8 + 2048, 0);
    if (src) return src;
    if ((src = sc.h2d(c->out_T21.p, T21, (size_t)c->n2 * 8, st))) return src;       // the maps first: they are checked on the
    if (T12 && (src = sc.h2d(c->out_T12.p, T12, (size_t)c->n1 * 8, st))) return src;  // device while the coordinates travel
    if ((src = check_maps_enqueue(c, c->out_T21.as<long long>(), T12 ? c->out_T12.as<long long>() : nullptr, st))) return src;
    if ((src = sc.h2d(c->in_xyz1.p, xyz1, b1, st)) || (src = sc.h2d(c->in_xyz2.p, xyz2, b2, st))) return src;
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm1.p, lm1, bl, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm2.p, lm2, bl, cudaMemcpyHostToDevice, st));
    if ((src = check_maps_verdict(c, st, "evaluate"))) return src;
    int rc = evaluate_device(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->in_lm1.as<double>(), c->in_lm2.as<double>(),
                             c->out_T21.as<long long>(), T12 ? c->out_T12.as<long long>() : nullptr, voxel_size, sample_proportion, st);
    if (rc) return rc;
    LIP_CUDA(c, cudaStreamSynchronize(st));
    scores_from_counts((const long long*)c->h_counts.p, (size_t)c->S, T12 != nullptr, scores, counts);
    c->last_launches = 2;
    return LIP_OK;
}

static int series_rows(lip_ctx* kc, const int64_t* row1, const int64_t* row2, int64_t s0, int64_t n_rows, const int64_t* off1,
                       const int64_t* off2, cudaStream_t st);
int lip_batch_evaluate_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, const double* d_lm1, const double* d_lm2,
                              const int64_t* d_T21, const int64_t* d_T12, double voxel_size, double sample_proportion,
                              int64_t* d_counts, const int64_t* row1, const int64_t* row2, int64_t n_rows, void* stream) {
    Range nvtx_("lip_batch_evaluate_device");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "evaluate: no batch planned");
    if (!d_xyz1 || !d_xyz2 || !d_lm1 || !d_lm2 || !d_T21 || !d_counts) return fail(c, LIP_EINVAL, "evaluate: null pointer");
    if (!(sample_proportion > 0.0)) return fail(c, LIP_EINVAL, "evaluate: sample_proportion must be > 0");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const int* rows = nullptr;
    if (row1 || row2) {                                      // series layout: d_xyz1 holds every mesh once, d_xyz2 is ignored
        if (!row1 || !row2 || n_rows <= 0) return fail(c, LIP_EINVAL, "evaluate: row1, row2 and n_rows go together");
        int rc = series_rows(c, row1, row2, 0, n_rows, c->h_off1.data(), c->h_off2.data(), st);
        if (rc) return rc;
        rows = c->d_rows.as<int>();
    }
    return evaluate_device(c, d_xyz1, d_xyz2, d_lm1, d_lm2, (const long long*)d_T21, (const long long*)d_T12, voxel_size,
                           sample_proportion, st, (long long*)d_counts, rows);
}

// double_direction_refinement on device-resident coordinates and maps (d_T21 rewritten in place)
extern "C++" {
template <typename IT>
static int launch_refine(lip_ctx* c, const LipArgs& a, const int* nn5, long long* d_T21, const long long* d_T12, IT* scratch,
                         size_t max_entries, size_t max_entries_all, cudaStream_t st) {
    int smem_cap = 0;
    LIP_CUDA(c, cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    smem_cap -= 1024;
    // Two layouts: every table in shared memory (fast steps, few segment pairs per SM) or only the hot ones (slower steps — the
    // cold tables are L2 accesses — but more pairs per SM).  One lane walks a pair from start to end, so the time is
    // waves x time per pair: measured 4.7 ms per maize pair with all tables in SMEM, ~8 ms with the hot ones only.
    const size_t smem_hot = std::min<size_t>((size_t)smem_cap, max_entries * sizeof(IT) + 16);
    const size_t smem_all = std::min<size_t>((size_t)smem_cap, max_entries_all * sizeof(IT) + 16);
    const size_t sm_smem = 227 * 1024;
    // (a sub-batch of a pipelined call shares the GPU with its siblings' kernels: count their segment pairs as well)
    const size_t in_flight = (size_t)c->S * (c->is_kid ? LIP_PIPELINE_CHUNKS : 1);
    auto waves = [&](size_t smem) { const size_t per_sm = std::max<size_t>(1, std::min<size_t>(32, sm_smem / (smem + 1024))); return (in_flight + per_sm * c->sm_count - 1) / (per_sm * c->sm_count); };
    static const int force = getenv("LIP_REFINE_LAYOUT") ? atoi(getenv("LIP_REFINE_LAYOUT")) : 0;     // 1: all in SMEM, 2: hot only
    const bool all = force ? force == 1 : (max_entries_all * sizeof(IT) + 16 <= (size_t)smem_cap && 1.0 * waves(smem_all) <= 1.7 * waves(smem_hot));
    if (all) {
        LIP_CUDA(c, cudaFuncSetAttribute(refine_kernel<IT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all));
        refine_kernel<IT, true><<<(unsigned)c->S, 32, smem_all, st>>>(a, nn5, d_T21, d_T12, scratch, (int)(smem_all / sizeof(IT)));
    } else {
        LIP_CUDA(c, cudaFuncSetAttribute(refine_kernel<IT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hot));
        refine_kernel<IT, false><<<(unsigned)c->S, 32, smem_hot, st>>>(a, nn5, d_T21, d_T12, scratch, (int)(smem_hot / sizeof(IT)));
    }
    LIP_CUDA(c, cudaGetLastError());
    return LIP_OK;
}
}  // extern "C++"
static int refine_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, long long* d_T21, const long long* d_T12,
                         cudaStream_t st) {
    const size_t n1 = (size_t)c->n1, n2 = (size_t)c->n2, S = (size_t)c->S;
    size_t max_entries = 0, max_entries_all = 0;
    int64_t max_n1 = 0, max_n = 0;
    for (size_t s = 0; s < S; ++s) {
        const int64_t a1 = c->h_off1[s + 1] - c->h_off1[s], a2 = c->h_off2[s + 1] - c->h_off2[s];
        max_entries = std::max(max_entries, (size_t)(2 * a2 + a1 + 4 + (a1 + 31) / 32 * 2));      // the hot tables of refine_kernel (shared memory)
        max_entries_all = std::max(max_entries_all, (size_t)(3 * a2 + 3 * a1 + 4 + (a1 + 31) / 32 * 2));   // ... and all of them
        max_n1 = std::max(max_n1, a1);
        max_n = std::max(max_n, std::max(a1, a2));
    }
    const bool narrow = max_n < 65535;                       // 16-bit tables: 0xFFFF = none
    const size_t esz = narrow ? 2 : 4;
    const size_t o_qb = align_up(n1 * 20), o_scratch = o_qb + align_up((S + 1) * 4);
    LIP_CUDA(c, c->misc.ensure(o_scratch + (3 * n2 + 4 * n1 + 16 * S) * esz + 16));   // cold tables always, hot ones when they do not fit SMEM
    LipArgs a = c->args;
    a.xyz1 = d_xyz1; a.xyz2 = d_xyz2;
    int* nn4 = c->misc.as<int>();                            // 5 neighbours per vertex of mesh1, row = [self, 4 nearest]
    int* d_qbase = (int*)(c->misc.as<char>() + o_qb);
    void* scratch = c->misc.as<char>() + o_scratch;
    std::vector<int>& qbase = c->h_qbase_ref;
    qbase.resize(S + 1);
    for (size_t s = 0; s <= S; ++s) qbase[s] = (int)c->h_off1[s];
    LIP_CUDA(c, cudaMemcpyAsync(d_qbase, qbase.data(), (S + 1) * 4, cudaMemcpyHostToDevice, st));
    int rc5 = launch_self_knn<5>(c, a, 0, 1, d_qbase, nn4, max_n1, st);
    if (rc5) return rc5;
    return narrow ? launch_refine<unsigned short>(c, a, nn4, d_T21, d_T12, (unsigned short*)scratch, max_entries, max_entries_all, st)
                  : launch_refine<int>(c, a, nn4, d_T21, d_T12, (int*)scratch, max_entries, max_entries_all, st);
}

int lip_batch_refine_host(lip_ctx* c, const double* xyz1, const double* xyz2, int64_t* T21, const int64_t* T12) {
    Range nvtx_("lip_batch_refine_host");
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "refine: no batch planned");
    if (!xyz1 || !xyz2 || !T21 || !T12) return fail(c, LIP_EINVAL, "refine: null pointer");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const size_t n1 = (size_t)c->n1, n2 = (size_t)c->n2, S = (size_t)c->S;
    for (size_t s = 0; s < S; ++s)
        if (c->h_off1[s + 1] - c->h_off1[s] < 5) return fail(c, LIP_EINVAL, "refine: every mesh1 needs at least 5 vertices (KDTree.query(k=5))");
    LIP_CUDA(c, c->in_xyz1.ensure(n1 * 24)); LIP_CUDA(c, c->in_xyz2.ensure(n2 * 24));
    LIP_CUDA(c, c->out_T21.ensure(n2 * 8)); LIP_CUDA(c, c->out_T12.ensure(n1 * 8));
    StageCall sc(c);
    int rc = sc.begin((n1 + n2) * 32 + 2048, n2 * 8);
    if (rc) return rc;
    // the maps are used as indices on the device: they go first and are range-checked there while the coordinates travel
    if ((rc = sc.h2d(c->out_T21.p, T21, n2 * 8, st)) || (rc = sc.h2d(c->out_T12.p, T12, n1 * 8, st)) ||
        (rc = check_maps_enqueue(c, c->out_T21.as<long long>(), c->out_T12.as<long long>(), st)) ||
        (rc = sc.h2d(c->in_xyz1.p, xyz1, n1 * 24, st)) || (rc = sc.h2d(c->in_xyz2.p, xyz2, n2 * 24, st)) ||
        (rc = check_maps_verdict(c, st, "refine"))) return rc;
    rc = refine_device(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->out_T21.as<long long>(), c->out_T12.as<long long>(), st);
    if (rc) return rc;
    if ((rc = sc.d2h(T21, c->out_T21.p, n2 * 8, st)) || (rc = sc.finish(st))) return rc;
    c->last_launches = 2;
    return LIP_OK;
}

int lip_batch_match_host(lip_ctx* c, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                         int refine, double voxel_size, double sample_proportion, int64_t* T21, int64_t* T12, double* scores) {
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "match: no batch planned");
    if (c->directions != LIP_DIR_BOTH) return fail(c, LIP_EINVAL, "match: the batch must be planned for both directions");
    if (c->map_bytes != 8) return fail(c, LIP_EINVAL, "match: the refinement and the scores work on int64 maps (lip_set_map_dtype(ctx, 64))");
    if (!xyz1 || !xyz2 || !lm1 || !lm2 || !T21 || !T12 || !scores) return fail(c, LIP_EINVAL, "match: null pointer");
    if (!(sample_proportion > 0.0)) return fail(c, LIP_EINVAL, "match: sample_proportion must be > 0");
    if (refine)
        for (int64_t s = 0; s < c->S; ++s)
            if (c->h_off1[s + 1] - c->h_off1[s] < 5) return fail(c, LIP_EINVAL, "match: refinement needs at least 5 vertices in every mesh1");
    cudaSetDevice(c->device);
    if (c->n_kids >= 2) {                                    // large batch: sub-batches pipelined like the hot path alone
        const MatchOpts mo{refine, voxel_size, sample_proportion};
        int prc = run_host_split(c, xyz1, xyz2, lm1, lm2, T21, T12, &mo);
        if (prc) return prc;
        int64_t launches = 0;
        for (int k = 0; k < c->n_kids; ++k) {
            lip_ctx* kc = c->kids[k];
            scores_from_counts((const long long*)kc->h_counts.p, (size_t)kc->S, true, scores + 3 * c->kid_seg[k], nullptr);
            launches += kc->last_launches;
        }
        c->last_launches = launches;
        return LIP_OK;
    }
    if (c->validate_inputs) {
        int vrc = check_finite(c, xyz1, xyz2, lm1, lm2);
        if (vrc) return vrc;
    }
    cudaStream_t st = c->stream;
    const size_t b1 = (size_t)c->n1 * 24, b2 = (size_t)c->n2 * 24, bl = (size_t)c->sumK * 24;
    LIP_CUDA(c, c->in_xyz1.ensure(b1)); LIP_CUDA(c, c->in_xyz2.ensure(b2));
    LIP_CUDA(c, c->in_lm1.ensure(bl)); LIP_CUDA(c, c->in_lm2.ensure(bl));
    LIP_CUDA(c, c->out_T21.ensure((size_t)c->n2 * 8)); LIP_CUDA(c, c->out_T12.ensure((size_t)c->n1 * 8));
    StageCall sc(c);
    int src = sc.begin(b1 + b2 + 1024, ((size_t)c->n1 + (size_t)c->n2) * 8 + 1024);
    if (src) return src;
    if ((src = sc.h2d(c->in_xyz1.p, xyz1, b1, st)) || (src = sc.h2d(c->in_xyz2.p, xyz2, b2, st))) return src;
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm1.p, lm1, bl, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_lm2.p, lm2, bl, cudaMemcpyHostToDevice, st));
    const double *dx1 = c->in_xyz1.as<double>(), *dx2 = c->in_xyz2.as<double>(), *dl1 = c->in_lm1.as<double>(), *dl2 = c->in_lm2.as<double>();
    long long *d21 = c->out_T21.as<long long>(), *d12 = c->out_T12.as<long long>();
    c->last_run_split = false;
    int rc = run_stages(c, dx1, dx2, dl1, dl2, d21, d12, nullptr, kFullPipeline, st);      // maps stay on the device
    if (rc) return rc;
    const int64_t hot_launches = c->last_launches;
    if (refine && (rc = refine_device(c, dx1, dx2, d21, d12, st)) != LIP_OK) return rc;
    if ((rc = evaluate_device(c, dx1, dx2, dl1, dl2, d21, d12, voxel_size, sample_proportion, st)) != LIP_OK) return rc;
    if ((rc = sc.d2h(T21, d21, (size_t)c->n2 * 8, st)) || (rc = sc.d2h(T12, d12, (size_t)c->n1 * 8, st)) || (rc = sc.finish(st))) return rc;
    scores_from_counts((const long long*)c->h_counts.p, (size_t)c->S, true, scores, nullptr);
    c->last_launches = hot_launches + (refine ? 2 : 0) + 2;
    return LIP_OK;
}

// interpolation_segment_level for the planned batch on device-resident arrays; d_err: one int, zeroed here
static int interpolate_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, const long long* d_T21, const long long* d_T12,
                              double alpha, double* d_out, int* d_err, cudaStream_t st) {
    LipArgs a = c->args;
    a.xyz1 = d_xyz1; a.xyz2 = d_xyz2;
    int64_t max_rows = 1;
    for (size_t s = 0; s < (size_t)c->S; ++s)
        max_rows = std::max(max_rows, (c->h_off1[s + 1] - c->h_off1[s]) + (c->h_off2[s + 1] - c->h_off2[s]));
    LIP_CUDA(c, cudaMemsetAsync(d_err, 0, 4, st));
    // rows of one segment pair over blockIdx.y: enough CTAs for a full wave, at most one per 256 rows
    const int64_t want = (4 * (int64_t)c->sm_count + c->S - 1) / c->S;
    const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>((max_rows + 255) / 256, std::max<int64_t>(want, 1)));
    interpolate_kernel<<<dim3((unsigned)c->S, gy), 256, 0, st>>>(a, d_T21, d_T12, alpha, d_out, d_err);
    LIP_CUDA(c, cudaGetLastError());
    c->last_launches = 1;
    return LIP_OK;
}

int lip_batch_interpolate_device(lip_ctx* c, const double* d_xyz1, const double* d_xyz2, const int64_t* d_T21, const int64_t* d_T12,
                                 double alpha, double* d_out, void* stream) {
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "interpolate: no batch planned");
    if (!d_xyz1 || !d_xyz2 || !d_T21 || !d_T12 || !d_out) return fail(c, LIP_EINVAL, "interpolate: null pointer");
    cudaSetDevice(c->device);
    LIP_CUDA(c, c->misc2.ensure(16));
    return interpolate_device(c, d_xyz1, d_xyz2, (const long long*)d_T21, (const long long*)d_T12, alpha, d_out,
                              c->misc2.as<int>(), (cudaStream_t)stream);
}

int lip_batch_interpolate_host(lip_ctx* c, const double* xyz1, const double* xyz2, const int64_t* T21, const int64_t* T12,
                               double alpha, double* out) {
    if (!c) return LIP_EINVAL;
    if (!c->planned || !c->user_plan || c->h_off1.empty()) return fail(c, LIP_ENOPLAN, "interpolate: no batch planned");
    if (!xyz1 || !xyz2 || !T21 || !T12 || !out) return fail(c, LIP_EINVAL, "interpolate: null pointer");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const size_t b1 = (size_t)c->n1 * 24, b2 = (size_t)c->n2 * 24;
    LIP_CUDA(c, c->in_xyz1.ensure(b1)); LIP_CUDA(c, c->in_xyz2.ensure(b2));
    LIP_CUDA(c, c->out_T21.ensure((size_t)c->n2 * 8)); LIP_CUDA(c, c->out_T12.ensure((size_t)c->n1 * 8));
    LIP_CUDA(c, c->misc.ensure(b1 + b2)); LIP_CUDA(c, c->misc2.ensure(16));
    StageCall sc(c);
    int rc = sc.begin(b1 + b2 + ((size_t)c->n1 + (size_t)c->n2) * 8 + 2048, b1 + b2 + 1024);
    if (rc) return rc;
    if ((rc = sc.h2d(c->in_xyz1.p, xyz1, b1, st)) || (rc = sc.h2d(c->in_xyz2.p, xyz2, b2, st)) ||
        (rc = sc.h2d(c->out_T21.p, T21, (size_t)c->n2 * 8, st)) || (rc = sc.h2d(c->out_T12.p, T12, (size_t)c->n1 * 8, st))) return rc;
    rc = interpolate_device(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->out_T21.as<long long>(),
                                c->out_T12.as<long long>(), alpha, c->misc.as<double>(), c->misc2.as<int>(), st);
    if (rc) return rc;
    int h_err = 0;
    if ((rc = sc.d2h(out, c->misc.p, b1 + b2, st))) return rc;
    LIP_CUDA(c, cudaMemcpyAsync(&h_err, c->misc2.p, 4, cudaMemcpyDeviceToHost, st));
    if ((rc = sc.finish(st))) return rc;
    if (h_err) return fail(c, LIP_EINVAL, "interpolate: map entry out of range");
    return LIP_OK;
}

int lip_icp_host(lip_ctx* c, int64_t P, const double* src, const int64_t* src_off, const double* tgt,
                 const int64_t* tgt_off, double* R, double* t, int32_t* iters, int32_t* status) {
    if (!c) return LIP_EINVAL;
    c->user_plan = false; c->n_kids = 0; c->n_dkids = 0;      // whatever happens below, the caller's batch plan is gone
    if (P <= 0 || !src || !src_off || !tgt || !tgt_off) return fail(c, LIP_EINVAL, "icp: bad arguments");
    cudaSetDevice(c->device);
    std::vector<int64_t> lm_off(P + 1);
    for (int64_t i = 0; i <= P; ++i) lm_off[i] = i;
    // mesh1 = target, mesh2 = source, one (unused) landmark per mesh -> one whole-mesh partition each
    int rc = plan_impl(c, P, tgt_off, src_off, lm_off.data(), LIP_DIR_T21);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    const size_t b1 = (size_t)c->n1 * 24, b2 = (size_t)c->n2 * 24;
    LIP_CUDA(c, c->in_xyz1.ensure(b1)); LIP_CUDA(c, c->in_xyz2.ensure(b2));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz1.p, tgt, b1, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz2.p, src, b2, cudaMemcpyHostToDevice, st));
    // the landmark of mesh m is read at lm[3*m]; any finite point will do (P <= n because no mesh is empty)
    rc = run_stages(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->in_xyz1.as<double>(), c->in_xyz2.as<double>(),
                    nullptr, nullptr, nullptr, ST_PARTITION | ST_ICP, st);
    if (rc) return rc;
    const LipArgs& a = c->args;
    const size_t K = (size_t)P;
    if (R) LIP_CUDA(c, cudaMemcpyAsync(R, a.icp_R + 9 * K, K * 72, cudaMemcpyDeviceToHost, st));
    if (t) LIP_CUDA(c, cudaMemcpyAsync(t, a.icp_t + 3 * K, K * 24, cudaMemcpyDeviceToHost, st));
    if (iters) LIP_CUDA(c, cudaMemcpyAsync(iters, a.icp_iters + K, K * 4, cudaMemcpyDeviceToHost, st));
    if (status) LIP_CUDA(c, cudaMemcpyAsync(status, a.icp_status + K, K * 4, cudaMemcpyDeviceToHost, st));
    LIP_CUDA(c, cudaStreamSynchronize(st));
    c->planned = false;                                      // internal plan: not a user batch
    return LIP_OK;
}

int lip_nn_host(lip_ctx* c, const double* src, int64_t n, const double* tgt, int64_t m, double* dist, int64_t* idx) {
    if (!c) return LIP_EINVAL;
    c->user_plan = false; c->n_kids = 0; c->n_dkids = 0;      // whatever happens below, the caller's batch plan is gone
    if (n <= 0 || m <= 0 || !src || !tgt || !idx) return fail(c, LIP_EINVAL, "nn: bad arguments");
    cudaSetDevice(c->device);
    const int64_t off1[2] = {0, m}, off2[2] = {0, n}, lmo[2] = {0, 1};
    int rc = plan_impl(c, 1, off1, off2, lmo, LIP_DIR_T21);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    LIP_CUDA(c, c->in_xyz1.ensure((size_t)m * 24)); LIP_CUDA(c, c->in_xyz2.ensure((size_t)n * 24));
    LIP_CUDA(c, c->out_T21.ensure((size_t)n * 8)); LIP_CUDA(c, c->misc.ensure((size_t)n * 8));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz1.p, tgt, (size_t)m * 24, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz2.p, src, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    rc = run_stages(c, c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), c->in_xyz1.as<double>(), c->in_xyz2.as<double>(),
                    c->out_T21.as<long long>(), nullptr, c->misc.as<double>(), ST_GRID | ST_QUERY, st);
    if (rc) return rc;
    LIP_CUDA(c, cudaMemcpyAsync(idx, c->out_T21.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (dist) LIP_CUDA(c, cudaMemcpyAsync(dist, c->misc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    LIP_CUDA(c, cudaStreamSynchronize(st));
    c->planned = false;
    return LIP_OK;
}

int lip_kabsch_host(lip_ctx* c, const double* A, const double* Bp, int64_t n, double* R, double* t, int32_t* status) {
    if (!c) return LIP_EINVAL;
    if (n <= 0 || n >= ((int64_t)1 << 30) || !A || !Bp || !R || !t) return fail(c, LIP_EINVAL, "kabsch: bad arguments");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    LIP_CUDA(c, c->in_xyz1.ensure((size_t)n * 24)); LIP_CUDA(c, c->in_xyz2.ensure((size_t)n * 24));
    LIP_CUDA(c, c->misc.ensure(128));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz1.p, A, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz2.p, Bp, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    double* d = c->misc.as<double>();
    kabsch_kernel<<<1, 256, 0, st>>>(c->in_xyz1.as<double>(), c->in_xyz2.as<double>(), (int)n, d, d + 9, (int*)(d + 12));
    LIP_CUDA(c, cudaGetLastError());
    double h[13];
    LIP_CUDA(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st));
    LIP_CUDA(c, cudaStreamSynchronize(st));
    memcpy(R, h, 72); memcpy(t, h + 9, 24);
    if (status) memcpy(status, h + 12, 4);
    c->last_launches = 1;
    return LIP_OK;
}

int lip_triangle_angles_host(lip_ctx* c, const double* abc, int64_t n, double* angles) {
    if (!c) return LIP_EINVAL;
    if (n <= 0 || n >= ((int64_t)1 << 30) || !abc || !angles) return fail(c, LIP_EINVAL, "triangle_angles: bad arguments");
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    LIP_CUDA(c, c->in_xyz1.ensure((size_t)n * 24)); LIP_CUDA(c, c->in_xyz2.ensure((size_t)n * 24));
    LIP_CUDA(c, cudaMemcpyAsync(c->in_xyz1.p, abc, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    triangle_angles_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(c->in_xyz1.as<double>(), (int)n, c->params.cos_den_eps,
                                                                  c->in_xyz2.as<double>());
    LIP_CUDA(c, cudaGetLastError());
    LIP_CUDA(c, cudaMemcpyAsync(angles, c->in_xyz2.p, (size_t)n * 24, cudaMemcpyDeviceToHost, st));
    LIP_CUDA(c, cudaStreamSynchronize(st));
    c->last_launches = 1;
    return LIP_OK;
}

}  // extern "C"
