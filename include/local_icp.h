/*
 * local_icp.h — C ABI of liblocalicp.so: B200 (sm_100a) local-ICP point-to-point matching.
 *
 * Drop-in boundary for the hot path of romi/4d_plant_analysis (paths below are relative to the
 * reference root).  The reference has no FFI of its own — its boundary is five Python functions —
 * so each entry point names the Python function whose arithmetic it replaces; INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; float64 coordinates in row-major (n,3) arrays exactly as numpy
 *     hands them over; int64 CSR offsets; int64 index outputs (numpy intp), like the reference.
 *   - *_host entry points take HOST pointers (pageable or pinned) and do H2D, kernels, D2H and a
 *     stream synchronise; *_device entry points take DEVICE pointers, enqueue on `stream`
 *     (a cudaStream_t passed as void*) and return without synchronising.
 *   - every function returns 0 on success, a negative LIP_E* code otherwise;
 *     lip_last_error(ctx) gives the message.
 *   - there is no CPU path: without a CUDA device lip_create fails with LIP_ENODEV.
 */
#ifndef LOCAL_ICP_H
#define LOCAL_ICP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LIP_ABI_VERSION 2 /* returned by lip_version(); the Python layer refuses a library of another version */

#define LIP_OK        0
#define LIP_ENODEV   -1   /* no CUDA device / CUDA runtime error at start-up */
#define LIP_EINVAL   -2   /* bad argument (empty mesh, K mismatch, K too large, null pointer ...) */
#define LIP_ECUDA    -3   /* CUDA runtime error, see lip_last_error */
#define LIP_ENOPLAN  -4   /* run called before plan */

#define LIP_DIR_T21   1   /* mesh2 -> mesh1 map  (p2p_matching_seg_icp.py:173) */
#define LIP_DIR_T12   2   /* mesh1 -> mesh2 map  (p2p_matching_seg_icp.py:175) */
#define LIP_DIR_BOTH  3

#define LIP_MAX_LANDMARKS 2048   /* per segment */

typedef struct lip_ctx lip_ctx;

/* Constants the reference hard-codes (registration/registration_icp.py:79,87;
 * p2p_matching_in_organ/p2p_matching_seg_icp.py:30-32,471).  lip_default_params fills the
 * reference values; reject_dist is the north star's correspondence-reject hook — the reference
 * rejects nothing, so the default is +inf. */
typedef struct lip_params {
    int32_t max_iter;        /* 500 */
    double  cost_rise_tol;   /* 1e-4 */
    double  cost_flat_tol;   /* 0.01 */
    double  obtuse_eps;      /* 1e-5 */
    double  cos_den_eps;     /* 1e-5 */
    double  reject_dist;     /* +inf: pairs farther than this are left out of the Kabsch sums */
} lip_params;

void lip_default_params(lip_params* p);

/* ---- context -------------------------------------------------------------------------------- */
int  lip_create(lip_ctx** ctx, int device);
void lip_destroy(lip_ctx* ctx);
const char* lip_last_error(const lip_ctx* ctx);   /* ctx may be NULL: start-up error */
int  lip_set_params(lip_ctx* ctx, const lip_params* p);
/* Final dense map (p2p_matching_seg_icp.py:452,488, KDTree(xyz1).query): 0 = exact search on landmark-bin tiles
 * on the CUDA cores (blend_tile_kernel), 1 = exact search in a hashed uniform grid (grid_build_kernel +
 * blend_nn_kernel + retry/far passes), 2 = the tiles of mode 0 with the scores computed by the 5th-generation tensor cores
 * (blend_tile_tc_kernel: tcgen05.mma kind::tf32 on split operands, accumulators in TMEM).  Same results in every mode; call
 * before lip_batch_plan. */
int  lip_set_map_mode(lip_ctx* ctx, int mode);
/* pinned host memory for the *_host entry points.  Optional: lip_batch_run_host stages ordinary (pageable) arrays through
 * pinned buffers of its own with a few copy threads (environment LIP_STAGING_THREADS, default min(6, cores / 2); 0 = leave
 * it to the driver), which costs about a quarter more time than pinned caller buffers instead of five times more. */
int  lip_host_alloc(void** p, uint64_t bytes);
int  lip_host_free(void* p);

/* on != 0: lip_batch_run_host rejects inputs holding a NaN or an infinity with LIP_EINVAL before anything is uploaded —
 * the ValueError sklearn's check_array raises at every KDTree construction of the reference (registration_icp.py:51,
 * p2p_matching_seg_icp.py:388-391).  The check runs on the copy threads.  Off by default. */
int  lip_set_input_validation(lip_ctx* ctx, int on);

/* ---- batched segment-pair maps --------------------------------------------------------------
 * Replaces, for S matched segment pairs at once, the pair of calls
 *   T21 = get_p2p_mapping_from_segments_using_icp(mesh1, mesh2, lm1, lm2, vs)
 *   T12 = get_p2p_mapping_from_segments_using_icp(mesh2, mesh1, lm2, lm1, vs)
 * (p2p_matching_in_organ/p2p_matching_seg_icp.py:383-489, call site :173-176), including
 * my_icp / get_best_fit_transform / get_NN_in_target (registration/registration_icp.py:8-98) and
 * the sklearn KDTree queries they make.
 *
 * lip_batch_plan: CSR description of the batch (host arrays, copied).  Planning the shape that is already planned
 *   (same offsets and directions) returns at once, so a caller may plan before every run.
 *   off1, off2 : (S+1) vertex offsets of mesh1 / mesh2 of every pair;  lm_off : (S+1) landmark
 *   offsets (lm1 and lm2 of a pair have the same count, p2p_matching_seg_icp.py:408-416).
 * lip_batch_run_*:
 *   xyz1 (off1[S],3), xyz2 (off2[S],3), lm1, lm2 (lm_off[S],3) float64.
 *   T21 (off2[S]) : T21[off2[s]+i] = index in mesh1 of pair s matched to vertex i of mesh2;
 *   T12 (off1[S]) likewise.  Either may be NULL when its direction is not requested.
 * lip_batch_run_host on a large batch (>= 400k vertices) runs as a pipeline of up to eight sub-batches on their own
 * streams, so that the copies in, the kernels and the copies out overlap (pinned host buffers make the copies
 * truly asynchronous: lip_host_alloc).  Results are identical to an unsplit run. */
int lip_batch_plan(lip_ctx* ctx, int64_t S, const int64_t* off1, const int64_t* off2,
                   const int64_t* lm_off, int directions);
int lip_batch_run_host(lip_ctx* ctx, const double* xyz1, const double* xyz2,
                       const double* lm1, const double* lm2, void* T21, void* T12);
int lip_batch_run_device(lip_ctx* ctx, const double* d_xyz1, const double* d_xyz2,
                         const double* d_lm1, const double* d_lm2,
                         void* d_T21, void* d_T12, void* stream);
/* Entry type of the dense maps written by lip_batch_run_* and lip_series_run_*: 64 (default: int64, numpy's intp, what the
 * reference returns from KDTree.query, p2p_matching_seg_icp.py:488), 32 (int32) or 16 (uint16: the entries are indices into ONE
 * segment's mesh, so 16 bits hold them whenever every mesh of the batch has fewer than 65,536 vertices — checked by
 * lip_batch_plan; a quarter of the device-to-host bytes for callers of the batched API).  The refinement / score /
 * interpolation entry points always take int64 maps. */
int lip_set_map_dtype(lip_ctx* ctx, int bits);
/* A time series — the loop over consecutive day pairs of the reference (run_registration_pipeline.py:64,100-104: frame k is
 * day2 of frame pair k-1 and day1 of frame pair k).  `frames` (n_rows,3) holds every mesh ONCE; row1[s] / row2[s] (S entries
 * each, host arrays) give the first row of mesh1 / mesh2 of segment pair s of the planned batch in it, so a mesh can serve any
 * number of segment pairs.  Sizes, landmarks and the layout of the outputs are those of lip_batch_plan.  Same results as
 * lip_batch_run_host on the gathered arrays, with one upload of every mesh instead of one per use.
 * lip_series_run_device: `d_frames`, landmarks and maps on the device, enqueued on `stream`, no synchronisation. */
int lip_series_run_host(lip_ctx* ctx, const double* frames, int64_t n_rows, const int64_t* row1, const int64_t* row2,
                        const double* lm1, const double* lm2, void* T21, void* T12);
int lip_series_run_device(lip_ctx* ctx, const double* d_frames, int64_t n_rows, const int64_t* row1, const int64_t* row2,
                          const double* d_lm1, const double* d_lm2, void* d_T21, void* d_T12, void* stream);

/* Per-landmark results of the last run, copied to host (any pointer may be NULL).
 * Row b of the "21" arrays is landmark b of the flattened lm arrays for the mesh2 -> mesh1 ICPs
 * (src = mesh2 vertices nearest lm2[b], tgt = mesh1 vertices nearest lm1[b]); "12" is the swap.
 *   R (sumK,3,3), t (sumK,3): my_icp's R, t (x -> R x + t) — AFTER the fill of
 *   p2p_matching_seg_icp.py:422-437 if filled != 0, raw (garbage where status==0) otherwise;
 *   iters: NN passes; status: 0 = no ICP (empty src or tgt partition), 1 = ran,
 *   2 = ran and a Kabsch solve was rank-deficient (R is not unique, SURVEY.md §7 hard part 2). */
int lip_batch_get_transforms(lip_ctx* ctx, int filled,
                             double* R21, double* t21, int32_t* iters21, int32_t* status21,
                             double* R12, double* t12, int32_t* iters12, int32_t* status12);
/* Landmark assignment of the last run: two nearest landmarks (local index in the pair's lm array)
 * and their distances for every vertex: KDTree(lm).query(xyz, k=2) of p2p_matching_seg_icp.py:388-392.
 * nn (n,2) int32, dist (n,2) float64; side 1 -> mesh1 vertices (off1 order), 2 -> mesh2. */
int lip_batch_get_landmark_nn(lip_ctx* ctx, int side, int32_t* nn, double* dist);
/* counters of the last run: [0] warp-per-partition ICPs (icp_group_kernel), [1] CTA-per-partition ICPs, [2] float64
 * re-evaluations of near-tied neighbours in ICP, [3] same in the final map, [4] final-map queries
 * left open by the warp-cooperative pass (retry_nn_kernel: ring 1 + ring 2 per query), [5] that were handed to the whole-mesh scan (far_nn_kernel), [6] kernels launched by the last run, [7] candidate pairs evaluated in
 * ring 1 of the final map (only counted while profiling is on) */
int lip_batch_get_counters(lip_ctx* ctx, int64_t counters[8]);

/* (source, candidate) pairs actually evaluated by the last run — the work figure of the FP32 roofline (8 flop per
 * pair): [0] icp_group_kernel, [1] icp_big_kernel, [2] final map.  Counted only while profiling is on. */
int lip_batch_get_pair_counts(lip_ctx* ctx, int64_t pairs[3]);

/* Per-kernel device times of the last run, measured with CUDA events on the stream the kernels were
 * launched on.  Enable with lip_set_profiling before the run; while it is on, the kernels run back to
 * back on one stream (normally the ICP kernels and the grid build run as concurrent branches).
 * ms[0] landmark_partition_kernel, [1] icp_schedule_kernel, [2] icp_group_kernel (warp per partition),
 * [3] icp_big_kernel (CTA per partition), [4] grid_build_kernel (+ its table memset; 0 in map mode 0),
 * [5] blend_kernel (map mode 0), [6] blend_tile_kernel (map mode 0) / blend_nn_kernel (mode 1) / nn_query_kernel,
 * [7] retry_nn_kernel, [8] far_nn_kernel (both 0 in map mode 0). */
#define LIP_N_STAGES 9
int lip_set_profiling(lip_ctx* ctx, int on);
int lip_batch_get_stage_ms(lip_ctx* ctx, float ms[LIP_N_STAGES]);
/* (2*sumK + 1) partition offsets of the last run: bin b = landmark b of [lm1 | lm2]; the vertices of
 * its mesh whose nearest landmark it is number bin_off[b+1] - bin_off[b]. */
int lip_batch_get_bin_offsets(lip_ctx* ctx, int32_t* bin_off);

/* ---- scores of the maps (the "score outputs" the driver returns, p2p_matching_seg_icp.py:184-189) ----------------
 * evaluate_p2p_matching (p2p_matching_in_organ/p2p_matching_evaluation.py:8-88, with get_mesh_landmarks of
 * landmark_utils.py:5-9) for every segment pair of the planned batch and the caller's maps T21 (off2[S]) / T12
 * (off1[S], may be NULL):  scores (S,3) = continuity, landmark precision, cycle consistency (NaN without T12), each a
 * ratio of integer counts; counts (S,6), optional = landmark pairs within 2*threshold, landmarks, continuity passes,
 * continuity samples, cycle passes, vertices of mesh2.  threshold = 2 * voxel_size (:32). */
int lip_batch_evaluate_host(lip_ctx* ctx, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                            const int64_t* T21, const int64_t* T12, double voxel_size, double sample_proportion,
                            double* scores, int64_t* counts);

/* The same on device-resident arrays (maps produced by lip_batch_run_device stay where they are): the integer counts (S,6)
 * are written to d_counts on `stream`, no synchronisation; scores = counts[2]/counts[3], counts[0]/counts[1], counts[4]/counts[5].
 * The maps are NOT range-checked (they come from this library).  row1 / row2 (host, S entries) non-NULL: series layout as in
 * lip_series_run_device — d_xyz1 holds every mesh once (n_rows rows), d_xyz2 is ignored. */
int lip_batch_evaluate_device(lip_ctx* ctx, const double* d_xyz1, const double* d_xyz2, const double* d_lm1, const double* d_lm2,
                              const int64_t* d_T21, const int64_t* d_T12, double voxel_size, double sample_proportion,
                              int64_t* d_counts, const int64_t* row1, const int64_t* row2, int64_t n_rows, void* stream);

/* double_direction_refinement (p2p_matching_in_organ/p2p_matching_fm.py:75-132, called right after the hot path at
 * p2p_matching_seg_icp.py:181-182) for every segment pair of the planned batch: T21 (off2[S]) is refined IN PLACE with
 * the reference's visiting order (it makes the map closer to injective using the cycle error |v2[i] - v2[T12[T21[i]]]|
 * and the 4 nearest neighbours of each target); T12 (off1[S]) is read only, as in the reference.  Every mesh1 needs at
 * least 5 vertices (KDTree.query(k=5), :79).  Map entries are segment-local indices; they are range-checked on the
 * device before anything dereferences them (LIP_EINVAL; the reference's numpy raises IndexError), also in
 * lip_batch_evaluate_host. */
int lip_batch_refine_host(lip_ctx* ctx, const double* xyz1, const double* xyz2, int64_t* T21, const int64_t* T12);

/* The whole per-segment body of get_p2p_mapping_organ_collection (p2p_matching_seg_icp.py:173-189) for the planned batch
 * in one call, everything between the upload of the inputs and the download of the results staying on the device:
 * T21 / T12 (:173-176), double_direction_refinement if `refine` != 0 (:181-182), evaluate_p2p_matching (:184-185).
 * T21 (off2[S]), T12 (off1[S]), scores (S,3) as in lip_batch_evaluate_host.  The batch must be planned LIP_DIR_BOTH. */
int lip_batch_match_host(lip_ctx* ctx, const double* xyz1, const double* xyz2, const double* lm1, const double* lm2,
                         int refine, double voxel_size, double sample_proportion, int64_t* T21, int64_t* T12, double* scores);

/* interpolation_segment_level (interpolation/interpolation_pcd.py:122-134, the consumer of the saved maps:
 * interpolation_plant_level :71-76) for every segment pair of the planned batch.  out holds, per segment pair s and
 * starting at row off1[s] + off2[s], the n2 rows (1 - alpha) * xyz1[T21] + alpha * xyz2 followed by the n1 rows
 * (1 - alpha) * xyz1 + alpha * xyz2[T12] — the reference's np.vstack([xyz_inter_0, xyz_inter_1]); (off1[S] + off2[S], 3)
 * float64 in all.  Products and sum are rounded separately, as numpy does: the rows equal the reference's bit for bit.
 * Map entries are segment-local; a negative entry counts from the end (numpy indexing), anything else out of range is
 * LIP_EINVAL.  lip_batch_interpolate_device: the same on device-resident arrays, enqueued on `stream` (a cudaStream_t),
 * no synchronisation; rows whose map entry is out of range are left unwritten. */
int lip_batch_interpolate_host(lip_ctx* ctx, const double* xyz1, const double* xyz2, const int64_t* T21, const int64_t* T12,
                               double alpha, double* out);
int lip_batch_interpolate_device(lip_ctx* ctx, const double* d_xyz1, const double* d_xyz2, const int64_t* d_T21,
                                 const int64_t* d_T12, double alpha, double* d_out, void* stream);

/* ---- the three registration_icp.py functions on their own ----------------------------------- */
/* my_icp for P independent (src, tgt) problems in CSR form (registration_icp.py:60-98):
 * src (src_off[P],3), tgt (tgt_off[P],3) -> R (P,3,3), t (P,3) with x -> R x + t, iters (P). */
int lip_icp_host(lip_ctx* ctx, int64_t P, const double* src, const int64_t* src_off,
                 const double* tgt, const int64_t* tgt_off,
                 double* R, double* t, int32_t* iters, int32_t* status);
/* get_NN_in_target (registration_icp.py:49-57): for each of n src points the exact Euclidean
 * nearest of m tgt points: dist (n) float64 (sqrt applied), idx (n) int64. */
int lip_nn_host(lip_ctx* ctx, const double* src, int64_t n, const double* tgt, int64_t m,
                double* dist, int64_t* idx);
/* get_best_fit_transform (registration_icp.py:8-46): Kabsch fit of n paired points A -> B:
 * R (3,3), t (3) with B ~ R A + t. */
int lip_kabsch_host(lip_ctx* ctx, const double* A, const double* B, int64_t n,
                    double* R, double* t, int32_t* status);
/* get_triangle_angles (p2p_matching_seg_icp.py:27-34) for n triangles: abc (n,3) side lengths ->
 * angles (n,3) in radians, opposite a, b, c. */
int lip_triangle_angles_host(lip_ctx* ctx, const double* abc, int64_t n, double* angles);

/* library/ABI version and the SM architecture the kernels were compiled for (100) */
int lip_version(void);
int lip_compiled_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* LOCAL_ICP_H */
