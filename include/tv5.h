/*
 * tv5.h — C ABI of libtv5: a B200 (sm_100a) two-view relative-pose engine.
 *
 * This is the drop-in boundary for the geometric core of Deep-SfM-Revisited's SFMnet.  Every
 * entry point names the reference interface it replaces (paths relative to
 * /root/reference/RANSAC_FiveP/essential_matrix/).  Plain pointers and sizes only; no torch
 * types.  All pointers marked "device" are CUDA device pointers on the context's device; all
 * work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 * and is asynchronous unless stated otherwise.  Functions return 0 on success or a negative
 * tv5 error code (tv5_strerror); they never print and never call exit() — unlike the
 * reference's CudaErrorCheck (essential_matrix.cu:17-24).
 *
 * Conventions (identical to the reference):
 *   x1, x2     [N,2] row-major float64, K^-1-normalised image coordinates; x2^T E x1 = 0
 *   E          3x3 row-major, unnormalised (||E||_F = sqrt(w^2+x^2+y^2+1)), arbitrary sign
 *   P          3x4 row-major [R | t], X2 = R X1 + t, ||t|| = 1, chosen by the unanimous
 *              cheirality vote of the 5 sample points (cheirality.cu:4-214)
 *   inlier     iff Sampson distance |x2'Ex1| / sqrt((Ex1)_0^2+(Ex1)_1^2+(E'x2)_0^2+(E'x2)_1^2)
 *              <= thr in float64, evaluated with the reference's exact operation order
 *              (kernel_functions.cu:231-264) — decisions are bit-identical to the reference.
 *   sets       [H,5] int32 minimal-set index table, hypothesis id h = thread*iters + it
 *              (thread in [0,512)), which reproduces the reference's tie-break order
 *              "first maximum over (thread, iteration, root)" (kernel_functions.cu:198,215;
 *              essential_matrix.cu:252).
 */
#ifndef TV5_H_
#define TV5_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TV5_VERSION 110 /* 0.1.1: winner record / pick, debug guard entry points, N-bucketed graphs */

/* error codes */
#define TV5_OK 0
#define TV5_ERR_INVALID (-1)   /* bad argument */
#define TV5_ERR_CUDA (-2)      /* CUDA runtime error (see tv5_last_cuda_error) */
#define TV5_ERR_NOMEM (-3)     /* workspace allocation failed */
#define TV5_ERR_NO_DEVICE (-4) /* no sm_100 device / device index out of range */

#define TV5_REF_THREADS 512 /* the reference's fixed 8 x 64 launch, essential_matrix.cu:201-203 */
#define TV5_MAX_SOLUTIONS 10

typedef struct tv5_ctx tv5_ctx;

/* Per-device context: owns the workspace (grown on demand, never shrunk) and the cached
 * reference-RNG index tables.  No global state — replaces the reference's module-level
 * __constant__ parameters (kernel_functions.cu:16-20) and its per-call cudaMallocManaged /
 * cudaFree (essential_matrix.cu:222-274).  Every entry point holds the context's mutex for the
 * duration of the call, so host threads sharing a context are serialised (never racing); different
 * contexts are independent.  One context owns one workspace: submissions made on
 * different streams are ordered inside the library (the later one first waits for the completion
 * event of the earlier one), so they never overlap on the device; use one context per stream for
 * concurrency.  If growing the workspace fails (TV5_ERR_NOMEM) the context stays usable: the next
 * call reallocates every buffer. */
int tv5_create(int device, tv5_ctx** out);
int tv5_destroy(tv5_ctx* ctx);
const char* tv5_strerror(int code);
int tv5_last_cuda_error(const tv5_ctx* ctx); /* cudaError_t of the last TV5_ERR_CUDA */
int tv5_version(void);
int tv5_device_sm_count(const tv5_ctx* ctx);

/* Result record written by the pose entry points (device, 8 x int32 per pair). */
typedef struct tv5_result {
  int32_t count;        /* inliers of the winner on the first n_full points (0 if none) */
  int32_t best_set;     /* winning hypothesis id h in [0,H), -1 if no hypothesis scored > 0 */
  int32_t best_root;    /* index of the winner inside its set's (cheirality-compacted) list */
  int32_t n_hypotheses; /* sum over sets of the number of scored solutions (M) */
  int32_t n_candidates; /* hypotheses re-scored exactly in float64 by the guard-band pass */
  int32_t fast_path;    /* 1 = float32 guard-band scorer used, 0 = all-float64 scorer */
  int32_t reserved[2];
} tv5_result;

/*
 * tv5_compute_pose — replaces ProjectionMatrixRansac (essential_matrix.cu:190-280, Python
 * `essential_matrix.computeP`) when with_cheirality != 0 and EssentialMatrixInitialise
 * (essential_matrix.cu:110-184, `essential_matrix.initialise`) when with_cheirality == 0.
 *
 *   x1, x2        device [N,2] float64
 *   sets          device [H,5] int32, H = TV5_REF_THREADS * iters; NULL => the table the
 *                 reference would draw (curand XORWOW, seed 1234, kernel_functions.cu:45-48,
 *                 269-300), generated on the device and cached per (N, iters)
 *   iters         num_ransac_iterations (minimal sets per reference thread)
 *   n_pre         num_test_points: points used to pick the best root inside a set
 *   n_full        num_ransac_test_points: points used to rank sets
 *   thr           inlier threshold on the unsquared Sampson distance
 *   E_out         device [9] float64
 *   P_out         device [12] float64 (zeros when with_cheirality == 0); may be NULL
 *   result        device tv5_result
 *   mask_out      device [n_full] uint8 inlier mask of the winner, or NULL  (extension: the
 *                 reference returns no mask)
 * Defined divergences from the reference's undefined behaviour: sets with
 * no surviving solution score nothing; indices are clamped to N-1; if no hypothesis has a
 * non-zero count E and P are zero and best_set = -1.
 */
int tv5_compute_pose(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                     const int32_t* sets, int iters, int n_pre, int n_full, double thr,
                     int with_cheirality, double* E_out, double* P_out, tv5_result* result,
                     uint8_t* mask_out);

/*
 * tv5_compute_pose_batch — B independent pairs in one stream-ordered submission (the
 * reference loops over pairs in Python, models/SFMnet.py:229-272).
 *   x1, x2        device [sum N_b, 2] float64, pairs concatenated
 *   pt_offsets    HOST [B+1] int64 prefix sums of N_b
 *   sets          device [B,H,5] int32 with indices local to each pair, or NULL (reference RNG
 *                 table per pair; all pairs with the same N share one table, as B successive
 *                 reference calls would)
 *   n_pre/n_full  <= 0 means "all N_b points of the pair" (what SFMnet passes)
 *   E_out [B,9], P_out [B,12] (or NULL), result [B], mask_out [sum N_b] or NULL: device
 */
int tv5_compute_pose_batch(tv5_ctx* ctx, void* stream, int B, const double* x1, const double* x2,
                           const int64_t* pt_offsets, const int32_t* sets, int iters, int n_pre,
                           int n_full, double thr, int with_cheirality, double* E_out,
                           double* P_out, tv5_result* result, uint8_t* mask_out);

/*
 * Host-buffer convenience used by non-torch callers and by bench.py's end-to-end leg: copies
 * x1/x2 (and sets, if given) from HOST memory, runs tv5_compute_pose_batch, copies E, P and
 * result back and synchronises the stream.  All pointers are HOST pointers.
 */
int tv5_compute_pose_batch_host(tv5_ctx* ctx, void* stream, int B, const double* x1,
                                const double* x2, const int64_t* pt_offsets, const int32_t* sets,
                                int iters, int n_pre, int n_full, double thr, int with_cheirality,
                                double* E_out, double* P_out, tv5_result* result);

/*
 * tv5_solve5 — five-point solver + cheirality for H minimal sets; replaces
 * compute_E_matrices_optimized (essential_matrix_5pt.cu:1224-1249) followed by
 * compute_P_matrices (cheirality.cu:4-214) as called at kernel_functions.cu:163,180.
 *   E_list  device [H,10,9]  solutions in ascending order of the hidden variable w; when
 *                            with_cheirality != 0 the list is compacted to the valid ones
 *   P_list  device [H,10,12] or NULL
 *   n_roots device [H] or NULL (number of real roots)
 *   n_valid device [H]       (number of entries in E_list / P_list)
 */
int tv5_solve5(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
               const int32_t* sets, int H, int with_cheirality, double* E_list, double* P_list,
               int32_t* n_roots, int32_t* n_valid);

/*
 * tv5_score — exact float64 Sampson inlier counts of M essential matrices on the first n_test
 * points; replaces the scoring loops + ComputeError<double> (kernel_functions.cu:184-214,
 * 231-264).  counts device [M] int32; masks device [M, ceil(n_test/32)] uint32 bit-packed
 * (bit k%32 of word k/32 = point k is an inlier) or NULL.
 */
int tv5_score(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int n_test,
              const double* E_list, int M, double thr, int32_t* counts, uint32_t* masks);

/*
 * tv5_score_bounds — the float32 guard-band scorer used by the fast path, exposed for tests
 * and roofline measurement: lo[m] <= exact count[m] <= hi[m] for every m (lo counts the
 * evaluations that are inliers under every rounding, hi adds the undecidable ones).
 * lo, hi device [M] int32.  Returns TV5_ERR_INVALID if the coordinates are not finite or too
 * large for the float32 bound (the pose entry points then use the float64 scorer).
 */
int tv5_score_bounds(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int n_test,
                     const double* E_list, int M, double thr, int32_t* lo, int32_t* hi);

/*
 * tv5_ref_rng_sets — the index table the reference draws for (N, iters): 512 curand XORWOW
 * states curand_init(1234, tid, 0), 5 x curand_uniform per set, index =
 * trunc(u * (N - 1 + 0.999999f)) in float32 (kernel_functions.cu:45-48, 269-300), clamped to
 * N-1.  sets_out device [512*iters, 5] int32.
 */
int tv5_ref_rng_sets(tv5_ctx* ctx, void* stream, int N, int iters, int32_t* sets_out);

/*
 * tv5_decompose / tv5_decompose_uv — E = U diag(1,1,0) V^T by three left and two right Givens
 * rotations; replace EssentialMatrixDecompose / EssentialMatrixDecomposeUV
 * (essential_matrix.cu:29-70 -> Edecomp, polish_E.cu:147-338; Python `decompose`,
 * `decomposeUV`).  Like the reference these two take HOST pointers and run on the calling
 * thread (a 3x3 matrix); results are bit-identical to the reference's.
 *   E host [9] row-major;  angles host [5] = (x, y, z, u, v);  U, V host [9] row-major.
 * tv5_decompose_batch is the device form: E device [B,9]; angles [B,5], U [B,9], V [B,9] device,
 * any of the three may be NULL.
 */
int tv5_decompose(const double* E, double* angles);
int tv5_decompose_uv(const double* E, double* U, double* V);
int tv5_decompose_batch(tv5_ctx* ctx, void* stream, const double* E, int B, double* angles,
                        double* U, double* V);

/*
 * tv5_optimise — iteratively re-weighted Gauss-Newton refinement of E on its five rotation
 * angles; replaces EssentialMatrixOptimise (essential_matrix.cu:76-105 ->
 * polish_E_robust_parametric, polish_E.cu:1470-1577; Python `optimise`), which the reference
 * runs on one CPU core.  Residual e = (V^T x1)_0 (U^T x2)_0 + (V^T x1)_1 (U^T x2)_1, weight 1 if
 * |e| < delta else alpha*delta/|e|; stops when |J^T W e|^2 < 1e-20 or after max_reps updates
 * (max_reps is clamped to 1,000,000).  Like the reference, a call that stops before its first
 * update returns the half-reduced working matrix of the decomposition (polish_E.cu:1545).
 *   x1, x2     device [N,2] float64
 *   mask       device [N] uint8 or NULL; points with mask == 0 are ignored (extension: lets the
 *              winner's inlier mask of tv5_compute_pose drive a local-optimisation step)
 *   E_io       device [9]: initial estimate in, refined E out (||E||_F = sqrt 2 once updated)
 *   iters_out  device int32 (number of updates applied) or NULL
 * One cooperative kernel per call; the per-iteration sums are reduced in a fixed order, so a call
 * is reproducible, but the order differs from the reference's sequential loop: results agree to
 * rounding (tests: 1e-9), not bit for bit.
 * tv5_optimise_batch: B independent problems, points concatenated, pt_offsets HOST [B+1],
 * mask device [sum N] or NULL, E_io device [B,9], iters_out device [B] or NULL.
 * tv5_optimise_host: all pointers HOST (what the reference's Python passes); copies, runs,
 * copies E back and synchronises.
 */
int tv5_optimise(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                 const uint8_t* mask, double* E_io, double delta, double alpha, int max_reps,
                 int32_t* iters_out);
int tv5_optimise_batch(tv5_ctx* ctx, void* stream, int B, const double* x1, const double* x2,
                       const int64_t* pt_offsets, const uint8_t* mask, double* E_io, double delta,
                       double alpha, int max_reps, int32_t* iters_out);
int tv5_optimise_host(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                      const double* E_init, double delta, double alpha, int max_reps, double* E_out);

/*
 * tv5_flow_to_points — optical flow -> K^-1-normalised correspondences in one pass; replaces the
 * per-image tensor chain in front of every pose solve in the reference: flow2coord
 * (models/SFMnet.py:298-318), the point selection of pose_by_ransac (:239-254), bmm(K^-1, .)
 * (:259-260), transpose/[:, :2]/contiguous (:262-263) and .double() (epipolar_utils.py:130).
 *   flow        device [B,2,H,W] float32 (channel 0 = dx, 1 = dy)
 *   Kinv        device [B,3,3] float32 inverse intrinsics
 *   mode        0 dense crop [margin,H-margin) x [margin,W-margin) row-major (SFMnet.py:240-241)
 *               1 integer pixel list: pts device int32 [sum n,2] = (x, y) (SFMnet.py:251-254)
 *               2 sub-pixel list: pts device float32 [sum n,2], bilinear, align_corners=True,
 *                 zero padding (cfg.SAMPLE_SP, SFMnet.py:244-249)
 *   pt_offsets  HOST [B+1] prefix sums of the list lengths (modes 1, 2; ignored for mode 0)
 *   x1_out, x2_out  device [sum n,2] float64 — exactly the arrays the reference passes to
 *               `essential_matrix.computeP`: all arithmetic in float32 as in the reference,
 *               widened at the end.
 * tv5_pose_from_flow — the same followed by tv5_compute_pose_batch on all points of each image
 * (n_pre = n_full = n, what SFMnet passes) in one stream-ordered submission without
 * intermediate tensors; E32_out [B,9] / P32_out [B,12] are the float32 E_mat / P_mat tensors
 * pose_by_ransac returns (SFMnet.py:186-187,272); E_out / P_out (float64, optional) the
 * unrounded ones.
 */
int tv5_flow_to_points(tv5_ctx* ctx, void* stream, const float* flow, int B, int H, int W,
                       const float* Kinv, int mode, int margin, const void* pts,
                       const int64_t* pt_offsets, double* x1_out, double* x2_out);
int tv5_pose_from_flow(tv5_ctx* ctx, void* stream, const float* flow, int B, int H, int W,
                       const float* Kinv, int mode, int margin, const void* pts,
                       const int64_t* pt_offsets, const int32_t* sets, int iters, double thr,
                       int with_cheirality, float* E32_out, float* P32_out, tv5_result* result,
                       double* E_out, double* P_out);

/*
 * tv5_plane_sweep — plane-sweep cost volume from the estimated pose in one launch; replaces the
 * label loop of PSNet.forward (models/PSNet.py:141-157) over inverse_warp
 * (models/inverse_warp.py:121-153), the direct consumer of P.
 *   ref_feat, tgt_feat  device [B,C,h,w] float32 (quarter-resolution features, C = 32)
 *   pose                device [B,3,4] float32 [R|t] "from ref to target" (P32_out of
 *                       tv5_pose_from_flow / the P_mat of pose_by_ransac)
 *   K, Kinv             device [B,3,3] float32 intrinsics at FEATURE resolution
 *                       (PSNet.py:130-133: K[:2,:] / 4, Kinv[:2,:2] * 4)
 *   nlabel, mindepth    depth of plane i = mindepth*nlabel/(i+1), or (i+1)*mindepth if by_depth
 *                       (cfg.PREDICT_BY_DEPTH)
 *   cost                device [B,2C,nlabel,h,w] float32: channels [0,C) = ref_feat on every plane,
 *                       [C,2C) = tgt_feat sampled bilinearly (zeros outside, align_corners=True)
 * float32 like the reference; agreement with torch's own kernels is to float32 rounding of the
 * sample position (tests state the bound).
 */
int tv5_plane_sweep(tv5_ctx* ctx, void* stream, const float* ref_feat, const float* tgt_feat,
                    const float* pose, const float* K, const float* Kinv, int B, int C, int h, int w,
                    int nlabel, float mindepth, int by_depth, float* cost);

/*
 * Hypothesis-sharded single pair (SURVEY.md section 8(e), configs[3]): every GPU solves and scores
 * its share of the minimal sets of ONE pair (tv5_compute_pose with the rows [set_offset,
 * set_offset + H_local) of the index table); the global winner is the reference's first maximum
 * over (thread, iteration, root) — its host std::max_element over the 512 per-thread results,
 * essential_matrix.cu:252 — taken across GPUs without a host round trip:
 *   tv5_winner_record  packs one rank's result into a 192-byte record (device):
 *                      u64 key = count << 32 | ~(global_set * 16 + root)  (0 when count == 0),
 *                      E[9], P[12] float64, n_hypotheses, n_candidates, fast_path, pad (int32)
 *   (the caller all-gathers the G records, e.g. ncclAllGather over NVLink)
 *   tv5_winner_pick    maximum key over G records -> E_out[9], P_out[12] (may be NULL), result
 *                      (count, GLOBAL best_set, best_root, summed n_hypotheses / n_candidates).
 * Both are stream-ordered single-CTA kernels; TV5_WINNER_RECORD_BYTES = 192.
 */
#define TV5_WINNER_RECORD_BYTES 192
int tv5_winner_record(tv5_ctx* ctx, void* stream, const double* E, const double* P,
                      const tv5_result* result, int set_offset, void* record_out);
int tv5_winner_pick(tv5_ctx* ctx, void* stream, const void* records, int G, double* E_out,
                    double* P_out, tv5_result* result_out);

/*
 * Testing aid (stands in for compute-sanitizer's memcheck / initcheck, which the B200 pool keeps
 * closed).  tv5_debug_guard must be called on a fresh context, before its first submission: from
 * then on every workspace buffer is allocated with a 256-byte guard zone on either side and its
 * payload filled with `poison_byte`.  tv5_debug_poison refills every payload (synchronises);
 * tv5_debug_check_guards synchronises and counts guard bytes that were overwritten — 0 means no
 * kernel wrote outside its buffers.  Results must not depend on the poison (tests run 0x00 / 0xFF /
 * 0x5A and compare bit for bit): no kernel reads workspace that this submission did not write.
 */
int tv5_debug_guard(tv5_ctx* ctx, int on, int poison_byte);
int tv5_debug_poison(tv5_ctx* ctx, int poison_byte);
int tv5_debug_check_guards(tv5_ctx* ctx, int64_t* corrupted_bytes_out, int32_t* n_buffers_out);
/* Self-test of the detector: writes one byte just outside the payload of the first guarded buffer
 * (back != 0: behind it, else in front of it); the next tv5_debug_check_guards must report it. */
int tv5_debug_stray_write(tv5_ctx* ctx, int back);

/* Testing/diagnostic switch: on != 0 makes the pose entry points score every hypothesis with the
 * float64 scorer (no float32 guard-band pass).  Results are identical by construction; the
 * tests use this to prove it. */
int tv5_set_force_exact(tv5_ctx* ctx, int on);

/* CUDA graphs (default on): a single-pair submission (tv5_compute_pose, the shape SFMnet calls) replays a
 * graph captured once per (N, iterations, flags) instead of issuing its memset and 13 kernel launches one
 * by one; on = 0 issues plain launches. */
int tv5_set_graphs(tv5_ctx* ctx, int on);

/* Early exit (opt-in, default off): the correspondences are scored in three stages (30 %, 52 %, 100 % of
 * each pair's points); after a stage every hypothesis whose upper bound on its FULL inlier count —
 * unseen points counted as inliers — is below the exact count of an actual hypothesis is dropped.
 * The winner, its count, E, P and mask are identical to scoring everything (tested); about 2/3 of
 * the Sampson evaluations are never made on RANSAC-typical data.  Off, every hypothesis is scored
 * against every correspondence, which is what the headline benchmark measures. */
int tv5_set_early_exit(tv5_ctx* ctx, int on);

/* Five-point solver organisation: on != 0 (default) = three kernels (front: null space + elimination +
 * determinant per set; roots: Sturm isolation per set; poses: Newton + E + cheirality per ROOT), on = 0 =
 * the fused one-kernel form.  Results are identical (tested); only the speed differs. */
int tv5_set_split_solver(tv5_ctx* ctx, int on);

/* Solver/scorer overlap (experimental, default off): a submission of >= 32 pairs is cut into up
 * to 8 chunks whose five-point solve runs on a low-priority internal stream concurrently with the
 * scoring of the previous chunk on a high-priority one; results are identical either way.
 * Measured on B200: no throughput gain. */
int tv5_set_overlap(tv5_ctx* ctx, int on);

/* Device-timed (CUDA events) measurements used by bench.py; both synchronise. */

/* FP32 FMA-chain peak of this GPU in TFLOP/s: mode 0 = scalar FFMA, 1 = packed FFMA2. */
int tv5_measure_fp32_peak(tv5_ctx* ctx, int mode, double* tflops_out);

/* Average duration in milliseconds of each pipeline stage over the launches since the last
 * reset (CUDA events recorded on the launching stream when profiling is enabled).
 * stage ids: 0 prep, 1 solve, 2 plan, 3 score_bounds, 4 candidates+exact, 5 finalize. */
#define TV5_N_STAGES 6
int tv5_profile_enable(tv5_ctx* ctx, int on);
int tv5_profile_read(tv5_ctx* ctx, double ms_out[TV5_N_STAGES], int64_t launches_out[TV5_N_STAGES],
                     int reset);

#ifdef __cplusplus
}
#endif
#endif /* TV5_H_ */
