// TEST INFRASTRUCTURE — not part of the product.
//
// "Instrumented twin" of the reference RANSAC_FiveP path.  This file contains NO reference
// code: it #includes the reference's own sources where they lie under /root/reference (the
// include path is supplied by oracle/build_ref.sh) and only adds thin entry points that expose
// intermediate results the stock extension does not return (sampled index tables, per-set
// E / P lists, per-hypothesis Sampson inlier counts).
//
// The same file is compiled twice:
//   * by nvcc  (REF_TWIN_CUDA defined)  -> oracle/_ref/libref_twin_cuda.so   (runs on the GPU box)
//   * by g++   (-D__host__= -D__device__=) -> oracle/_ref/libref_host.so     (runs anywhere)
//
// Reference entry points wrapped (file:line in /root/reference/RANSAC_FiveP/essential_matrix):
//   compute_E_matrices_optimized   essential_matrix_5pt.cu:1224
//   compute_P_matrices             cheirality.cu:4
//   ComputeError<double>           kernel_functions.cu:231   (CUDA build only; needs curand)
//   SetupRandomState / RandomInt   kernel_functions.cu:45 / :269 (CUDA build only)
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Five-point solve + cheirality for H minimal sets.  x1/x2: [N,2] row-major, sets: [H,5].
// E_all   [H,10,9]  all real-root solutions in reference order (ascending hidden variable)
// n_roots [H]
// E_valid [H,10,9]  cheirality-compacted list (what the reference scores in computeP)
// P_valid [H,10,12]
// n_valid [H]
// Arrays are zero-filled first, so the reference's stale-slot quirks show up
// as zeros instead of garbage.  Pointers are device pointers in the CUDA build.
int ref_solve_sets(const double* x1, const double* x2, int N, const int32_t* sets, int H,
                   double* E_all, int32_t* n_roots, double* E_valid, double* P_valid,
                   int32_t* n_valid);

// Sampson scoring with the reference's own ComputeError: counts[m] = #{k < n_test : err <= thr}.
// err_out ([M,n_test] or NULL) receives the raw error values.
int ref_score(const double* x1, const double* x2, int n_test, const double* E_list, int M,
              double thr, int32_t* counts, double* err_out);

#ifdef REF_TWIN_CUDA
// Index table drawn exactly as the reference does: curand_init(1234, tid, 0) per thread,
// 5 x RandomInt per iteration; out[(tid*iters + it)*5 + i].  512 threads (8 x 64).
int ref_rng_sets(int N, int iters, int32_t* out);

// The reference's whole computeP on its own kernels (see ref_twin_cuda.cu); x1/x2/E_out/P_out
// device pointers, count_out host pointer.  Synchronous.  Returns a cudaError_t.
int ref_compute_pose(const double* x1, const double* x2, int N, int n_pre, int n_full, int iters,
                     double thr, double* E_out, double* P_out, int32_t* count_out, int use_managed);
#endif

#ifdef __cplusplus
}
#endif
