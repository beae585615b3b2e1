/*
 * tv5_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double precision) of the reference's two-view relative-pose hot
 * path, i.e. what `essential_matrix.computeP` / `initialise` compute
 * (/root/reference/RANSAC_FiveP/essential_matrix/...).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against
 * (a) oracle/_ref/libref_host.so = the reference's own solver + cheirality sources compiled with
 *     g++ from /root/reference (when present), and
 * (b) tests/golden/ fixtures generated from that library by tests/golden/make_golden.py, and
 * (c) on the GPU box, oracle/_ref/libref_twin_cuda.so + the unmodified reference extension.
 */
#ifndef TV5_ORACLE_H_
#define TV5_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Sampson distance exactly as the reference evaluates it (kernel_functions.cu:231-264, with the
 * FMA contraction order nvcc 12.9 emits for sm_100a; see DESIGN.md "exact scoring sequence"). */
double tv5o_sampson_err(const double E[9], double x1, double y1, double x2, double y2);

/* counts[m] = #{k < n : err(E_m, point k) <= thr};  mask ([M,n] bytes) optional. */
void tv5o_score(const double* x1, const double* x2, int n, const double* E_list, int M,
                double thr, int32_t* counts, uint8_t* mask);

/* Five-point solver (compute_E_matrices_optimized, essential_matrix_5pt.cu:1224-1249).
 * q, qp: 5 x 2 normalised image points (first / second image).  E_out: up to 10 matrices,
 * row-major, unnormalised, ordered by ascending hidden variable w.  Returns #solutions.
 * w_out (optional) receives the roots. */
int tv5o_solve5(const double q[5][2], const double qp[5][2], double E_out[10][9], double* w_out);

/* Intermediate stages, exposed for tests. */
void tv5o_nullspace_basis(const double q[5][2], const double qp[5][2], double B[4][9]);
void tv5o_hidden_poly(const double q[5][2], const double qp[5][2], double poly[11]);
int tv5o_real_roots(const double* p, int degree, double* roots);

/* Cheirality filter + P extraction (compute_P_matrices, cheirality.cu:4-214) on the 5 sample
 * points.  Compacts E in place; returns the number kept. */
int tv5o_cheirality(const double q[5][2], const double qp[5][2], double E[][9], int n,
                    double P[][12]);

/* Index draw from a uniform float in (0,1] (RandomInt, kernel_functions.cu:269-278). */
int32_t tv5o_index_from_uniform(float u, int N);

/* Whole RANSAC with the reference's selection order (kernel_functions.cu:140-226 +
 * essential_matrix.cu:252): hypothesis id h = thread*iters + it over `sets` [H,5] with
 * H = n_threads*iters.  Known, documented divergences from the reference's undefined
 * behaviour: sets with no surviving solution are skipped (they do not re-score a stale slot).
 * Outputs: E[9], P[12] (zeros when with_cheirality == 0), returns best count;
 * best_set / best_root receive the winning hypothesis; mask ([n_full] bytes) optional. */
int tv5o_ransac(const double* x1, const double* x2, int N, const int32_t* sets, int n_threads,
                int iters, int n_pre, int n_full, double thr, int with_cheirality, double E[9],
                double P[12], int32_t* best_set, int32_t* best_root, uint8_t* mask);

/* Per-set dump used by golden fixtures: solutions after (optional) cheirality + counts. */
void tv5o_solve_sets(const double* x1, const double* x2, const int32_t* sets, int H,
                     int with_cheirality, double* E_list /*[H,10,9]*/, double* P_list /*[H,10,12]*/,
                     int32_t* n_roots /*[H]*/, int32_t* n_valid /*[H]*/);

/* E = U diag(1,1,0) V^T by three left and two right Givens rotations (Edecomp, polish_E.cu:147-244);
 * E is overwritten with the reduced matrix exactly as the reference does.  U, V row-major. */
void tv5o_decompose_uv(double E[9], double U[9], double V[9]);

/* The five Givens angles (x, y, z, u, v) of the same decomposition (Edecomp, polish_E.cu:246-338). */
void tv5o_decompose_angles(double E[9], double par[5]);

/* Iteratively re-weighted least-squares refinement of E on n correspondences
 * (polish_E_robust_parametric, polish_E.cu:1470-1577): residual eps = first two components of
 * (x1 V).(x2 U), weight 1 if |eps| < delta else alpha*delta/|eps|, Gauss-Newton on the 5 angles,
 * stops when |J^T W eps|^2 < 1e-20 or after max_reps updates.  E is updated in place. */
void tv5o_optimise(double E[9], const double* x1, const double* x2, int n, double delta,
                   double alpha, int max_reps);

#ifdef __cplusplus
}
#endif
#endif
