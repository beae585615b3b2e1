// TEST INFRASTRUCTURE — CUDA build of the instrumented twin (see ref_twin_common.h).
// The reference translation unit is included verbatim from /root/reference.
#define REF_TWIN_CUDA 1
#include <cuda_runtime.h>
#include "common.h"              // reference
#include "kernel_functions.cu"   // reference (pulls in sturm.cu, essential_matrix_5pt.cu, cheirality.cu ...)
#include "ref_twin_common.h"

namespace {

__global__ void twin_solve(const double* x1, const double* x2, const int32_t* sets, int H,
                           double* E_all, int32_t* n_roots, double* E_valid, double* P_valid,
                           int32_t* n_valid) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  Matches_n<5> q, qp;
  for (int i = 0; i < 5; ++i) {
    int idx = sets[h * 5 + i];
    q[i][0] = x1[2 * idx]; q[i][1] = x1[2 * idx + 1]; q[i][2] = 1.0;
    qp[i][0] = x2[2 * idx]; qp[i][1] = x2[2 * idx + 1]; qp[i][2] = 1.0;
  }
  Ematrix Es[10];
  for (int m = 0; m < 10; ++m) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Es[m][i][j] = 0.0;
  int nr = 0;
  compute_E_matrices_optimized(q, qp, Es, nr);
  n_roots[h] = nr;
  for (int m = 0; m < 10; ++m) for (int i = 0; i < 9; ++i) E_all[(h * 10 + m) * 9 + i] = (m < nr) ? (&Es[m][0][0])[i] : 0.0;
  Pmatrix Ps[10];
  for (int m = 0; m < 10; ++m) for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) Ps[m][i][j] = 0.0;
  int nv = nr;
  compute_P_matrices(q, qp, Es, (double*)0, Ps, nv, 5);
  n_valid[h] = nv;
  for (int m = 0; m < 10; ++m) {
    for (int i = 0; i < 9; ++i) E_valid[(h * 10 + m) * 9 + i] = (m < nv) ? (&Es[m][0][0])[i] : 0.0;
    for (int i = 0; i < 12; ++i) P_valid[(h * 10 + m) * 12 + i] = (m < nv) ? (&Ps[m][0][0])[i] : 0.0;
  }
}

__global__ void twin_score(const double* x1, const double* x2, int n_test, const double* E_list,
                           int M, double thr, int32_t* counts, double* err_out) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  Ematrix E;
  for (int i = 0; i < 9; ++i) (&E[0][0])[i] = E_list[m * 9 + i];
  int c = 0;
  for (int k = 0; k < n_test; ++k) {
    double error;
    double q_test[3] = {x1[2 * k], x1[2 * k + 1], 1.0};
    double qp_test[3] = {x2[2 * k], x2[2 * k + 1], 1.0};
    ComputeError<double>(q_test, qp_test, E, error);
    if (error <= thr) c++;
    if (err_out) err_out[(size_t)m * n_test + k] = error;
  }
  counts[m] = c;
}

__global__ void twin_rng(curandState* state, int N, int iters, int32_t* out) {
  int gi = threadIdx.x + blockDim.x * blockIdx.x;
  for (int it = 0; it < iters; ++it)
    for (int i = 0; i < 5; ++i)
      out[(gi * iters + it) * 5 + i] = RandomInt(state, gi, 0, N - 1);
}

}  // namespace

extern "C" int ref_solve_sets(const double* x1, const double* x2, int N, const int32_t* sets, int H,
                              double* E_all, int32_t* n_roots, double* E_valid, double* P_valid,
                              int32_t* n_valid) {
  (void)N;
  twin_solve<<<(H + 63) / 64, 64>>>(x1, x2, sets, H, E_all, n_roots, E_valid, P_valid, n_valid);
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_score(const double* x1, const double* x2, int n_test, const double* E_list, int M,
                         double thr, int32_t* counts, double* err_out) {
  twin_score<<<(M + 63) / 64, 64>>>(x1, x2, n_test, E_list, M, thr, counts, err_out);
  return (int)cudaDeviceSynchronize();
}

extern "C" int ref_rng_sets(int N, int iters, int32_t* out) {
  curandState* state = nullptr;
  cudaError_t e = cudaMalloc(&state, 512 * sizeof(curandState));
  if (e != cudaSuccess) return (int)e;
  SetupRandomState<<<8, 64>>>(1234ULL, state);
  twin_rng<<<8, 64>>>(state, N, iters, out);
  e = cudaDeviceSynchronize();
  cudaFree(state);
  return (int)e;
}
