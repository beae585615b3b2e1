// TEST INFRASTRUCTURE — the reference's own RANSAC kernels behind a C entry point
// (see ref_twin_common.h).  The reference translation unit is included verbatim from
// /root/reference; nothing of it is copied here.
#define REF_TWIN_CUDA 1
#include <cuda_runtime.h>
#include "common.h"              // reference
#include "kernel_functions.cu"   // reference
#include "ref_twin_common.h"

// Host flow of ProjectionMatrixRansac (reference essential_matrix.cu:190-280) around the
// reference's own, unmodified kernels SetupRandomState and EstimateProjectionMatrix<5>
// (kernel_functions.cu:45-48, 140-226), same <<<8, 64>>> launch, same __constant__ uploads, same
// host-side first-max over the 512 per-thread counts.  essential_matrix.cu itself cannot be
// included here because it needs ATen; only its ~40 host lines are restated.  use_managed != 0
// allocates the work buffers with cudaMallocManaged exactly as the reference does, otherwise
// with cudaMalloc + an explicit copy of the 512 counts.
extern "C" int ref_compute_pose(const double* x1, const double* x2, int N, int n_pre, int n_full,
                                int iters, double thr, double* E_out, double* P_out,
                                int32_t* count_out, int use_managed) {
  const int threads = 64, blocks = 8, total = threads * blocks;
  int* num_inliers = nullptr;
  double (*Es)[3][3] = nullptr;
  double (*Ps)[3][4] = nullptr;
  curandState* state = nullptr;
  cudaError_t e;
#define TWIN_CHECK(x) do { e = (x); if (e != cudaSuccess) return (int)e; } while (0)
  if (use_managed) {
    TWIN_CHECK(cudaMallocManaged((void**)&num_inliers, total * sizeof(int)));
    TWIN_CHECK(cudaMallocManaged((void**)&Es, total * 9 * sizeof(double)));
    TWIN_CHECK(cudaMallocManaged((void**)&Ps, total * 12 * sizeof(double)));
    TWIN_CHECK(cudaMallocManaged((void**)&state, total * sizeof(curandState)));
  } else {
    TWIN_CHECK(cudaMalloc((void**)&num_inliers, total * sizeof(int)));
    TWIN_CHECK(cudaMalloc((void**)&Es, total * 9 * sizeof(double)));
    TWIN_CHECK(cudaMalloc((void**)&Ps, total * 12 * sizeof(double)));
    TWIN_CHECK(cudaMalloc((void**)&state, total * sizeof(curandState)));
  }
  TWIN_CHECK(cudaMemcpyToSymbol(c_num_points, &N, sizeof(int)));
  TWIN_CHECK(cudaMemcpyToSymbol(c_num_test_points, &n_pre, sizeof(int)));
  TWIN_CHECK(cudaMemcpyToSymbol(c_ransac_num_test_points, &n_full, sizeof(int)));
  TWIN_CHECK(cudaMemcpyToSymbol(c_ransac_num_iterations, &iters, sizeof(int)));
  TWIN_CHECK(cudaMemcpyToSymbol(c_inlier_threshold, &thr, sizeof(double)));
  SetupRandomState<<<blocks, threads>>>(1234ULL, state);
  EstimateProjectionMatrix<5><<<blocks, threads>>>(x1, x2, state, num_inliers, Es, Ps);
  TWIN_CHECK(cudaPeekAtLastError());
  TWIN_CHECK(cudaDeviceSynchronize());
  int host_counts[512];
  const int* counts = num_inliers;
  if (!use_managed) {
    TWIN_CHECK(cudaMemcpy(host_counts, num_inliers, sizeof(host_counts), cudaMemcpyDeviceToHost));
    counts = host_counts;
  }
  int best = 0;
  for (int i = 1; i < total; ++i) if (counts[i] > counts[best]) best = i;  // first maximum
  *count_out = counts[best];
  TWIN_CHECK(cudaMemcpy(E_out, &Es[best], 72, cudaMemcpyDeviceToDevice));
  TWIN_CHECK(cudaMemcpy(P_out, &Ps[best], 96, cudaMemcpyDeviceToDevice));
  cudaFree(num_inliers); cudaFree(Es); cudaFree(Ps); cudaFree(state);
  return 0;
#undef TWIN_CHECK
}
