// TEST INFRASTRUCTURE — host (g++) build of the instrumented twin (see ref_twin_common.h).
// The reference's solver / cheirality sources are all `__host__ __device__`; with both macros
// defined empty they compile as plain C++.  kernel_functions.cu itself needs curand, so the
// Sampson error is NOT taken from the reference here: ref_score() in this build forwards to the
// oracle restatement (oracle/tv5_oracle.c, tv5o_sampson_err) and is only a convenience.
#include <string.h>
#include "common.h"                 // reference
#include "polydet.cu"               // reference
#include "sturm.cu"                 // reference
#include "polyquotient.cu"          // reference
#include "cheirality.cu"            // reference
#include "essential_matrix_5pt.cu"  // reference
#include "ref_twin_common.h"

extern "C" int ref_solve_sets(const double* x1, const double* x2, int N, const int32_t* sets, int H,
                              double* E_all, int32_t* n_roots, double* E_valid, double* P_valid,
                              int32_t* n_valid) {
  (void)N;
  for (int h = 0; h < H; ++h) {
    Matches_n<5> q, qp;
    for (int i = 0; i < 5; ++i) {
      int idx = sets[h * 5 + i];
      q[i][0] = x1[2 * idx]; q[i][1] = x1[2 * idx + 1]; q[i][2] = 1.0;
      qp[i][0] = x2[2 * idx]; qp[i][1] = x2[2 * idx + 1]; qp[i][2] = 1.0;
    }
    Ematrix Es[10];
    memset(Es, 0, sizeof(Es));
    int nr = 0;
    compute_E_matrices_optimized(q, qp, Es, nr);
    n_roots[h] = nr;
    memset(&E_all[(size_t)h * 90], 0, 90 * sizeof(double));
    memcpy(&E_all[(size_t)h * 90], Es, (size_t)nr * 9 * sizeof(double));
    Pmatrix Ps[10];
    memset(Ps, 0, sizeof(Ps));
    int nv = nr;
    compute_P_matrices(q, qp, Es, (double*)0, Ps, nv, 5);
    n_valid[h] = nv;
    memset(&E_valid[(size_t)h * 90], 0, 90 * sizeof(double));
    memset(&P_valid[(size_t)h * 120], 0, 120 * sizeof(double));
    memcpy(&E_valid[(size_t)h * 90], Es, (size_t)nv * 9 * sizeof(double));
    memcpy(&P_valid[(size_t)h * 120], Ps, (size_t)nv * 12 * sizeof(double));
  }
  return 0;
}
