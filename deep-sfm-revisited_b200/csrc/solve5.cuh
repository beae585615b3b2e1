// solve5.cuh — five-point relative-pose minimal solver, one minimal set per thread, float64.
//
// Replaces (does not port) the reference's compute_E_matrices_optimized
// (RANSAC_FiveP/essential_matrix/essential_matrix_5pt.cu:1224-1249) and compute_P_matrices
// (cheirality.cu:4-214).  Formulation (Nister 2004, hidden-variable form):
//   1. null space of the 5x9 epipolar system  -> E(w,x,y) = w B0 + x B1 + y B2 + B3
//   2. det E = 0 and 2 E E^T E - tr(E E^T) E = 0 -> 10 cubics = a 10x20 coefficient matrix over
//      the monomials  [x3 x2y xy2 y3 x2w x2 xyw xy y2w y2 | xw2 xw x yw2 yw y w3 w2 w 1]
//   3. Gauss-Jordan with partial pivoting on the first 10 columns
//   4. rows (x2w,x2), (xyw,xy), (y2w,y2) combine to a 3x3 polynomial matrix in w whose
//      determinant is the degree-10 hidden-variable polynomial
//   5. real roots by Sturm sequence + bisection isolation + safeguarded Newton
//   6. (x,y) from the 3x3 null vector, E per root; roots ascending in w
//   7. twisted-pair decomposition in closed form (Horn 1990) and a unanimous positive-depth
//      vote of the 5 sample points  -> P = [R | t]
//
// Output compatibility with the reference: the null-space basis B0..B3 is *defined* the way the
// reference defines it (orthonormal completion of the five constraint rows by the fixed
// sequence of essential_matrix_5pt.cu:639-649 under modified Gram-Schmidt, :651-668), so the
// hidden variable w, the order of the solutions and the scale of the returned (unnormalised) E
// are the same as the reference's.  Everything after the basis is an independent derivation.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifndef TV5_ISOLATE_TOP_IN_REGS
#define TV5_ISOLATE_TOP_IN_REGS 1
#endif

namespace tv5 {

#ifdef TV5_SOLVE_PROFILE
__device__ unsigned long long g_solve_prof[16];
#endif

// Completion rows of the 9x9 system (filled by tv5_create from the recurrence
// ran <- 3.18730379 * ran; ran <- 2 (ran - floor(ran)) - 1, started at 3.18730379).
__constant__ double c_completion[4][9];

// ------------------------------------------------------------------------------------------
// sorted-index tables for homogeneous polynomials in v = (w, x, y, 1)
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int q2(int i, int j) {  // i <= j, 10 quadratic monomials
  return i * 4 - (i * (i - 1)) / 2 + (j - i);
}
// column of the sorted cubic monomial (i<=j<=k) in the 10x20 matrix; W=0 X=1 Y=2 Z=3
__host__ __device__ constexpr int col3(int i, int j, int k) {
  return (i == 1 && j == 1 && k == 1) ? 0    // x3
       : (i == 1 && j == 1 && k == 2) ? 1    // x2y
       : (i == 1 && j == 2 && k == 2) ? 2    // xy2
       : (i == 2 && j == 2 && k == 2) ? 3    // y3
       : (i == 0 && j == 1 && k == 1) ? 4    // x2w
       : (i == 1 && j == 1 && k == 3) ? 5    // x2
       : (i == 0 && j == 1 && k == 2) ? 6    // xyw
       : (i == 1 && j == 2 && k == 3) ? 7    // xy
       : (i == 0 && j == 2 && k == 2) ? 8    // y2w
       : (i == 2 && j == 2 && k == 3) ? 9    // y2
       : (i == 0 && j == 0 && k == 1) ? 10   // xw2
       : (i == 0 && j == 1 && k == 3) ? 11   // xw
       : (i == 1 && j == 3 && k == 3) ? 12   // x
       : (i == 0 && j == 0 && k == 2) ? 13   // yw2
       : (i == 0 && j == 2 && k == 3) ? 14   // yw
       : (i == 2 && j == 3 && k == 3) ? 15   // y
       : (i == 0 && j == 0 && k == 0) ? 16   // w3
       : (i == 0 && j == 0 && k == 3) ? 17   // w2
       : (i == 0 && j == 3 && k == 3) ? 18   // w
       : 19;                                 // 1
}
__host__ __device__ constexpr int col3_any(int i, int j, int k) {  // (i<=j), k anywhere
  return k < i ? col3(k, i, j) : (k < j ? col3(i, k, j) : col3(i, j, k));
}

// Q += a * b  (linear x linear -> quadratic)
__device__ __forceinline__ void quad_acc(double (&Q)[10], const double (&a)[4], const double (&b)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Q[i <= j ? q2(i, j) : q2(j, i)] += a[i] * b[j];
}

// row += Q * a  (quadratic x linear -> cubic, written straight into matrix-column order)
__device__ __forceinline__ void cubic_acc(double (&row)[20], const double (&Q)[10], const double (&a)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) row[col3_any(i, j, k)] += Q[q2(i, j)] * a[k];
}

// ------------------------------------------------------------------------------------------
// 1. null-space basis
// ------------------------------------------------------------------------------------------
__device__ inline void nullspace_basis(const double (&q)[5][2], const double (&qp)[5][2],
                                       double (&B)[4][9]) {
  double A[9][9];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double a0 = qp[i][0], a1 = qp[i][1], b0 = q[i][0], b1 = q[i][1];
    A[i][0] = a0 * b0; A[i][1] = a0 * b1; A[i][2] = a0;
    A[i][3] = a1 * b0; A[i][4] = a1 * b1; A[i][5] = a1;
    A[i][6] = b0;      A[i][7] = b1;      A[i][8] = 1.0;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) A[5 + i][j] = c_completion[i][j];
  // modified Gram-Schmidt, row by row
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) s += A[r][j] * A[r][j];
    const double f = 1.0 / sqrt(s);
#pragma unroll
    for (int j = 0; j < 9; ++j) A[r][j] *= f;
#pragma unroll
    for (int i = r + 1; i < 9; ++i) {
      double d = 0.0;
#pragma unroll
      for (int j = 0; j < 9; ++j) d += A[r][j] * A[i][j];
#pragma unroll
      for (int j = 0; j < 9; ++j) A[i][j] -= d * A[r][j];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < 9; ++c) B[k][c] = A[5 + k][c];
}

// ------------------------------------------------------------------------------------------
// 2. the ten cubic constraints  ->  M[10][20]
// ------------------------------------------------------------------------------------------
__device__ inline void build_constraints(const double (&B)[4][9], double (*M)[20]) {
  double e[3][3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) e[i][j][k] = B[k][3 * i + j];

  // G[i][q] = (E E^T)(i,q), symmetric: 6 quadratics
  double G[3][3][10];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int qq = i; qq < 3; ++qq) {
#pragma unroll
      for (int t = 0; t < 10; ++t) G[i][qq][t] = 0.0;
#pragma unroll
      for (int p = 0; p < 3; ++p) quad_acc(G[i][qq], e[i][p], e[qq][p]);
    }
  // Lambda = E E^T - (1/2) tr(E E^T) I
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    const double h = 0.5 * (G[0][0][t] + G[1][1][t] + G[2][2][t]);
    G[0][0][t] -= h; G[1][1][t] -= h; G[2][2][t] -= h;
  }
  // rows 1..9: (Lambda E)(i,j) = sum_q Lambda(i,q) E(q,j)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double row[20];
#pragma unroll
      for (int t = 0; t < 20; ++t) row[t] = 0.0;
#pragma unroll
      for (int qq = 0; qq < 3; ++qq) cubic_acc(row, i <= qq ? G[i][qq] : G[qq][i], e[qq][j]);
#pragma unroll
      for (int t = 0; t < 20; ++t) M[1 + 3 * i + j][t] = row[t];
    }
  // row 0: det E, cofactor expansion along the first column
  {
    double row[20];
#pragma unroll
    for (int t = 0; t < 20; ++t) row[t] = 0.0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int a = (t + 1) % 3, b = (t + 2) % 3;
      double m[10], neg[4];
#pragma unroll
      for (int u = 0; u < 10; ++u) m[u] = 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) neg[u] = -e[b][1][u];
      quad_acc(m, e[a][1], e[b][2]);
      quad_acc(m, neg, e[a][2]);
      cubic_acc(row, m, e[t][0]);
    }
#pragma unroll
    for (int t = 0; t < 20; ++t) M[0][t] = row[t];
  }
}

// ------------------------------------------------------------------------------------------
// 3. Gauss-Jordan on columns 0..9 (only rows 4..9 are needed afterwards)
// ------------------------------------------------------------------------------------------
__device__ inline bool eliminate(double (*M)[20]) {
  // Fully unrolled over the pivot column so that every inner loop has compile-time bounds and
  // addresses; the only dynamic index is the pivot row p.
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 10; ++c) {
    int p = c;
    double best = fabs(M[c][c]);
#pragma unroll
    for (int r = c + 1; r < 10; ++r) {
      const double v = fabs(M[r][c]);
      if (v > best) { best = v; p = r; }
    }
    ok = ok && (best > 1e-300) && (best < 1e300);
    double prow[20];
#pragma unroll
    for (int j = c; j < 20; ++j) prow[j] = M[p][j];
    if (p != c) {
#pragma unroll
      for (int j = c; j < 20; ++j) M[p][j] = M[c][j];
    }
    const double inv = 1.0 / prow[c];
#pragma unroll
    for (int j = c + 1; j < 20; ++j) prow[j] *= inv;
#pragma unroll
    for (int j = c + 1; j < 20; ++j) M[c][j] = prow[j];
#pragma unroll
    for (int r = c + 1; r < 10; ++r) {
      const double f = M[r][c];
#pragma unroll
      for (int j = c + 1; j < 20; ++j) M[r][j] = fma(-f, prow[j], M[r][j]);
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int c = 9; c >= 5; --c) {
    double prow[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) prow[j] = M[c][10 + j];
#pragma unroll
    for (int r = 4; r < c; ++r) {
      const double f = M[r][c];
#pragma unroll
      for (int j = 0; j < 10; ++j) M[r][10 + j] = fma(-f, prow[j], M[r][10 + j]);
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// 4. 3x3 polynomial matrix and its determinant
//    Bp[r][c][k]: coefficient of w^k; c = 0,1 (x,y: degree 3), c = 2 (constant: degree 4)
// ------------------------------------------------------------------------------------------
__device__ inline void hidden_matrix(const double (*M)[20], double (&Bp)[3][3][5]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const double* a = M[4 + 2 * r];  // monomial * w
    const double* b = M[5 + 2 * r];  // monomial
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int o = 10 + 3 * c;      // columns (v w2, v w, v)
      Bp[r][c][0] = a[o + 2];
      Bp[r][c][1] = a[o + 1] - b[o + 2];
      Bp[r][c][2] = a[o] - b[o + 1];
      Bp[r][c][3] = -b[o];
      Bp[r][c][4] = 0.0;
    }
    Bp[r][2][0] = a[19];
    Bp[r][2][1] = a[18] - b[19];
    Bp[r][2][2] = a[17] - b[18];
    Bp[r][2][3] = a[16] - b[17];
    Bp[r][2][4] = -b[16];
  }
}

__device__ inline void hidden_determinant(const double (&Bp)[3][3][5], double (&poly)[11]) {
#pragma unroll
  for (int i = 0; i < 11; ++i) poly[i] = 0.0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int r0 = t, r1 = (t + 1) % 3, r2 = (t + 2) % 3;
    double two[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) two[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        two[i + j] += Bp[r1][0][i] * Bp[r2][1][j] - Bp[r2][0][i] * Bp[r1][1][j];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) poly[i + j] += Bp[r0][2][j] * two[i];
  }
}

// ------------------------------------------------------------------------------------------
// 5. real roots of a degree-10 polynomial: Sturm chain, bisection isolation, Newton polish
// ------------------------------------------------------------------------------------------
struct SturmChain {
  double c[11][11];  // c[k][i]: coefficient of u^i in chain member k
  int deg[11];
  int n;             // last valid member index
};

__device__ __forceinline__ double horner(const double* c, int deg, double x) {
  double f = c[deg];
  for (int i = deg - 1; i >= 0; --i) f = fma(f, x, c[i]);
  return f;
}

__device__ inline int sturm_changes(const SturmChain& s, double x) {
  int ch = 0;
  double lf = horner(s.c[0], s.deg[0], x);
  for (int k = 1; k <= s.n; ++k) {
    const double f = horner(s.c[k], s.deg[k], x);
    if (lf == 0.0 || lf * f < 0.0) ++ch;
    lf = f;
  }
  return ch;
}

__device__ inline int sturm_changes_inf(const SturmChain& s, bool neg) {
  int ch = 0;
  double lf = s.c[0][s.deg[0]];
  if (neg && (s.deg[0] & 1)) lf = -lf;
  for (int k = 1; k <= s.n; ++k) {
    double f = s.c[k][s.deg[k]];
    if (neg && (s.deg[k] & 1)) f = -f;
    if (lf == 0.0 || lf * f < 0.0) ++ch;
    lf = f;
  }
  return ch;
}

// Builds the chain for the monic polynomial already stored in s.c[0] (degree 10).
__device__ inline void sturm_build(SturmChain& s) {
  const double kSmall = 1.0e-12;  // a remainder coefficient below this is zero (lead coefficients are +-1)
  s.deg[0] = 10;
  s.deg[1] = 9;
  {
    const double f = fabs(s.c[0][10] * 10.0);
    for (int i = 1; i <= 10; ++i) s.c[1][i - 1] = s.c[0][i] * i / f;
  }
  int k = 2;
  for (; k <= 10; ++k) {
    const double* u = s.c[k - 2];
    const double* v = s.c[k - 1];
    const int du = s.deg[k - 2], dv = s.deg[k - 1];
    double* r = s.c[k];
    for (int i = 0; i <= du; ++i) r[i] = u[i];
    const double lead = v[dv];  // +-1
    for (int t = du - dv; t >= 0; --t) {
      const double f = r[dv + t] * lead;  // division by +-1
      for (int j = 0; j < dv; ++j) r[j + t] = fma(-f, v[j], r[j + t]);
    }
    int d = dv - 1;
    while (d >= 0 && fabs(r[d]) < kSmall) { r[d] = 0.0; --d; }
    if (d <= 0) {  // constant (or zero) remainder ends the chain; negate as Sturm requires
      s.deg[k] = 0;
      r[0] = -r[0];
      break;
    }
    const double g = -1.0 / fabs(r[d]);  // negate and normalise
    for (int i = 0; i <= d; ++i) r[i] *= g;
    s.deg[k] = d;
  }
  s.n = k > 10 ? 10 : k;
}

// One root strictly isolated in [lo,hi] (Sturm count 1).  Newton with a bisection safeguard on
// a sign-changing bracket; if the end points do not change sign, bisect on Sturm counts.
__device__ inline double refine_root(const SturmChain& s, double lo, double hi, int vlo) {
  const double* p = s.c[0];
  double flo = horner(p, 10, lo), fhi = horner(p, 10, hi);
  if (flo == 0.0) return lo;
  if (fhi == 0.0) return hi;
  if ((flo < 0.0) == (fhi < 0.0)) {
    for (int it = 0; it < 64; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (!(mid > lo && mid < hi)) break;
      if (vlo - sturm_changes(s, mid) == 0) lo = mid; else hi = mid;
      flo = horner(p, 10, lo); fhi = horner(p, 10, hi);
      if ((flo < 0.0) != (fhi < 0.0)) break;
    }
    if ((flo < 0.0) == (fhi < 0.0)) return 0.5 * (lo + hi);
  }
  double x = 0.5 * (lo + hi);
  for (int it = 0; it < 48; ++it) {
    double f = p[10], df = 0.0;
#pragma unroll
    for (int i = 9; i >= 0; --i) { df = fma(df, x, f); f = fma(f, x, p[i]); }
    if (f == 0.0) return x;
    if ((f < 0.0) == (flo < 0.0)) lo = x; else hi = x;
    double xn = x - f / df;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (!(xn > lo && xn < hi)) return x;  // bracket collapsed to neighbouring doubles
    if (fabs(xn - x) <= 1.0e-15 * fabs(xn)) return xn;
    x = xn;
  }
  return x;
}

// roots ascending; returns the count.  poly[0..10] ascending powers of w.
// Generic (variable-degree chain, local-memory) implementation: used only when a Sturm
// remainder loses more than one degree (degenerate polynomials).
__device__ __noinline__ int real_roots_deg10_generic(const double (&poly)[11], double (&roots)[10]) {
  const double lead = poly[10];
  if (lead == 0.0) return 0;
  SturmChain s;
  const double inv = 1.0 / lead;
  bool finite = true;
#pragma unroll
  for (int i = 0; i <= 10; ++i) {
    s.c[0][i] = poly[i] * inv;
    finite = finite && (fabs(s.c[0][i]) < 1e300);
  }
  if (!finite) return 0;
  s.c[0][10] = 1.0;
  // w = u / fac so that the monic polynomial in u has |constant term| = 1 when it was > 10
  double fac = 1.0;
  const double val0 = fabs(s.c[0][0]);
  if (val0 > 10.0) {
    fac = pow(val0, -0.1);
    double mult = fac;
    for (int i = 9; i >= 0; --i) { s.c[0][i] *= mult; mult *= fac; }
  }
  sturm_build(s);
  const int vneg = sturm_changes_inf(s, true), vpos = sturm_changes_inf(s, false);
  if (vneg - vpos <= 0) return 0;
  // Cauchy bound: every root of a monic polynomial has |u| <= 1 + max |c_i|
  double bound = 0.0;
#pragma unroll
  for (int i = 0; i < 10; ++i) bound = fmax(bound, fabs(s.c[0][i]));
  bound += 1.0;
  int vlo0 = sturm_changes(s, -bound), vhi0 = sturm_changes(s, bound);
  int total = vlo0 - vhi0;
  if (total <= 0) return 0;
  if (total > 10) total = 10;

  // explicit interval stack; the left half is always processed first -> ascending roots
  double slo[12], shi[12];
  int svlo[12], svhi[12];
  int sp = 0, nr = 0;
  slo[0] = -bound; shi[0] = bound; svlo[0] = vlo0; svhi[0] = vhi0; sp = 1;
  while (sp > 0 && nr < 10) {
    --sp;
    double lo = slo[sp], hi = shi[sp];
    int vlo = svlo[sp], vhi = svhi[sp];
    int n = vlo - vhi;
    if (n <= 0) continue;
    if (n == 1) { roots[nr++] = refine_root(s, lo, hi, vlo); continue; }
    bool split = false;
    for (int it = 0; it < 200; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (!(mid > lo && mid < hi)) break;
      const int vmid = sturm_changes(s, mid);
      const int n1 = vlo - vmid, n2 = vmid - vhi;
      if (n1 > 0 && n2 > 0) {
        if (sp + 2 <= 12) {
          slo[sp] = mid; shi[sp] = hi; svlo[sp] = vmid; svhi[sp] = vhi; ++sp;  // right, popped later
          slo[sp] = lo; shi[sp] = mid; svlo[sp] = vlo; svhi[sp] = vmid; ++sp;  // left, popped next
          split = true;
        }
        break;
      }
      if (n1 == 0) lo = mid; else hi = mid;
    }
    if (!split) {  // roots closer than floating-point resolution: report them at the midpoint
      const double mid = 0.5 * (lo + hi);
      for (int i = 0; i < n && nr < 10; ++i) roots[nr++] = mid;
    }
  }
  const double back = 1.0 / fac;
  for (int i = 0; i < nr; ++i) roots[i] *= back;
  return nr;
}

// ------------------------------------------------------------------------------------------
// 5b. fast path: the generic situation where every Sturm remainder drops exactly one degree.
//     All loops have compile-time bounds, so the 66 chain coefficients live in registers and a
//     sign-change count is 55 independent-ish FMAs instead of 55 dependent local-memory loads.
//     Control flow is kept warp-friendly: one flat isolation loop (one Sturm count per trip),
//     then one bracketed Newton loop per root.
// ------------------------------------------------------------------------------------------
#ifdef TV5_SOLVE_PROFILE
#define TV5_RTICK(i) do { if ((threadIdx.x & 31) == 0) { long long t__ = clock64(); atomicAdd(&g_solve_prof[i], (unsigned long long)(t__ - rt_prev)); rt_prev = t__; } } while (0)
#else
#define TV5_RTICK(i)
#endif
struct FastChain {
  double c[11][11];  // member k has degree 10-k; only c[k][0 .. 10-k] is used
};

__device__ __forceinline__ int fast_changes(const FastChain& s, double x) {
  double f[11];
  const double x2 = x * x;
#pragma unroll
  for (int k = 0; k <= 10; ++k) {
    const int deg = 10 - k;
    if (deg >= 5) {   // long members: even and odd parts as two independent chains in x^2
      const int te = deg & ~1, to = (deg - 1) | 1;   // top even / odd exponent
      double ve = s.c[k][te], vo = s.c[k][to];
#pragma unroll
      for (int i = te - 2; i >= 0; i -= 2) ve = fma(ve, x2, s.c[k][i]);
#pragma unroll
      for (int i = to - 2; i >= 1; i -= 2) vo = fma(vo, x2, s.c[k][i]);
      f[k] = fma(vo, x, ve);
    } else {
      double v = s.c[k][deg];
#pragma unroll
      for (int i = deg - 1; i >= 0; --i) v = fma(v, x, s.c[k][i]);
      f[k] = v;
    }
  }
  int ch = 0;
#pragma unroll
  for (int k = 1; k <= 10; ++k) ch += (f[k - 1] == 0.0 || f[k - 1] * f[k] < 0.0) ? 1 : 0;
  return ch;
}

__device__ __forceinline__ int fast_changes_inf(const FastChain& s, bool neg) {
  int ch = 0;
#pragma unroll
  for (int k = 1; k <= 10; ++k) {
    double a = s.c[k - 1][11 - k], b = s.c[k][10 - k];
    if (neg && ((11 - k) & 1)) a = -a;
    if (neg && ((10 - k) & 1)) b = -b;
    ch += (a == 0.0 || a * b < 0.0) ? 1 : 0;
  }
  return ch;
}

// returns false if some remainder's leading coefficient is negligible (-> generic path)
__device__ __forceinline__ bool fast_build(FastChain& s) {
  const double kSmall = 1.0e-12;
  {
    const double f = fabs(s.c[0][10] * 10.0);
#pragma unroll
    for (int i = 1; i <= 10; ++i) s.c[1][i - 1] = s.c[0][i] * i / f;
  }
  bool ok = true;
#pragma unroll
  for (int k = 2; k <= 10; ++k) {
    const int dv = 11 - k;             // degree of member k-1; member k-2 has degree dv+1
    const double lead = s.c[k - 1][dv];  // +-1
    double r[12];
#pragma unroll
    for (int i = 0; i <= dv + 1; ++i) r[i] = s.c[k - 2][i];
    const double f1 = r[dv + 1] * lead;
#pragma unroll
    for (int j = 0; j < dv; ++j) r[j + 1] = fma(-f1, s.c[k - 1][j], r[j + 1]);
    const double f0 = r[dv] * lead;
#pragma unroll
    for (int j = 0; j < dv; ++j) r[j] = fma(-f0, s.c[k - 1][j], r[j]);
    if (k == 10) {
      s.c[10][0] = -r[0];
    } else {
      const double top = fabs(r[dv - 1]);
      ok = ok && (top >= kSmall);
      const double g = -1.0 / top;
#pragma unroll
      for (int i = 0; i < dv; ++i) s.c[k][i] = r[i] * g;
    }
  }
  return ok;
}

__device__ __forceinline__ void eval_p_dp(const double (&p)[11], double x, double& f, double& df) {
  f = p[10];
  df = 0.0;
#pragma unroll
  for (int i = 9; i >= 0; --i) { df = fma(df, x, f); f = fma(f, x, p[i]); }
}

// The same two values with short dependency chains (the solver is FP64-latency bound): p and its
// derivative are split into even and odd parts, four independent Horner chains in x^2 of length
// <= 5 instead of one chain of 20.  dq = chain member 1 = p' / 10 (fast_build).
__device__ __forceinline__ void eval_p_dp_split(const double (&p)[11], const double (&dq)[11], double x,
                                                double& f, double& df) {
  const double x2 = x * x;
  double fe = p[10], fo = p[9], de = dq[8], dd = dq[9];
#pragma unroll
  for (int i = 8; i >= 0; i -= 2) fe = fma(fe, x2, p[i]);
#pragma unroll
  for (int i = 7; i >= 1; i -= 2) fo = fma(fo, x2, p[i]);
#pragma unroll
  for (int i = 6; i >= 0; i -= 2) de = fma(de, x2, dq[i]);
#pragma unroll
  for (int i = 7; i >= 1; i -= 2) dd = fma(dd, x2, dq[i]);
  f = fma(fo, x, fe);
  df = 10.0 * fma(dd, x, de);
}

// Newton part of the refinement: p has opposite signs at lo and hi (flo = p(lo)).  dq = p'/10.
__device__ __forceinline__ double newton_bracketed(const double (&p)[11], const double (&dq)[11], double lo,
                                                   double hi, double flo) {
  // bracketed Newton (bisect when Newton leaves the bracket or converges too slowly)
  double xl = flo < 0.0 ? lo : hi, xh = flo < 0.0 ? hi : lo;
  double x = 0.5 * (lo + hi), dxold = fabs(hi - lo), dx = dxold, f, df;
  eval_p_dp_split(p, dq, x, f, df);
  for (int it = 0; it < 64; ++it) {
#ifdef TV5_SOLVE_COUNT
    atomicAdd(&g_solve_prof[10], 1ull);
#endif
    const bool bisect = (((x - xh) * df - f) * ((x - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df));
    dxold = dx;
    if (bisect) { dx = 0.5 * (xh - xl); x = xl + dx; } else { dx = f / df; x -= dx; }
    if (!(fabs(dx) > 1.0e-15 * fabs(x))) break;
    eval_p_dp_split(p, dq, x, f, df);
    if (f == 0.0) break;
    if (f < 0.0) xl = x; else xh = x;
  }
  return x;
}

// One root isolated in [lo,hi] by the Sturm counts (vlo - vhi == 1): make the bracket one with a
// sign change of p.  Returns 0 = bracket ready (flo = p(lo)), 1 = root found exactly / bracket
// collapsed (lo = the answer).
__device__ __forceinline__ int prepare_bracket(const FastChain& s, double& lo, double& hi, int vlo, double& flo) {
  const double (&p)[11] = s.c[0];
  double fhi, d;
  eval_p_dp(p, lo, flo, d);
  eval_p_dp(p, hi, fhi, d);
  if (flo == 0.0) return 1;
  if (fhi == 0.0) { lo = hi; return 1; }
  if ((flo < 0.0) == (fhi < 0.0)) {  // rare: no sign change at the ends; shrink with Sturm counts
    for (int it = 0; it < 60; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (!(mid > lo && mid < hi)) break;
      double fm;
      eval_p_dp(p, mid, fm, d);
      if (vlo - fast_changes(s, mid) == 0) { lo = mid; flo = fm; } else { hi = mid; fhi = fm; }
      if ((flo < 0.0) != (fhi < 0.0)) break;
    }
    if ((flo < 0.0) == (fhi < 0.0)) { lo = 0.5 * (lo + hi); return 1; }
  }
  return 0;
}

__device__ __forceinline__ double fast_refine(const FastChain& s, double lo, double hi, int vlo) {
  double flo;
  if (prepare_bracket(s, lo, hi, vlo, flo)) return lo;
#ifdef TV5_SOLVE_COUNT
  atomicAdd(&g_solve_prof[11], 1ull);
#endif
  return newton_bracketed(s.c[0], s.c[1], lo, hi, flo);
}

// Isolation of the real roots of poly (ascending powers of w) in the scaled variable u = w * fac.
// Returns the number of intervals ni (ascending); interval i holds exactly one root of the monic
// scaled polynomial s.c[0] in [ilo, ihi] with Sturm count ivlo at ilo, or is a collapsed cluster
// (ivlo < 0, root := ilo).  Returns -1 when the generic (local-memory) path must be used.
__device__ inline int isolate_roots_deg10(const double (&poly)[11], FastChain& s, double& back,
                                          double (&ilo)[10], double (&ihi)[10], int (&ivlo)[10]) {
  back = 1.0;
  const double lead = poly[10];
  if (lead == 0.0) return 0;
#ifdef TV5_SOLVE_PROFILE
  long long rt_prev = clock64();
#endif
  const double inv = 1.0 / lead;
  bool finite = true;
#pragma unroll
  for (int i = 0; i <= 10; ++i) {
    s.c[0][i] = poly[i] * inv;
    finite = finite && (fabs(s.c[0][i]) < 1e300);
  }
  if (!finite) return 0;
  s.c[0][10] = 1.0;
  double fac = 1.0;
  const double val0 = fabs(s.c[0][0]);
  if (val0 > 10.0) {  // same variable scaling rule as the generic path
    fac = exp2(-0.1 * log2(val0));
    double mult = fac;
#pragma unroll
    for (int i = 9; i >= 0; --i) { s.c[0][i] *= mult; mult *= fac; }
  }
  back = 1.0 / fac;
  if (!fast_build(s)) return -1;
  TV5_RTICK(6);
  if (fast_changes_inf(s, true) - fast_changes_inf(s, false) <= 0) return 0;
  double bound = 0.0;
#pragma unroll
  for (int i = 0; i < 10; ++i) bound = fmax(bound, fabs(s.c[0][i]));
  bound += 1.0;
  const int vlo0 = fast_changes(s, -bound), vhi0 = fast_changes(s, bound);
  if (vlo0 - vhi0 <= 0) return 0;

  // flat isolation loop: every trip performs at most one Sturm count.  The interval being worked on
  // (the top of the stack) lives in registers; only the right halves set aside by a split go through
  // the (local-memory) stack, so a trip that merely narrows an interval touches no memory at all
  // (the loop used to reload and store its four stack fields on every trip: 44 % of solve_roots'
  // stall samples were long-scoreboard waits on those local accesses).  Same trips, same order.
#if TV5_ISOLATE_TOP_IN_REGS
  double slo[11], shi[11];
  int svlo[11], svhi[11];
  int sp = 0, ni = 0;                                   // sp: intervals waiting on the stack
  double lo = -bound, hi = bound;
  int vlo = vlo0, vhi = vhi0;
  for (int trip = 0; trip < 400; ++trip) {
#ifdef TV5_SOLVE_COUNT
    atomicAdd(&g_solve_prof[8], 1ull);
#endif
    const int n = vlo - vhi;
    bool pop = false;
    if (n <= 0) {
      pop = true;
    } else if (n == 1) {
      if (ni < 10) { ilo[ni] = lo; ihi[ni] = hi; ivlo[ni] = vlo; ++ni; }
      pop = true;
    } else {
      const double mid = 0.5 * (lo + hi);
      if (!(mid > lo && mid < hi) || ni + n > 10) {  // unresolvable cluster: report at the midpoint
        for (int i = 0; i < n && ni < 10; ++i) { ilo[ni] = mid; ihi[ni] = mid; ivlo[ni] = -1; ++ni; }
        pop = true;
      } else {
#ifdef TV5_SOLVE_COUNT
        atomicAdd(&g_solve_prof[9], 1ull);
#endif
        const int vmid = fast_changes(s, mid);
        const int n1 = vlo - vmid, n2 = vmid - vhi;
        if (n1 > 0 && n2 > 0 && sp < 11) {   // split: right half set aside, left half next
          slo[sp] = mid; shi[sp] = hi; svlo[sp] = vmid; svhi[sp] = vhi; ++sp;
          hi = mid; vhi = vmid;
        } else if (n1 == 0) {
          lo = mid; vlo = vmid;
        } else {
          hi = mid; vhi = vmid;
        }
      }
    }
    if (pop) {
      if (sp == 0) break;
      --sp;
      lo = slo[sp]; hi = shi[sp]; vlo = svlo[sp]; vhi = svhi[sp];
    }
  }
#else
  double slo[12], shi[12];
  int svlo[12], svhi[12];
  int sp = 1, ni = 0;
  slo[0] = -bound; shi[0] = bound; svlo[0] = vlo0; svhi[0] = vhi0;
  for (int trip = 0; trip < 400 && sp > 0; ++trip) {
#ifdef TV5_SOLVE_COUNT
    atomicAdd(&g_solve_prof[8], 1ull);
#endif
    const int t = sp - 1;
    const int n = svlo[t] - svhi[t];
    if (n <= 0) { --sp; continue; }
    if (n == 1) {
      if (ni < 10) { ilo[ni] = slo[t]; ihi[ni] = shi[t]; ivlo[ni] = svlo[t]; ++ni; }
      --sp;
      continue;
    }
    const double lo = slo[t], hi = shi[t];
    const double mid = 0.5 * (lo + hi);
    if (!(mid > lo && mid < hi) || ni + n > 10) {  // unresolvable cluster: report at the midpoint
      for (int i = 0; i < n && ni < 10; ++i) { ilo[ni] = mid; ihi[ni] = mid; ivlo[ni] = -1; ++ni; }
      --sp;
      continue;
    }
#ifdef TV5_SOLVE_COUNT
    atomicAdd(&g_solve_prof[9], 1ull);
#endif
    const int vmid = fast_changes(s, mid);
    const int n1 = svlo[t] - vmid, n2 = vmid - svhi[t];
    if (n1 > 0 && n2 > 0 && sp < 12) {   // split: right half below, left half on top (popped first)
      slo[t] = mid; svlo[t] = vmid;                       // right: [mid, hi]
      slo[sp] = lo; shi[sp] = mid; svlo[sp] = n1 + vmid; svhi[sp] = vmid; ++sp;
    } else if (n1 == 0) {
      slo[t] = mid; svlo[t] = vmid;
    } else {
      shi[t] = mid; svhi[t] = vmid;
    }
  }
#endif
  TV5_RTICK(7);
  return ni;
}

// roots ascending; returns the count.  poly[0..10] ascending powers of w.
__device__ inline int real_roots_deg10(const double (&poly)[11], double (&roots)[10]) {
  FastChain s;
  double back, ilo[10], ihi[10];
  int ivlo[10];
  const int ni = isolate_roots_deg10(poly, s, back, ilo, ihi, ivlo);
  if (ni < 0) return real_roots_deg10_generic(poly, roots);
  for (int i = 0; i < ni; ++i)
    roots[i] = (ivlo[i] < 0 ? ilo[i] : fast_refine(s, ilo[i], ihi[i], ivlo[i])) * back;
  return ni;
}

// ------------------------------------------------------------------------------------------
// 6. E for one root
// ------------------------------------------------------------------------------------------
__device__ inline bool essential_from_root(const double (&B)[4][9], const double (&Bp)[3][3][5],
                                           double w, double (&E)[9]) {
  double m[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double f = Bp[r][c][4];
#pragma unroll
      for (int k = 3; k >= 0; --k) f = fma(f, w, Bp[r][c][k]);
      m[r][c] = f;
    }
  // (x, y, 1) spans the null space of m: take the best-conditioned cross product of two rows
  double best = -1.0, bx = 0.0, by = 0.0, bz = 1.0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int a = t, b = (t + 1) % 3;
    const double cx = m[a][1] * m[b][2] - m[a][2] * m[b][1];
    const double cy = m[a][2] * m[b][0] - m[a][0] * m[b][2];
    const double cz = m[a][0] * m[b][1] - m[a][1] * m[b][0];
    const double nrm = cx * cx + cy * cy + cz * cz;
    if (nrm > best) { best = nrm; bx = cx; by = cy; bz = cz; }
  }
  const double x = bx / bz, y = by / bz;
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    E[c] = w * B[0][c] + x * B[1][c] + y * B[2][c] + B[3][c];
    ok = ok && (fabs(E[c]) < 1e300);
  }
  return ok;
}

// ------------------------------------------------------------------------------------------
// 7. cheirality: closed-form twisted pair + unanimous vote of the 5 sample points
//    returns true and P = [R | t] (||t|| = 1) when one of the four candidates puts all five
//    points in front of both cameras.
// ------------------------------------------------------------------------------------------
__device__ inline bool pose_from_essential(const double (&E)[9], const double (&q)[5][2],
                                           const double (&qp)[5][2], double (&P)[12]) {
  // b b^T = (1/2) tr(E E^T) I - E E^T   (E = [b]x R)
  double EEt[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      EEt[i][j] = E[3 * i] * E[3 * j] + E[3 * i + 1] * E[3 * j + 1] + E[3 * i + 2] * E[3 * j + 2];
  const double half_tr = 0.5 * (EEt[0][0] + EEt[1][1] + EEt[2][2]);
  if (!(half_tr > 0.0)) return false;
  double bb[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) bb[i][j] = (i == j ? half_tr : 0.0) - EEt[i][j];
  int k = 0;
  if (bb[1][1] > bb[k][k]) k = 1;
  if (bb[2][2] > bb[k][k]) k = 2;
  if (!(bb[k][k] > 0.0)) return false;
  const double s = 1.0 / sqrt(bb[k][k]);
  const double b[3] = {bb[k][0] * s, bb[k][1] * s, bb[k][2] * s};  // |b|^2 = half_tr
  // cofactor matrix of E
  double C[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      C[i][j] = E[3 * i1 + j1] * E[3 * i2 + j2] - E[3 * i1 + j2] * E[3 * i2 + j1];
    }
  // [b]x E
  double bE[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    bE[0][j] = b[1] * E[6 + j] - b[2] * E[3 + j];
    bE[1][j] = b[2] * E[j] - b[0] * E[6 + j];
    bE[2][j] = b[0] * E[3 + j] - b[1] * E[j];
  }
  const double inv = 1.0 / half_tr;
  const double bn = 1.0 / sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);  // exactly unit t
  const double t[3] = {b[0] * bn, b[1] * bn, b[2] * bn};
  // |b|^2 R = Cof(E)^T -+ [b]x E : the two rotations of the twisted pair
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    double R[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = (C[i][j] + (which ? bE[i][j] : -bE[i][j])) * inv;
    int front = 0, behind = 0;
#pragma unroll
    for (int p = 0; p < 5; ++p) {
      const double x1[3] = {q[p][0], q[p][1], 1.0}, x2[3] = {qp[p][0], qp[p][1], 1.0};
      double a[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = R[i][0] * x1[0] + R[i][1] * x1[1] + R[i][2];
      // z2 x2 = z1 a + t :  z1 = -(x2 x t).(x2 x a)/|x2 x a|^2 ,  z2 = (a x t).(a x x2)/|a x x2|^2
      const double c1[3] = {x2[1] * t[2] - x2[2] * t[1], x2[2] * t[0] - x2[0] * t[2], x2[0] * t[1] - x2[1] * t[0]};
      const double c2[3] = {x2[1] * a[2] - x2[2] * a[1], x2[2] * a[0] - x2[0] * a[2], x2[0] * a[1] - x2[1] * a[0]};
      const double c3[3] = {a[1] * t[2] - a[2] * t[1], a[2] * t[0] - a[0] * t[2], a[0] * t[1] - a[1] * t[0]};
      const double s1 = -(c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2]);
      const double s2 = -(c3[0] * c2[0] + c3[1] * c2[1] + c3[2] * c2[2]);
      front += (s1 > 0.0) + (s2 > 0.0);
      behind += (s1 < 0.0) + (s2 < 0.0);
    }
    double sgn;
    if (front == 10) sgn = 1.0;
    else if (behind == 10) sgn = -1.0;
    else continue;
    // E from an ill-conditioned root is only approximately essential, and the closed form then
    // returns an R that is only approximately orthonormal: re-orthonormalise (Gram-Schmidt on
    // the first two rows, third by cross product), as the reference's Givens construction
    // guarantees by design.  For a proper E this changes R at the 1e-16 level.
    {
      double n0 = 1.0 / sqrt(R[0][0] * R[0][0] + R[0][1] * R[0][1] + R[0][2] * R[0][2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) R[0][j] *= n0;
      const double d = R[1][0] * R[0][0] + R[1][1] * R[0][1] + R[1][2] * R[0][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) R[1][j] -= d * R[0][j];
      const double n1 = 1.0 / sqrt(R[1][0] * R[1][0] + R[1][1] * R[1][1] + R[1][2] * R[1][2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) R[1][j] *= n1;
      R[2][0] = R[0][1] * R[1][2] - R[0][2] * R[1][1];
      R[2][1] = R[0][2] * R[1][0] - R[0][0] * R[1][2];
      R[2][2] = R[0][0] * R[1][1] - R[0][1] * R[1][0];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      P[4 * i] = R[i][0]; P[4 * i + 1] = R[i][1]; P[4 * i + 2] = R[i][2];
      P[4 * i + 3] = sgn * t[i];
    }
    return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// driver: one minimal set.  E_out/P_out are per-thread global slices [10][9] / [10][12].
// Returns n_valid (after cheirality when requested); n_roots_out receives the real-root count.
// ------------------------------------------------------------------------------------------
#ifdef TV5_SOLVE_PROFILE
#define TV5_TICK(i) do { if ((threadIdx.x & 31) == 0) { long long t__ = clock64(); atomicAdd(&g_solve_prof[i], (unsigned long long)(t__ - t_prev)); t_prev = t__; } } while (0)
#else
#define TV5_TICK(i)
#endif

__device__ inline int solve_minimal_set(const double (&q)[5][2], const double (&qp)[5][2],
                                        bool with_cheirality, double* E_out, double* P_out,
                                        int* n_roots_out) {
  double B[4][9];
  double Bp[3][3][5];
  double poly[11];
  *n_roots_out = 0;
#ifdef TV5_SOLVE_PROFILE
  long long t_prev = clock64();
#endif
  {
    double M[10][20];
    nullspace_basis(q, qp, B);
    TV5_TICK(0);
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 9; ++c) ok = ok && (fabs(B[3][c]) <= 1.0);  // false on NaN (degenerate set)
    if (!ok) return 0;
    build_constraints(B, M);
    TV5_TICK(1);
    if (!eliminate(M)) return 0;
    hidden_matrix(M, Bp);
    TV5_TICK(2);
  }
  hidden_determinant(Bp, poly);
  TV5_TICK(3);
  double roots[10];
  const int nr = real_roots_deg10(poly, roots);
  TV5_TICK(4);
  *n_roots_out = nr;
  int nv = 0;
  for (int i = 0; i < nr; ++i) {
    double E[9], P[12];
    if (!essential_from_root(B, Bp, roots[i], E)) continue;
    if (with_cheirality) {
      if (!pose_from_essential(E, q, qp, P)) continue;
      if (P_out)
#pragma unroll
        for (int c = 0; c < 12; ++c) P_out[12 * nv + c] = P[c];
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) E_out[9 * nv + c] = E[c];
    ++nv;
  }
  TV5_TICK(5);
  return nv;
}

}  // namespace tv5
