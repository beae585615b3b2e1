/*
 * tv5_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see tv5_oracle.h).
 *
 * A plain-C restatement of the algorithm the reference runs per RANSAC hypothesis:
 *   null space (Gram-Schmidt, fixed completion rows) -> 10 cubic constraints in (w;x,y) ->
 *   pivoted reduction to a 3x3 polynomial matrix -> degree-10 determinant in the hidden
 *   variable w -> Sturm isolation + bracketed refinement -> E per root -> cheirality/P ->
 *   Sampson scoring -> first-max selection.
 * Every function cites the reference lines it follows ("ref:" = path relative to
 * /root/reference/RANSAC_FiveP/essential_matrix/).  Nothing here is used by the product path.
 *
 * Parity status: PINNED — against the reference's own sources compiled for the host and for the GPU
 * (oracle/_ref/, build_ref.sh), against fixtures generated from them (the .npz files under tests/golden) and, for the
 * polish functions, against the reference extension imported in the build container (DESIGN.md
 * section 6; tests/test_oracle.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared tv5_oracle.c -o libtv5_oracle.so -lm
 * (-ffp-contract=off: the only fused operations are the explicit fma() calls in
 * tv5o_sampson_err, which reproduce the contraction nvcc applies to the reference kernel.)
 */
#include "tv5_oracle.h"

#include <math.h>
#include <string.h>

#define DEG 10

/* ------------------------------------------------------------------------------------------
 * Sampson distance.  ref: kernel_functions.cu:231-264 (ComputeError<double>).
 * Operation order = the SASS nvcc 12.9 emits for sm_100a with default flags (--fmad=true):
 * the `* 1.0` homogeneous terms fold away, `a*b + c*d` contracts to fma(c,d, a*b), sqrt and
 * division are IEEE round-to-nearest.  Verified bit-for-bit on the GPU box against the
 * reference's own function by tests/test_gpu_parity.py::test_oracle_sampson_bits_match_reference_gpu.
 * ---------------------------------------------------------------------------------------- */
double tv5o_sampson_err(const double E[9], double x1, double y1, double x2, double y2) {
  double Ex0 = fma(E[1], y1, E[0] * x1) + E[2];
  double Ex1 = fma(E[4], y1, E[3] * x1) + E[5];
  double Ex2 = fma(E[7], y1, E[6] * x1) + E[8];
  double tE0 = fma(E[3], y2, E[0] * x2) + E[6];
  double tE1 = fma(E[4], y2, E[1] * x2) + E[7];
  double num = Ex2 + fma(y2, Ex1, x2 * Ex0);
  double den = fma(tE1, tE1, fma(tE0, tE0, fma(Ex0, Ex0, Ex1 * Ex1)));
  return fabs(num / sqrt(den));
}

/* ref: kernel_functions.cu:187-197 (the `error <= c_inlier_threshold` loops). */
void tv5o_score(const double* x1, const double* x2, int n, const double* E_list, int M,
                double thr, int32_t* counts, uint8_t* mask) {
  for (int m = 0; m < M; ++m) {
    const double* E = E_list + 9 * (size_t)m;
    int c = 0;
    for (int k = 0; k < n; ++k) {
      double err = tv5o_sampson_err(E, x1[2 * k], x1[2 * k + 1], x2[2 * k], x2[2 * k + 1]);
      int in = (err <= thr); /* NaN -> outlier, as in the reference */
      c += in;
      if (mask) mask[(size_t)m * n + k] = (uint8_t)in;
    }
    counts[m] = c;
  }
}

/* ref: kernel_functions.cu:269-278 (RandomInt with min_int = 0, max_int = N-1), float math. */
int32_t tv5o_index_from_uniform(float u, int N) {
  float r = u;
  r *= ((float)(N - 1) - 0.0f + 0.999999f); /* (max_int - min_int + 0.999999f) in float */
  r += 0.0f;
  return (int32_t)truncf(r);
}

/* ------------------------------------------------------------------------------------------
 * Null space.  ref: essential_matrix_5pt.cu:680-699 (constraint rows qp (x) q) and :631-678
 * (completion by a fixed pseudo-random sequence, modified Gram-Schmidt, basis = rows 5..8).
 * B[k][c]: k = 0..3 multiplies the unknowns (w, x, y, 1); c = 3*i + j indexes E(i,j).
 * ---------------------------------------------------------------------------------------- */
void tv5o_nullspace_basis(const double q[5][2], const double qp[5][2], double B[4][9]) {
  double M[9][9];
  for (int i = 0; i < 5; ++i) {
    const double a[3] = {qp[i][0], qp[i][1], 1.0};
    const double b[3] = {q[i][0], q[i][1], 1.0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[i][3 * r + c] = a[r] * b[c];
  }
  double ran = 3.18730379;
  for (int i = 5; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      ran *= 3.18730379;
      ran = 2.0 * (ran - floor(ran)) - 1.0;
      M[i][j] = ran;
    }
  for (int r = 0; r < 9; ++r) {
    double s = 0.0;
    for (int j = 0; j < 9; ++j) s += M[r][j] * M[r][j];
    double f = 1.0 / sqrt(s);
    for (int j = 0; j < 9; ++j) M[r][j] *= f;
    for (int i = r + 1; i < 9; ++i) {
      double d = 0.0;
      for (int j = 0; j < 9; ++j) d += M[r][j] * M[i][j];
      for (int j = 0; j < 9; ++j) M[i][j] -= d * M[r][j];
    }
  }
  for (int k = 0; k < 4; ++k)
    for (int c = 0; c < 9; ++c) B[k][c] = M[5 + k][c];
}

/* ------------------------------------------------------------------------------------------
 * Polynomial algebra in the 4 unknowns v = (w, x, y, z=1).  Homogeneous quadratics/cubics are
 * stored on sorted index tuples (i<=j<=k), as the reference's poly4_2 / poly4_3 do
 * (ref: essential_matrix_5pt.cu:26-120).
 * ---------------------------------------------------------------------------------------- */
static void quad_addmul(double Q[4][4], const double* a, const double* b) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      int lo = i < j ? i : j, hi = i < j ? j : i;
      Q[lo][hi] += a[i] * b[j];
    }
}

static void cub_addmul(double C[4][4][4], const double Q[4][4], const double* a, double s) {
  for (int i = 0; i < 4; ++i)
    for (int j = i; j < 4; ++j)
      for (int k = 0; k < 4; ++k) {
        int t0 = i, t1 = j, t2 = k; /* insert k into the sorted pair (i,j) */
        if (t2 < t1) { int t = t1; t1 = t2; t2 = t; }
        if (t1 < t0) { int t = t0; t0 = t1; t1 = t; }
        C[t0][t1][t2] += s * Q[i][j] * a[k];
      }
}

/* ref: essential_matrix_5pt.cu:370-426 (mono_coeff): cubic in (w,x,y,z) -> coefficients of the
 * 10 monomials {1,x,y,xx,xy,yy,xxx,xxy,xyy,yyy}, split by power of w into A[deg][row][col]. */
static void store_row(double A[5][10][10], int n, const double C[4][4][4]) {
  enum { W = 0, X = 1, Y = 2, Z = 3 };
  A[0][n][0] = C[Z][Z][Z]; A[0][n][1] = C[X][Z][Z]; A[0][n][2] = C[Y][Z][Z];
  A[0][n][3] = C[X][X][Z]; A[0][n][4] = C[X][Y][Z]; A[0][n][5] = C[Y][Y][Z];
  A[0][n][6] = C[X][X][X]; A[0][n][7] = C[X][X][Y]; A[0][n][8] = C[X][Y][Y];
  A[0][n][9] = C[Y][Y][Y];
  A[1][n][0] = C[W][Z][Z]; A[1][n][1] = C[W][X][Z]; A[1][n][2] = C[W][Y][Z];
  A[1][n][3] = C[W][X][X]; A[1][n][4] = C[W][X][Y]; A[1][n][5] = C[W][Y][Y];
  A[2][n][0] = C[W][W][Z]; A[2][n][1] = C[W][W][X]; A[2][n][2] = C[W][W][Y];
  A[3][n][0] = C[W][W][W];
}

/* ref: essential_matrix_5pt.cu:428-474 (EEeqns_5pt): row 0 = det E, rows 1..9 =
 * 2 E E^T E - tr(E E^T) E, entry by entry. */
static void build_constraints(const double B[4][9], double A[5][10][10]) {
  double e[3][3][4]; /* e[i][j][k]: coefficient of unknown k in E(i,j) */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 4; ++k) e[i][j][k] = B[k][3 * i + j];
  memset(A, 0, sizeof(double) * 5 * 10 * 10);

  double tr[4][4];
  memset(tr, 0, sizeof(tr));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) quad_addmul(tr, e[i][j], e[i][j]);

  /* determinant: ref :334-349 (polydet4) */
  {
    double C[4][4][4];
    memset(C, 0, sizeof(C));
    const int rows[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    for (int t = 0; t < 3; ++t) {
      int a = rows[t][0], b = rows[t][1];
      double m[4][4], neg[4];
      memset(m, 0, sizeof(m));
      quad_addmul(m, e[a][1], e[b][2]);
      for (int k = 0; k < 4; ++k) neg[k] = -e[b][1][k];
      quad_addmul(m, neg, e[a][2]);
      cub_addmul(C, m, e[t][0], 1.0);
    }
    store_row(A, 0, C);
  }

  int eqn = 1;
  for (int i = 0; i < 3; ++i) {
    double C[3][4][4][4];
    memset(C, 0, sizeof(C));
    for (int qq = 0; qq < 3; ++qq) {
      double EEt[4][4];
      memset(EEt, 0, sizeof(EEt));
      for (int p = 0; p < 3; ++p) quad_addmul(EEt, e[i][p], e[qq][p]);
      for (int j = 0; j < 3; ++j) cub_addmul(C[j], EEt, e[qq][j], 2.0);
    }
    for (int j = 0; j < 3; ++j) {
      cub_addmul(C[j], tr, e[i][j], -1.0);
      store_row(A, eqn++, C[j]);
    }
  }
}

/* Row operation limited to the structurally non-zero columns: degree-0 columns 0..ncol0,
 * 6 / 3 / 1 columns in degrees 1 / 2 / 3.  ref: essential_matrix_5pt.cu:713-781. */
static void row_sub(double A[5][10][10], int dst, int src, double fac, int ncol0) {
  for (int j = 0; j <= ncol0; ++j) A[0][dst][j] -= fac * A[0][src][j];
  for (int j = 0; j < 6; ++j) A[1][dst][j] -= fac * A[1][src][j];
  for (int j = 0; j < 3; ++j) A[2][dst][j] -= fac * A[2][src][j];
  A[3][dst][0] -= fac * A[3][src][0];
}

/* ref: :808-850 (pivot): bring the largest |A[0][i][last]|, i <= last, to row `last`. */
static void pivot_rows(double A[5][10][10], int last) {
  double best = fabs(A[0][last][last]);
  int row = last;
  for (int i = 0; i < last; ++i)
    if (fabs(A[0][i][last]) > best) { row = i; best = fabs(A[0][i][last]); }
  if (row == last) return;
  for (int d = 0; d < 4; ++d) {
    int n = d == 0 ? last + 1 : (d == 1 ? 6 : (d == 2 ? 3 : 1));
    for (int j = 0; j < n; ++j) {
      double t = A[d][last][j]; A[d][last][j] = A[d][row][j]; A[d][row][j] = t;
    }
  }
}

static void sweep_above(double A[5][10][10], int row, int col, int deg) {
  double piv = A[deg][row][col];
  for (int i = 0; i < row; ++i) row_sub(A, i, row, A[deg][i][col] / piv, col);
}

static void sweep_below(double A[5][10][10], int row, int col, int deg, int lastrow) {
  double piv = A[deg][row][col];
  for (int i = row + 1; i <= lastrow; ++i) row_sub(A, i, row, A[deg][i][col] / piv, col);
}

/* ref: :852-900 (reduce_Ematrix). */
static void reduce_to_3x3(double A[5][10][10]) {
  for (int last = 9; last >= 3; --last) { pivot_rows(A, last); sweep_above(A, last, last, 0); }
  sweep_below(A, 3, 3, 0, 5);
  sweep_below(A, 4, 4, 0, 5);
  sweep_above(A, 2, 5, 1);
  sweep_above(A, 1, 4, 1);
  sweep_below(A, 0, 3, 1, 5);
  sweep_below(A, 1, 4, 1, 5);
  sweep_below(A, 2, 5, 1, 5);
  for (int i = 0; i < 3; ++i) {
    double fac = A[1][i][3 + i] / A[0][3 + i][3 + i];
    A[4][i][0] = -A[3][i + 3][0] * fac;
    for (int j = 0; j < 3; ++j) {
      A[3][i][j] -= A[2][i + 3][j] * fac;
      A[2][i][j] -= A[1][i + 3][j] * fac;
      A[1][i][j] -= A[0][i + 3][j] * fac;
    }
  }
}

/* ref: :902-948 (one_cofactor, compute_determinant): det of the 3x3 polynomial matrix. */
static void hidden_determinant(double A[5][10][10], double poly[DEG + 1]) {
  memset(poly, 0, sizeof(double) * (DEG + 1));
  const int cyc[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
  for (int t = 0; t < 3; ++t) {
    int r0 = cyc[t][0], r1 = cyc[t][1], r2 = cyc[t][2];
    double two[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i <= 3; ++i)
      for (int j = 0; j <= 3; ++j)
        two[i + j] += A[i][r1][1] * A[j][r2][2] - A[i][r2][1] * A[j][r1][2];
    for (int i = 0; i <= 6; ++i)
      for (int j = 0; j <= 4; ++j) poly[i + j] += A[j][r0][0] * two[i];
  }
}

void tv5o_hidden_poly(const double q[5][2], const double qp[5][2], double poly[11]) {
  double B[4][9], A[5][10][10];
  tv5o_nullspace_basis(q, qp, B);
  build_constraints(B, A);
  reduce_to_3x3(A);
  hidden_determinant(A, poly);
}

/* ------------------------------------------------------------------------------------------
 * Real roots by Sturm sequence.  ref: sturm.cu (constants :13-18).
 * ---------------------------------------------------------------------------------------- */
#define RELERROR 1.0e-12
#define MAXPOW 32
#define MAXIT 800
#define SMALL_ENOUGH 1.0e-12
#define MAXORD 20

typedef struct { int ord; double coef[MAXORD + 1]; } spoly;

static double horner(int ord, const double* c, double x) {
  double f = c[ord];
  for (int i = ord - 1; i >= 0; --i) f = x * f + c[i];
  return f;
}

/* ref: sturm.cu:283-322 (modp) */
static int poly_mod(const spoly* u, const spoly* v, spoly* r) {
  for (int i = 0; i <= u->ord; ++i) r->coef[i] = u->coef[i];
  if (v->coef[v->ord] < 0.0) {
    for (int k = u->ord - v->ord - 1; k >= 0; k -= 2) r->coef[k] = -r->coef[k];
    for (int k = u->ord - v->ord; k >= 0; --k)
      for (int j = v->ord + k - 1; j >= k; --j)
        r->coef[j] = -r->coef[j] - r->coef[v->ord + k] * v->coef[j - k];
  } else {
    for (int k = u->ord - v->ord; k >= 0; --k)
      for (int j = v->ord + k - 1; j >= k; --j)
        r->coef[j] -= r->coef[v->ord + k] * v->coef[j - k];
  }
  int k = v->ord - 1;
  while (k >= 0 && fabs(r->coef[k]) < SMALL_ENOUGH) { r->coef[k] = 0.0; --k; }
  r->ord = k < 0 ? 0 : k;
  return r->ord;
}

/* ref: sturm.cu:331-360 (buildsturm) */
static int build_chain(int ord, spoly* s) {
  s[0].ord = ord;
  s[1].ord = ord - 1;
  double f = fabs(s[0].coef[ord] * ord);
  for (int i = 1; i <= ord; ++i) s[1].coef[i - 1] = s[0].coef[i] * i / f;
  int n = 2;
  while (poly_mod(&s[n - 2], &s[n - 1], &s[n])) {
    double g = -fabs(s[n].coef[s[n].ord]);
    for (int i = s[n].ord; i >= 0; --i) s[n].coef[i] /= g;
    ++n;
  }
  s[n].coef[0] = -s[n].coef[0];
  return n;
}

/* ref: sturm.cu:369-386 (numchanges) */
static int sign_changes(int np, const spoly* s, double a) {
  int ch = 0;
  double lf = horner(s[0].ord, s[0].coef, a);
  for (int i = 1; i <= np; ++i) {
    double f = horner(s[i].ord, s[i].coef, a);
    if (lf == 0.0 || lf * f < 0) ++ch;
    lf = f;
  }
  return ch;
}

/* ref: sturm.cu:394-439 (numroots, non_neg = false) */
static int count_roots(int np, const spoly* s, int* atneg, int* atpos) {
  int p = 0, n = 0;
  double lf = s[0].coef[s[0].ord];
  for (int i = 1; i <= np; ++i) {
    double f = s[i].coef[s[i].ord];
    if (lf == 0.0 || lf * f < 0) ++p;
    lf = f;
  }
  lf = (s[0].ord & 1) ? -s[0].coef[s[0].ord] : s[0].coef[s[0].ord];
  for (int i = 1; i <= np; ++i) {
    double f = (s[i].ord & 1) ? -s[i].coef[s[i].ord] : s[i].coef[s[i].ord];
    if (lf == 0.0 || lf * f < 0) ++n;
    lf = f;
  }
  *atneg = n;
  *atpos = p;
  return n - p;
}

/* Value of p scaled by |x|^-ord outside [-1,1] (same sign and zeros as p, no overflow).  The
 * reference achieves the same by inverting the interval, ref: sturm.cu:208-275 (modrf). */
static double scaled_eval(int ord, const double* c, double x) {
  if (fabs(x) <= 1.0) return horner(ord, c, x);
  double t = 1.0 / x, f = c[0];
  for (int i = 1; i <= ord; ++i) f = t * f + c[i];
  return ((ord & 1) && x < 0.0) ? -f : f;
}

/* One simple root in [a,b] with a sign change: regula falsi with Illinois damping and a
 * bisection safeguard.  Plays the role of ref: sturm.cu:43-152 (modrf_pos) but iterates to the
 * floating-point limit instead of the reference's 1e-12 relative stop.  Returns 0 when the end
 * points do not bracket a sign change (the caller then bisects on Sturm counts). */
static int refine_bracket(int ord, const double* c, double a, double b, double* root) {
  double fa = scaled_eval(ord, c, a), fb = scaled_eval(ord, c, b);
  if (fa == 0.0) { *root = a; return 1; }
  if (fb == 0.0) { *root = b; return 1; }
  if (fa * fb > 0.0 || fa != fa || fb != fb) return 0;
  int side = 0;
  for (int it = 0; it < 400; ++it) {
    double x = (fb * a - fa * b) / (fb - fa);
    if (!(x > a && x < b) || (it % 4) == 3) x = 0.5 * (a + b);
    if (!(x > a && x < b)) break; /* interval has collapsed to adjacent doubles */
    double fx = scaled_eval(ord, c, x);
    if (fx == 0.0) { *root = x; return 1; }
    if ((fa < 0) == (fx < 0)) {
      a = x; fa = fx;
      if (side == -1) fb *= 0.5;
      side = -1;
    } else {
      b = x; fb = fx;
      if (side == 1) fa *= 0.5;
      side = 1;
    }
    if (fabs(b - a) <= 4.0e-16 * fmax(fabs(a), fabs(b))) break;
  }
  *root = fabs(fa) < fabs(fb) ? a : b;
  return 1;
}

/* ref: sturm.cu:450-555 (sbisect) */
static void isolate(int np, const spoly* s, double lo, double hi, int atlo, int athi,
                    double* roots, int depth) {
  int nroot = atlo - athi;
  double mid = 0.5 * (lo + hi);
  if (nroot <= 0) return;
  if (depth >= 40) { for (int i = 0; i < nroot; ++i) roots[i] = mid; return; }
  if (nroot == 1) {
    if (refine_bracket(s[0].ord, s[0].coef, lo, hi, &roots[0])) return;
    for (int its = 0; its < MAXIT; ++its) {
      mid = (lo + hi) / 2;
      int atmid = sign_changes(np, s, mid);
      if (fabs(mid) > RELERROR) {
        if (fabs((hi - lo) / mid) < RELERROR) break;
      } else if (fabs(hi - lo) < RELERROR) break;
      if (atlo - atmid == 0) lo = mid; else hi = mid;
    }
    roots[0] = mid;
    return;
  }
  for (int its = 0; its < MAXIT; ++its) {
    mid = (lo + hi) / 2;
    int atmid = sign_changes(np, s, mid);
    int n1 = atlo - atmid, n2 = atmid - athi;
    if (n1 != 0 && n2 != 0) {
      isolate(np, s, lo, mid, atlo, atmid, roots, depth + 1);
      isolate(np, s, mid, hi, atmid, athi, roots + n1, depth + 1);
      return;
    }
    if (n1 == 0) lo = mid; else hi = mid;
  }
  for (int i = 0; i < nroot; ++i) roots[i] = mid; /* roots too close together */
}

/* ref: sturm.cu:557-676 (find_real_roots_sturm) */
int tv5o_real_roots(const double* p, int degree, double* roots) {
  spoly s[MAXORD + 2];
  memset(&s[0], 0, sizeof(s[0]));
  double norm = 1.0 / p[degree];
  for (int i = 0; i <= degree; ++i) s[0].coef[i] = p[i] * norm;
  double val0 = fabs(s[0].coef[0]), fac = 1.0;
  if (val0 > 10.0) {
    fac = pow(val0, -1.0 / degree);
    double mult = fac;
    for (int i = degree - 1; i >= 0; --i) { s[0].coef[i] *= mult; mult *= fac; }
  }
  for (int i = 0; i <= degree; ++i)
    if (!(fabs(s[0].coef[i]) < INFINITY)) return 0; /* NaN/Inf input: no solutions */
  int np = build_chain(degree, s);
  int atmin, atmax;
  int n = count_roots(np, s, &atmin, &atmax);
  if (n <= 0) return 0;
  double lo = -1.0, hi = 1.0;
  int ch = sign_changes(np, s, lo);
  for (int i = 0; ch != atmin && i != MAXPOW; ++i) { lo *= 10.0; ch = sign_changes(np, s, lo); }
  if (ch != atmin) atmin = ch;
  ch = sign_changes(np, s, hi);
  for (int i = 0; ch != atmax && i != MAXPOW; ++i) { hi *= 10.0; ch = sign_changes(np, s, hi); }
  if (ch != atmax) atmax = ch;
  n = atmin - atmax;
  if (n <= 0) return 0;
  if (n > degree) n = degree;
  isolate(np, s, lo, hi, atmin, atmax, roots, 0);
  for (int i = 0; i < n; ++i) roots[i] /= fac;
  return n;
}

/* ref: essential_matrix_5pt.cu:476-507 (null_space_solve_3x3_half_pivot) and :955-1015
 * (compute_E_matrix). */
static void e_from_root(const double B[4][9], double A[5][10][10], double w, double E[9]) {
  double w2 = w * w, w3 = w2 * w, w4 = w3 * w;
  double M[3][3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      M[i][j] = A[0][i][j] + w * A[1][i][j] + w2 * A[2][i][j] + w3 * A[3][i][j];
    M[i][0] += w4 * A[4][i][0];
  }
  int p1;
  double f0 = fabs(M[0][2]), f1 = fabs(M[1][2]), f2 = fabs(M[2][2]);
  if (f0 > f1) p1 = (f0 > f2) ? 0 : 2; else p1 = (f1 > f2) ? 1 : 2;
  int r1 = (p1 + 1) % 3, r2 = (p1 + 2) % 3;
  double fac = M[r1][2] / M[p1][2];
  M[r1][0] -= fac * M[p1][0];
  M[r1][1] -= fac * M[p1][1];
  fac = M[r2][2] / M[p1][2];
  M[r2][0] -= fac * M[p1][0];
  M[r2][1] -= fac * M[p1][1];
  int p2 = fabs(M[r1][1]) > fabs(M[r2][1]) ? r1 : r2;
  double x = -M[p2][0] / M[p2][1];
  double y = -(M[p1][0] + M[p1][1] * x) / M[p1][2];
  for (int c = 0; c < 9; ++c) E[c] = w * B[0][c] + x * B[1][c] + y * B[2][c] + B[3][c];
}

/* ref: essential_matrix_5pt.cu:1224-1249 (compute_E_matrices_optimized) */
int tv5o_solve5(const double q[5][2], const double qp[5][2], double E_out[10][9], double* w_out) {
  double B[4][9], A[5][10][10], poly[DEG + 1], roots[MAXORD];
  tv5o_nullspace_basis(q, qp, B);
  build_constraints(B, A);
  reduce_to_3x3(A);
  hidden_determinant(A, poly);
  for (int i = 0; i <= DEG; ++i)
    if (!(fabs(poly[i]) < INFINITY)) return 0;
  if (poly[DEG] == 0.0) return 0;
  int n = tv5o_real_roots(poly, DEG, roots);
  for (int i = 0; i < n; ++i) {
    e_from_root(B, A, roots[i], E_out[i]);
    if (w_out) w_out[i] = roots[i];
  }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Cheirality + P.  ref: cheirality.cu:4-214 (compute_P_matrices, focal == NULL, npoints = 5).
 * E = U diag(1,1,0) V^T by two/three Givens rotations; the four (R, +-t) candidates are tested
 * with closed-form depth signs on the 5 sample points; a candidate is kept only on a unanimous
 * vote.  P = [R | +-u3].
 * ---------------------------------------------------------------------------------------- */
int tv5o_cheirality(const double q[5][2], const double qp[5][2], double E[][9], int n,
                    double P[][12]) {
  int kept = 0;
  for (int m = 0; m < n; ++m) {
    double Ut[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double Vt[3][3];
    memcpy(Vt, E[m], sizeof(Vt));
    for (int i = 0; i <= 1; ++i)
      for (int k = i + 1; k < 3; ++k) {
        double a = Vt[i][i], b = Vt[k][i];
        double s = sqrt(a * a + b * b);
        if (s == 0.0) continue;
        a /= s;
        b /= s;
        Vt[i][i] = s;
        Vt[k][i] = 0.0;
        for (int j = i + 1; j < 3; ++j) {
          double c = Vt[i][j], d = Vt[k][j];
          Vt[i][j] = a * c + b * d;
          Vt[k][j] = a * d - b * c;
        }
        if (k == 1) {
          Ut[0][0] = Ut[1][1] = a;
          Ut[1][0] = -b;
          Ut[0][1] = b;
          Ut[2][2] = 1.0;
        } else {
          for (int j = 0; j < 3; ++j) {
            double t = a * Ut[i][j] + b * Ut[k][j];
            Ut[k][j] = -b * Ut[i][j] + a * Ut[k][j];
            Ut[i][j] = t;
          }
        }
      }
    double scale = 1.0 / sqrt(Vt[0][0] * Vt[0][0] + Vt[0][1] * Vt[0][1] + Vt[0][2] * Vt[0][2]);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 3; ++j) Vt[i][j] *= scale;
    Vt[2][0] = Vt[0][1] * Vt[1][2] - Vt[0][2] * Vt[1][1];
    Vt[2][1] = Vt[0][2] * Vt[1][0] - Vt[0][0] * Vt[1][2];
    Vt[2][2] = Vt[0][0] * Vt[1][1] - Vt[0][1] * Vt[1][0];

    int c0a = 0, c0b = 0, c1a = 0, c1b = 0;
    for (int pt = 0; pt < 5; ++pt) {
      double a0 = q[pt][0], a1 = q[pt][1], b0 = qp[pt][0], b1 = qp[pt][1];
      double Vx0 = a0 * Vt[0][0] + a1 * Vt[0][1] + Vt[0][2];
      double Vx2 = a0 * Vt[2][0] + a1 * Vt[2][1] + Vt[2][2];
      double Ux1 = b0 * Ut[1][0] + b1 * Ut[1][1] + Ut[1][2];
      double Ux2 = b0 * Ut[2][0] + b1 * Ut[2][1] + Ut[2][2];
      double d1 = Vx0 * Ux2 + Vx2 * Ux1;
      double d2 = -Vx0 * Ux2 + Vx2 * Ux1;
      if (-Ux1 / d1 > 0.0) ++c0a;
      if (Vx0 / d1 > 0.0) ++c0b;
      if (-Ux1 / d2 > 0.0) ++c1a;
      if (-Vx0 / d2 > 0.0) ++c1b;
    }
    int c0 = c0a + c0b, c1 = c1a + c1b;
    double rs, ts; /* sign of the (U0 V1 - U1 V0) part of R, sign of t */
    if (c0 == 10) { rs = 1.0; ts = 1.0; }
    else if (c0 == 0) { rs = 1.0; ts = -1.0; }
    else if (c1 == 10) { rs = -1.0; ts = 1.0; }
    else if (c1 == 0) { rs = -1.0; ts = -1.0; }
    else continue;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j)
        P[kept][4 * i + j] =
            rs * (Ut[0][i] * Vt[1][j] - Ut[1][i] * Vt[0][j]) + Ut[2][i] * Vt[2][j];
      P[kept][4 * i + 3] = ts * Ut[2][i];
    }
    if (m > kept) memcpy(E[kept], E[m], 9 * sizeof(double));
    ++kept;
  }
  return kept;
}

static void gather_set(const double* x1, const double* x2, const int32_t* set, double q[5][2],
                       double qp[5][2]) {
  for (int i = 0; i < 5; ++i) { /* ref: kernel_functions.cu:284-300 (SelectSubset) */
    int idx = set[i];
    q[i][0] = x1[2 * idx]; q[i][1] = x1[2 * idx + 1];
    qp[i][0] = x2[2 * idx]; qp[i][1] = x2[2 * idx + 1];
  }
}

void tv5o_solve_sets(const double* x1, const double* x2, const int32_t* sets, int H,
                     int with_cheirality, double* E_list, double* P_list, int32_t* n_roots,
                     int32_t* n_valid) {
  for (int h = 0; h < H; ++h) {
    double q[5][2], qp[5][2], E[10][9], P[10][12];
    memset(E, 0, sizeof(E));
    memset(P, 0, sizeof(P));
    gather_set(x1, x2, sets + 5 * (size_t)h, q, qp);
    int nr = tv5o_solve5(q, qp, E, 0);
    int nv = nr;
    if (with_cheirality) nv = tv5o_cheirality(q, qp, E, nr, P);
    for (int m = nv; m < 10; ++m) { memset(E[m], 0, sizeof(E[m])); memset(P[m], 0, sizeof(P[m])); }
    n_roots[h] = nr;
    n_valid[h] = nv;
    memcpy(E_list + 90 * (size_t)h, E, sizeof(E));
    if (P_list) memcpy(P_list + 120 * (size_t)h, P, sizeof(P));
  }
}

/* ref: kernel_functions.cu:140-226 (EstimateProjectionMatrix<5>; with_cheirality == 0 gives
 * :53-135 EstimateEssentialMatrix<5>) and essential_matrix.cu:252 (first max over threads). */
int tv5o_ransac(const double* x1, const double* x2, int N, const int32_t* sets, int n_threads,
                int iters, int n_pre, int n_full, double thr, int with_cheirality, double E[9],
                double P[12], int32_t* best_set, int32_t* best_root, uint8_t* mask) {
  (void)N;
  int global_best = 0, gset = -1, groot = -1;
  memset(E, 0, 9 * sizeof(double));
  memset(P, 0, 12 * sizeof(double));
  for (int t = 0; t < n_threads; ++t) {
    int thread_best = 0, tset = -1, troot = -1;
    double tE[9], tP[12];
    memset(tE, 0, sizeof(tE));
    memset(tP, 0, sizeof(tP));
    for (int it = 0; it < iters; ++it) {
      int h = t * iters + it;
      double q[5][2], qp[5][2], Es[10][9], Ps[10][12];
      memset(Ps, 0, sizeof(Ps));
      gather_set(x1, x2, sets + 5 * (size_t)h, q, qp);
      int n = tv5o_solve5(q, qp, Es, 0);
      if (with_cheirality) n = tv5o_cheirality(q, qp, Es, n, Ps);
      if (n == 0) continue; /* documented divergence: no stale-slot scoring */
      int sub_best = 0, sub_idx = 0;
      for (int j = 0; j < n; ++j) {
        int32_t c;
        tv5o_score(x1, x2, n_pre, Es[j], 1, thr, &c, 0);
        if (c > sub_best) { sub_best = c; sub_idx = j; }
      }
      int32_t full;
      tv5o_score(x1, x2, n_full, Es[sub_idx], 1, thr, &full, 0);
      if (full > thread_best) {
        thread_best = full; tset = h; troot = sub_idx;
        memcpy(tE, Es[sub_idx], sizeof(tE));
        memcpy(tP, Ps[sub_idx], sizeof(tP));
      }
    }
    if (thread_best > global_best) { /* std::max_element: first maximum wins */
      global_best = thread_best; gset = tset; groot = troot;
      memcpy(E, tE, sizeof(tE));
      memcpy(P, tP, sizeof(tP));
    }
  }
  if (best_set) *best_set = gset;
  if (best_root) *best_root = groot;
  if (mask) {
    if (gset >= 0) { int32_t c; tv5o_score(x1, x2, n_full, E, 1, thr, &c, mask); }
    else memset(mask, 0, (size_t)n_full);
  }
  return global_best;
}

/* ------------------------------------------------------------------------------------------
 * Decomposition E = U diag(1,1,0) V^T and IRLS refinement.  ref: polish_E.cu.
 * ---------------------------------------------------------------------------------------- */
static void givens_from(double a, double b, double* c, double* s) { /* ref: polish_E.cu:160-165 */
  double sc = sqrt(a * a + b * b);
  *c = a / sc;
  *s = b / sc;
}

static void decompose_core(double E[9], double cs[10]) { /* ref: polish_E.cu:147-226 / :246-330 */
  double cx, cy, cz, cu, cv, sx, sy, sz, su, sv, t;
  givens_from(E[0], -E[3], &cz, &sz);               /* rows 0,1: eliminate E[1][0] */
  for (int j = 0; j < 3; ++j) { t = E[j] * cz - E[3 + j] * sz; E[3 + j] = E[j] * sz + E[3 + j] * cz; E[j] = t; }
  givens_from(E[0], -E[6], &cy, &sy);               /* rows 0,2: eliminate E[2][0] */
  for (int j = 0; j < 3; ++j) { t = E[j] * cy - E[6 + j] * sy; E[6 + j] = E[j] * sy + E[6 + j] * cy; E[j] = t; }
  givens_from(E[4], -E[7], &cx, &sx);               /* rows 1,2: eliminate E[2][1] */
  for (int j = 1; j < 3; ++j) E[3 + j] = E[3 + j] * cx - E[6 + j] * sx;
  givens_from(E[4], -E[5], &cu, &su);               /* columns 1,2: eliminate E[1][2] */
  E[2] = su * E[1] + cu * E[2];
  givens_from(E[0], -E[2], &cv, &sv);               /* columns 0,2: eliminate E[0][2] */
  cs[0] = cx; cs[1] = sx; cs[2] = cy; cs[3] = sy; cs[4] = cz; cs[5] = sz; cs[6] = cu; cs[7] = su; cs[8] = cv; cs[9] = sv;
}

void tv5o_decompose_uv(double E[9], double U[9], double V[9]) {
  double k[10];
  decompose_core(E, k);
  const double cx = k[0], sx = k[1], cy = k[2], sy = k[3], cz = k[4], sz = k[5], cu = k[6], su = k[7], cv = k[8], sv = k[9];
  U[0] = cy * cz;  U[1] = -cz * sx * sy + cx * sz; U[2] = cx * cz * sy + sx * sz;
  U[3] = -cy * sz; U[4] = cx * cz + sx * sy * sz;  U[5] = cz * sx - cx * sy * sz;
  U[6] = -sy;      U[7] = -cy * sx;                U[8] = cx * cy;
  V[0] = cv;       V[1] = 0.0; V[2] = sv;
  V[3] = -su * sv; V[4] = cu;  V[5] = cv * su;
  V[6] = -cu * sv; V[7] = -su; V[8] = cu * cv;
}

void tv5o_decompose_angles(double E[9], double par[5]) {
  double k[10];
  decompose_core(E, k);
  for (int i = 0; i < 5; ++i) par[i] = atan2(k[2 * i + 1], k[2 * i]);
}

static void rot_cols(double M[9], int c1, int c2, double angle) { /* ref: polish_E.cu:128-145 (Gright) */
  double c = cos(angle), s = sin(angle);
  for (int i = 0; i < 3; ++i) {
    double t = M[3 * i + c1] * c - M[3 * i + c2] * s;
    M[3 * i + c2] = M[3 * i + c1] * s + M[3 * i + c2] * c;
    M[3 * i + c1] = t;
  }
}

static void solve5(double A[5][5], double b[5]) { /* ref: polish_E.cu:340-448 (solve_5x5) */
  for (int row = 0; row < 5; ++row) {
    int mr = row;
    double mv = fabs(A[row][row]);
    for (int i = row + 1; i < 5; ++i) if (fabs(A[i][row]) > mv) { mv = fabs(A[i][row]); mr = i; }
    if (mr != row) {
      for (int j = row; j < 5; ++j) { double t = A[row][j]; A[row][j] = A[mr][j]; A[mr][j] = t; }
      double t = b[row]; b[row] = b[mr]; b[mr] = t;
    }
    for (int i = row + 1; i < 5; ++i) {
      double f = A[i][row] / A[row][row];
      for (int j = row + 1; j < 5; ++j) A[i][j] -= f * A[row][j];
      b[i] -= f * b[row];
    }
  }
  for (int i = 4; i >= 0; --i) {
    for (int j = i + 1; j < 5; ++j) b[i] -= A[i][j] * b[j];
    b[i] /= A[i][i];
  }
}

void tv5o_optimise(double E[9], const double* x1, const double* x2, int n, double delta,
                   double alpha, int max_reps) {
  double U[9], V[9];
  tv5o_decompose_uv(E, U, V);
  for (int rep = 0;; ++rep) {
    double g[5] = {0, 0, 0, 0, 0}, H[5][5];
    memset(H, 0, sizeof(H));
    for (int k = 0; k < n; ++k) {
      double p[3], q[3], J[5];
      for (int j = 0; j < 3; ++j) {
        p[j] = x1[2 * k] * V[j] + x1[2 * k + 1] * V[3 + j] + V[6 + j];
        q[j] = x2[2 * k] * U[j] + x2[2 * k + 1] * U[3 + j] + U[6 + j];
      }
      double e = p[0] * q[0] + p[1] * q[1];
      double w = (fabs(e) < delta) ? 1.0 : alpha * delta / fabs(e);
      J[0] = -p[1] * q[2]; J[1] = -p[0] * q[2]; J[2] = p[1] * q[0] - p[0] * q[1];
      J[3] = -p[2] * q[1]; J[4] = -p[2] * q[0];
      for (int i = 0; i < 5; ++i) {
        g[i] += J[i] * -e * w;
        for (int j = 0; j < 5; ++j) H[i][j] += w * J[i] * J[j];
      }
    }
    double mag = 0.0;
    for (int i = 0; i < 5; ++i) mag += g[i] * g[i];
    if (mag < 1e-20) break;
    for (int i = 0; i < 3; ++i)                       /* ref: :60-66 (Eprod) */
      for (int j = 0; j < 3; ++j) E[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1];
    if (rep == max_reps) break;
    solve5(H, g);
    rot_cols(U, 0, 1, g[2]); rot_cols(U, 0, 2, g[1]); rot_cols(U, 1, 2, g[0]);   /* ref: :450-472 (update) */
    rot_cols(V, 1, 2, g[3]); rot_cols(V, 0, 2, g[4]);
  }
}
