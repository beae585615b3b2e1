// solve5_coop.cuh — warp-cooperative front end of the five-point solver.
//
// One warp owns 32 minimal sets (lane = set) for the per-set phases (null space, determinant,
// roots, E, pose).  The two phases that would need a 10x20 float64 matrix per thread — building
// the ten cubic constraints and the pivoted Gauss-Jordan elimination — are done cooperatively
// instead: three sets at a time, ten lanes per set, ONE MATRIX ROW PER LANE, all 20 coefficients
// of the row in registers, pivot search and pivot-row broadcast by warp shuffles.  Nothing of the
// matrix ever touches local memory; per warp only the null-space bases (36 doubles per set) and
// the six reduced rows that survive the elimination (60 doubles per set) go through shared
// memory.  (Replaces the per-thread build_constraints + eliminate of solve5.cuh on the hot path;
// same mathematics: Nister's 10x20 system, partial pivoting.)
#pragma once
#include "solve5.cuh"

namespace tv5 {

// 1 = pivot row broadcast through shared memory with 128-bit accesses instead of shuffles: half the
// LSU wavefronts, but the store -> sync -> load latency per pivot costs more (solve 3.32 -> 3.43 ms)
#ifndef TV5_COOP_SMEM_BROADCAST
#define TV5_COOP_SMEM_BROADCAST 0
#endif
// 1 = pivot search by one redux.sync (max over the group's 10 lanes) instead of four dependent
// rotation shuffles.  Measured on B200: MUCH slower (solver 3.32 -> 4.14 ms per 1.05 M sets) — redux.sync
// with three different member masks in one warp is evidently executed mask by mask.  Off.
#ifndef TV5_COOP_REDUX
#define TV5_COOP_REDUX 0
#endif
// 1 = every lane takes the reciprocal of its own pivot candidate while the pivot search is in flight
// and the pivot lane's reciprocal is fetched by one shuffle, instead of fetching the pivot and
// inverting it afterwards: same value, same operation (bit-identical), a shorter dependent chain.
// Measured on B200: no gain (solver 3.11 -> 3.16 ms per 1.05 M sets).  Off.
#ifndef TV5_COOP_SPEC_RCP
#define TV5_COOP_SPEC_RCP 0
#endif
constexpr int kCoopBasisDoubles = 36;
constexpr int kCoopRowsDoubles = 60;
#ifndef TV5_COOP_JAM
#define TV5_COOP_JAM 1
#endif
constexpr int kCoopJam = TV5_COOP_JAM;      // rounds carried through the elimination together (2 measured slower: spills)
#ifndef TV5_COOP_STRIDE
#define TV5_COOP_STRIDE 33
#endif
constexpr int kCoopStride = TV5_COOP_STRIDE;  // padded lane stride: conflict-free for fixed-set/varying-element access too
constexpr int kCoopPointDoubles = 20;

// sB[e][lane]: basis coefficient e = k*9 + c of the lane's set (k: unknown w,x,y,1; c = 3i+j)
// sR[e][lane]: e = r*10 + j, r = 0..5 <-> pivot columns 4..9, j <-> matrix columns 10..19
// One lane's constraint row r of set sw:  sum_t Q_t * L_t  (Q quadratic, L linear in (w,x,y,1))
//      det row:      Q_t = E(a,1)E(b,2) - E(b,1)E(a,2),  L_t = E(t,0),  (a,b) = (t+1,t+2) mod 3
//      row (i,j):    Q_t = sum_p E(i,p)E(t,p) - [t==i] tr/2,  L_t = E(t,j)
template <int S>
__device__ __forceinline__ void coop_build_row(const double (*sB)[S], int sw, int gb, bool isdet,
                                               int ri, int rj, double (&row)[20]) {
  const unsigned FULL = 0xffffffffu;
  double Q[3][10];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
#pragma unroll
    for (int u = 0; u < 10; ++u) Q[t][u] = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int a = (t + 1) % 3, b = (t + 2) % 3;
      const int ia = isdet ? (p == 0 ? 3 * a + 1 : 3 * b + 1) : 3 * ri + p;
      const int ib = isdet ? (p == 0 ? 3 * b + 2 : 3 * a + 2) : 3 * t + p;
      const double wt = isdet ? (p == 0 ? 1.0 : (p == 1 ? -1.0 : 0.0)) : 1.0;
      double la[4], lb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { la[k] = wt * sB[k * 9 + ia][sw]; lb[k] = sB[k * 9 + ib][sw]; }
      quad_acc(Q[t], la, lb);
    }
  }
  // trace of E E^T = G(0,0) + G(1,1) + G(2,2): the lanes of rows (0,0), (1,0), (2,0) hold them
#pragma unroll
  for (int u = 0; u < 10; ++u) {
    const double diag = ri == 0 ? Q[0][u] : (ri == 1 ? Q[1][u] : Q[2][u]);
    const double tr = __shfl_sync(FULL, diag, gb + 1) + __shfl_sync(FULL, diag, gb + 4) +
                      __shfl_sync(FULL, diag, gb + 7);
    const double h = isdet ? 0.0 : 0.5 * tr;
    Q[0][u] -= ri == 0 ? h : 0.0;
    Q[1][u] -= ri == 1 ? h : 0.0;
    Q[2][u] -= ri == 2 ? h : 0.0;
  }
#pragma unroll
  for (int u = 0; u < 20; ++u) row[u] = 0.0;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int il = 3 * t + rj;  // det row: rj = 0 -> E(t,0)
    double L[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) L[k] = sB[k * 9 + il][sw];
    cubic_acc(row, Q[t], L);
  }
}

// sB[e][lane]: basis coefficient e = k*9 + c of the lane's set (k: unknown w,x,y,1; c = 3i+j)
// sR[e][lane]: e = r*10 + j, r = 0..5 <-> pivot columns 4..9, j <-> matrix columns 10..19
// kCoopJam rounds (x 3 sets) are carried through the elimination together.
template <int S>
__device__ __forceinline__ void coop_constraints_eliminate(const double (*sB)[S], double (*sR)[S],
                                                           int* sOk, int lane, int n_sets = 32) {
  const unsigned FULL = 0xffffffffu;
#if TV5_COOP_SMEM_BROADCAST
  __shared__ __align__(16) double sP[3][22];           // pivot row of each group (stride 22: distinct banks)
#endif
  const int g = lane < 30 ? lane / 10 : 2;            // lanes 30, 31 shadow group 2's shuffles
  const int r = lane < 30 ? lane - 10 * g : lane - 20;  // rows 10, 11 do not take part
  const int gb = 10 * g;
  const bool rowlane = r < 10;
  const unsigned gmask = lane < 30 ? (0x3ffu << gb) : 0xC0000000u;   // lanes 30, 31 reduce among themselves
  (void)gmask;
  const bool isdet = (r == 0) || !rowlane;
  const int ri = isdet ? 0 : (r - 1) / 3;
  const int rj = isdet ? 0 : (r - 1) - 3 * ri;

  const int n_rounds = (n_sets + 2) / 3;   // three sets per round
  for (int round2 = 0; round2 < (n_rounds + kCoopJam - 1) / kCoopJam; ++round2) {
    int sw[kCoopJam];
    bool active[kCoopJam];
    double row[kCoopJam][20];
#pragma unroll
    for (int v = 0; v < kCoopJam; ++v) {
      const int sw_raw = (kCoopJam * round2 + v) * 3 + g;
      sw[v] = sw_raw < n_sets ? sw_raw : n_sets - 1;
      active[v] = rowlane && sw_raw < n_sets;
      coop_build_row(sB, sw[v], gb, isdet, ri, rj, row[v]);
    }
    // ---- Gauss-Jordan with partial pivoting on columns 0..9, rows spread over the group's lanes
    int pivcol[kCoopJam];
    bool bad[kCoopJam];
    double pivval[kCoopJam];
#pragma unroll
    for (int v = 0; v < kCoopJam; ++v) { pivcol[v] = -1; bad[v] = false; pivval[v] = 1.0; }
#pragma unroll
    for (int c = 0; c < 10; ++c) {
      // partial pivoting: all-reduce (max is idempotent) over the 10-lane ring by rotations
      // 1, 2, 4, 8 of a packed key = float magnitude bits (low 5 bits replaced by the lane id)
      unsigned key[kCoopJam];
#pragma unroll
      for (int v = 0; v < kCoopJam; ++v) {
        const float mag = fabsf((float)row[v][c]);
        key[v] = (rowlane && pivcol[v] < 0 && mag == mag) ? ((__float_as_uint(mag) & ~31u) | (unsigned)lane) : 0u;
      }
#if TV5_COOP_SPEC_RCP
      double inv_self[kCoopJam];
#pragma unroll
      for (int v = 0; v < kCoopJam; ++v) inv_self[v] = __drcp_rn(row[v][c]);   // independent of the search below
#endif
#if TV5_COOP_REDUX
#pragma unroll
      for (int v = 0; v < kCoopJam; ++v) key[v] = __reduce_max_sync(gmask, key[v]);
#else
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const int src = rowlane ? gb + (r + o) % 10 : gb;
#pragma unroll
        for (int v = 0; v < kCoopJam; ++v) {
          const unsigned other = __shfl_sync(FULL, key[v], src);
          key[v] = other > key[v] ? other : key[v];
        }
      }
#endif
      // The pivot row is NOT normalised while eliminating: with multiplier g = row[c] / pivot (0 on
      // the pivot lane itself) every lane does  row[j] -= g * pivot_row[j]  — one fma per column and
      // no select; the surviving rows are divided by their pivot once, at the end.
      int bl[kCoopJam];
      double g_mul[kCoopJam];
#pragma unroll
      for (int v = 0; v < kCoopJam; ++v) {
        bl[v] = (key[v] >> 5) ? (int)(key[v] & 31u) : gb;
        const float best = __uint_as_float(key[v] & ~31u);
        bad[v] = bad[v] || !(best > 0.f) || !(best < 3.0e38f);
#if TV5_COOP_SPEC_RCP
        const double ipiv = __shfl_sync(FULL, inv_self[v], bl[v]);
        g_mul[v] = (lane == bl[v]) ? 0.0 : -(row[v][c] * ipiv);
        if (lane == bl[v]) { pivcol[v] = c; pivval[v] = row[v][c]; }
#else
        const double piv = __shfl_sync(FULL, row[v][c], bl[v]);
        g_mul[v] = (lane == bl[v]) ? 0.0 : -(row[v][c] * __drcp_rn(piv));
        if (lane == bl[v]) { pivcol[v] = c; pivval[v] = piv; }
#endif
      }
#if TV5_COOP_SMEM_BROADCAST
      // pivot row to the group through shared memory, two columns per 128-bit access (half the LSU
      // wavefronts of two 32-bit shuffles per column)
      static_assert(kCoopJam == 1, "shared-memory broadcast is written for one round in flight");
      if (lane == bl[0]) {
#pragma unroll
        for (int j2 = (c + 1) & ~1; j2 < 20; j2 += 2)
          *reinterpret_cast<double2*>(&sP[g][j2]) = make_double2(row[0][j2], row[0][j2 + 1]);
      }
      __syncwarp();
#pragma unroll
      for (int j2 = (c + 1) & ~1; j2 < 20; j2 += 2) {
        const double2 pv = *reinterpret_cast<const double2*>(&sP[g][j2]);
        if (j2 > c) row[0][j2] = fma(g_mul[0], pv.x, row[0][j2]);
        row[0][j2 + 1] = fma(g_mul[0], pv.y, row[0][j2 + 1]);
      }
      __syncwarp();
#else
#pragma unroll
      for (int j = c + 1; j < 20; ++j) {
#pragma unroll
        for (int v = 0; v < kCoopJam; ++v) row[v][j] = fma(g_mul[v], __shfl_sync(FULL, row[v][j], bl[v]), row[v][j]);
      }
#endif
    }
#pragma unroll
    for (int v = 0; v < kCoopJam; ++v) {
      const unsigned badmask = __ballot_sync(FULL, bad[v] && rowlane);
      __syncwarp();   // sR may alias sB: every lane of the group has finished reading the set's basis
      if (active[v]) {
        if (pivcol[v] >= 4) {
          const double sc = __drcp_rn(pivval[v]);
#pragma unroll
          for (int j = 0; j < 10; ++j) sR[(pivcol[v] - 4) * 10 + j][sw[v]] = row[v][10 + j] * sc;
        }
        if (r == 0) sOk[sw[v]] = ((badmask >> gb) & 0x3ffu) ? 0 : 1;
      }
    }
  }
}

// 3x3 polynomial matrix from the six reduced rows (see hidden_matrix in solve5.cuh)
template <int S>
__device__ __forceinline__ void hidden_matrix_from_rows(const double (*sR)[S], int lane,
                                                        double (&Bp)[3][3][5]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double a[10], b[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) { a[j] = sR[(2 * r) * 10 + j][lane]; b[j] = sR[(2 * r + 1) * 10 + j][lane]; }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int o = 3 * c;
      Bp[r][c][0] = a[o + 2];
      Bp[r][c][1] = a[o + 1] - b[o + 2];
      Bp[r][c][2] = a[o] - b[o + 1];
      Bp[r][c][3] = -b[o];
      Bp[r][c][4] = 0.0;
    }
    Bp[r][2][0] = a[9];
    Bp[r][2][1] = a[8] - b[9];
    Bp[r][2][2] = a[7] - b[8];
    Bp[r][2][3] = a[6] - b[7];
    Bp[r][2][4] = -b[6];
  }
}

// Whole solve for the lane's set.  All 32 lanes of the warp must call this (lanes without a set
// pass valid = false and still take part in the cooperative phase).  Register pressure is kept
// down for the root finder (66 chain coefficients in registers) by parking everything else in
// shared memory meanwhile: sQ = the five point pairs, sB = basis, sR = reduced rows.
// emit(j, E) is called for every solution kept (j = its index in the set's list) while E is still
// in registers.
struct NoEmit {
  __device__ __forceinline__ void operator()(int, const double (&)[9]) const {}
};
// reload(q, qp) (optional) re-gathers the five point pairs for the pose phase; without it they
// are parked in sQ during the root phase.  Re-gathering saves 5 KB of shared memory per warp
// (8 instead of 7 resident warps per SM).
struct NoReload {
  static constexpr bool kEnabled = false;
  __device__ __forceinline__ void operator()(double (&)[5][2], double (&)[5][2]) const {}
};
template <typename Emit = NoEmit, typename Reload = NoReload>
__device__ inline int solve_minimal_set_coop(bool valid, const double (&q_in)[5][2],
                                             const double (&qp_in)[5][2], bool with_cheirality,
                                             double* E_out, double* P_out, int* n_roots_out,
                                             double (*sB)[kCoopStride], double (*sR)[kCoopStride], double (*sQ)[kCoopStride],
                                             int* sOk, Emit emit = Emit(), Reload reload = Reload()) {
  const int lane = threadIdx.x & 31;
  *n_roots_out = 0;
  bool ok = valid;
#ifdef TV5_SOLVE_PROFILE
  long long t_prev = clock64();
#endif
  {
    double B[4][9];
    nullspace_basis(q_in, qp_in, B);
#pragma unroll
    for (int c = 0; c < 9; ++c) ok = ok && (fabs(B[3][c]) <= 1.0);  // false on NaN (degenerate set)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int c = 0; c < 9; ++c) sB[k * 9 + c][lane] = ok ? B[k][c] : (k == 3 ? 1.0 : 0.0);
    if (!Reload::kEnabled) {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        sQ[4 * i][lane] = q_in[i][0]; sQ[4 * i + 1][lane] = q_in[i][1];
        sQ[4 * i + 2][lane] = qp_in[i][0]; sQ[4 * i + 3][lane] = qp_in[i][1];
      }
    }
  }
  __syncwarp();
  TV5_TICK(0);
  coop_constraints_eliminate(sB, sR, sOk, lane);
  __syncwarp();
  TV5_TICK(1);
  if (!ok || !sOk[lane]) return 0;
  double roots[10];
  int nr;
  {
    double Bp[3][3][5];
    double poly[11];
    hidden_matrix_from_rows(sR, lane, Bp);
    hidden_determinant(Bp, poly);
    TV5_TICK(3);
    nr = real_roots_deg10(poly, roots);
    TV5_TICK(4);
  }
  *n_roots_out = nr;
  asm volatile("" ::: "memory");  // reload (do not keep alive) what was parked in shared memory
  double Bp[3][3][5], B[4][9], q[5][2], qp[5][2];
  hidden_matrix_from_rows(sR, lane, Bp);
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < 9; ++c) B[k][c] = sB[k * 9 + c][lane];
  if (Reload::kEnabled) {
    reload(q, qp);
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      q[i][0] = sQ[4 * i][lane]; q[i][1] = sQ[4 * i + 1][lane];
      qp[i][0] = sQ[4 * i + 2][lane]; qp[i][1] = sQ[4 * i + 3][lane];
    }
  }
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
    emit(nv, E);
    ++nv;
  }
  TV5_TICK(5);
  return nv;
}

}  // namespace tv5
