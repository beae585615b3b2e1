// solve5_split.cuh — the five-point solver as three kernels instead of one.
//
// The fused kernel (solve_sets) holds, per lane, the union of every phase's live state (255
// registers, 8 warps per SM) and runs the two per-root phases — Newton refinement and E + pose —
// as per-lane loops whose trip count is the warp's maximum root count (mean 4.6, maximum 10).
// Here each phase is its own launch at its own register budget and lane mapping:
//
//   solve_front   lane = set (warp-cooperative constraints + elimination as before): null space,
//                 10x20 system, 3x3 polynomial matrix, degree-10 determinant  -> record per set
//   solve_roots   thread = set: Sturm chain, root isolation                  -> (set, interval) list
//   solve_poses   thread = ROOT (compacted list, no divergence over the root count): bracketed
//                 Newton, E from the root, twisted pair + cheirality, float32 hypothesis record
//
// Between the kernels a set is a 96-double record in global memory (written coalesced from shared
// memory by solve_front, ~0.8 GB per 10^6 sets, read through L2).  Solutions are stored at their
// ROOT index (ascending w) and a per-set bit mask says which ones survived; hypothesis ids carry
// the root index, and the compacted index the C ABI promises is recovered by a popcount
// (finalize) or by compact_solutions (tv5_solve5).
#pragma once
#include "solve5_coop.cuh"

#ifndef TV5_POSES_PREFETCH
#define TV5_POSES_PREFETCH 1
#endif

namespace tv5 {

constexpr int kRecDoubles = 96;     // record per set
constexpr int kRecB = 0;            // [36] null-space basis, k*9 + c
constexpr int kRecBp = 36;          // [45] hidden-variable matrix, (r*3 + c)*5 + k
constexpr int kRecPoly = 82;        // [11] determinant polynomial (solve_front: raw; solve_roots: monic, scaled)
constexpr int kRecBack = 93;        // root of the scaled polynomial * back = w
constexpr int kRecOk = 94;          // 1.0 when the set produced a polynomial
// (offsets are even so that every field can be read with 16-byte loads)

struct RootEntry {                  // one isolated real root
  double lo, hi;                    // bracket with a sign change (or lo = root when exact != 0)
  int32_t set;                      // minimal set (hypothesis id) inside the image pair
  int32_t r_exact;                  // root index | exact << 8
};

// ---- solve_front: grid (ceil(H/32), B), block 32 --------------------------------------------
// sets_per_warp (32, 16 or 8): lanes >= sets_per_warp idle in the per-set phases and the elimination
// runs ceil(sets_per_warp / 3) rounds — fewer sets per warp shorten the critical path of small
// submissions (one 4096-set pair: 128 warps x 11 rounds -> 512 warps x 3 rounds).
// sB and sR may ALIAS (kFrontAlias: sR rows 0..35 = sB rows 0..35, one 60-row array): column s of the
// basis is dead once the round of set s has built its rows, which is before that round stores the
// set's reduced rows into the same column — provided the basis part of the record has already been
// written to global memory, which is done right after the null-space phase.  15.8 KB instead of
// 25.4 KB of shared memory per warp: 12 resident warps per SM (register limit) instead of 8.
template <int S, typename Gather>
__device__ __forceinline__ void solve_front_warp(bool valid, const Gather& gather, double* __restrict__ rec_warp,
                                                 int n_sets_here, double (*sB)[S],
                                                 double (*sR)[S], int* sOk, int sets_per_warp = 32) {
  const int lane = threadIdx.x & 31;
  bool ok = valid;
  {
    double q[5][2], qp[5][2], B[4][9];
    gather(q, qp);
    nullspace_basis(q, qp, B);
#pragma unroll
    for (int c = 0; c < 9; ++c) ok = ok && (fabs(B[3][c]) <= 1.0);  // false on NaN (degenerate set)
    if (lane < sets_per_warp) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 9; ++c) sB[k * 9 + c][lane] = ok ? B[k][c] : (k == 3 ? 1.0 : 0.0);
    }
  }
  __syncwarp();
  // basis part of the records (doubles 0..35), coalesced: lanes walk one record
  for (int sidx = 0; sidx < n_sets_here; ++sidx) {
    double* __restrict__ rec = rec_warp + (size_t)sidx * kRecDoubles;
    rec[lane] = sB[lane][sidx];
    if (lane < 4) rec[32 + lane] = sB[32 + lane][sidx];
  }
  coop_constraints_eliminate(sB, sR, sOk, lane, sets_per_warp);
  __syncwarp();
  ok = ok && (lane < sets_per_warp) && sOk[min(lane, sets_per_warp - 1)];
  {
    double Bp[3][3][5], poly[11];
    const int col = min(lane, sets_per_warp - 1);   // idle lanes read a valid column, write nothing
    hidden_matrix_from_rows(sR, col, Bp);
    hidden_determinant(Bp, poly);
    __syncwarp();                                    // idle lanes have read their neighbour's column
    if (lane < sets_per_warp) {
    // park Bp, poly and the flag in this lane's column of sR (its rows are consumed)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < 5; ++k) sR[(r * 3 + c) * 5 + k][lane] = Bp[r][c][k];
    sR[45][lane] = 0.0;
#pragma unroll
    for (int i = 0; i < 11; ++i) sR[46 + i][lane] = ok ? poly[i] : 0.0;
    sR[57][lane] = 0.0;
    sR[58][lane] = ok ? 1.0 : 0.0;
    }
  }
  __syncwarp();
  // coalesced copy-out of doubles 36..95: lanes walk one record
  for (int sidx = 0; sidx < n_sets_here; ++sidx) {
    double* __restrict__ rec = rec_warp + (size_t)sidx * kRecDoubles;
    rec[36 + lane] = sR[lane][sidx];
    const int e = 32 + lane;                       // sR rows 32..58, then padding up to double 95
    if (e < 60) rec[36 + e] = e < 59 ? sR[e][sidx] : 0.0;
  }
}

// ---- solve_roots: one thread per set ----------------------------------------------------------
// Returns the number of real roots n; alloc(n) (called once, when n > 0) returns where the set's n
// root entries go — they are written there directly, in ascending order, with `set` filled in.
template <typename Alloc>
__device__ inline int solve_roots_set(double* __restrict__ rec, int set, Alloc alloc) {
  if (!(rec[kRecOk] == 1.0)) return 0;
  double poly[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) poly[i] = rec[kRecPoly + i];
  FastChain s;
  double back, ilo[10], ihi[10];
  int ivlo[10];
  int ni = isolate_roots_deg10(poly, s, back, ilo, ihi, ivlo);
  if (ni < 0) {  // generic chain (rare): roots come back refined, in w
    double roots[10];
    ni = real_roots_deg10_generic(poly, roots);
    if (ni > 0) {
      RootEntry* __restrict__ dst = alloc(ni);
      for (int i = 0; i < ni; ++i) {
        RootEntry e;
        e.lo = e.hi = roots[i];
        e.set = set;
        e.r_exact = i | (1 << 8);
        dst[i] = e;
      }
    }
    rec[kRecBack] = 1.0;
    return ni;
  }
  if (ni > 0) {
    RootEntry* __restrict__ dst = alloc(ni);
    for (int i = 0; i < ni; ++i) {
      double lo = ilo[i], hi = ihi[i], flo = 0.0;
      int exact = 1;
      if (ivlo[i] >= 0) exact = prepare_bracket(s, lo, hi, ivlo[i], flo);
      RootEntry e;
      e.lo = lo;
      e.hi = exact ? lo : hi;
      e.set = set;
      e.r_exact = i | (exact << 8);
      dst[i] = e;
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) rec[kRecPoly + i] = s.c[0][i];
    rec[kRecBack] = back;
  }
  return ni;
}

// ---- solve_poses: one thread per isolated root --------------------------------------------------
// Returns true when the root gives a solution; E (and P when with_cheirality) are filled.
template <typename Gather>
__device__ inline bool solve_pose_root(const double* __restrict__ rec, const RootEntry& en, bool with_cheirality,
                                       const Gather& gather, double (&E)[9], double (&P)[12]) {
  const double2* __restrict__ rec2 = reinterpret_cast<const double2*>(rec);
#if TV5_POSES_PREFETCH
  // the basis and the hidden-variable matrix (first five 128-byte lines of the record) are needed only
  // after the Newton refinement: ask for them now, the refinement runs while they arrive
#pragma unroll
  for (int l = 0; l < 5; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 16 * l));
#endif
  double w;
  if (en.r_exact >> 8) {
    w = en.lo * rec[kRecBack];
  } else {
    double p[12], dq[11];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double2 v = rec2[kRecPoly / 2 + i];     // poly[0..10], back
      p[2 * i] = v.x; p[2 * i + 1] = v.y;
    }
    const double (&pp)[11] = *reinterpret_cast<const double (*)[11]>(&p[0]);
#pragma unroll
    for (int i = 1; i <= 10; ++i) dq[i - 1] = p[i] * i / 10.0;   // as fast_build (p is monic)
    dq[10] = 0.0;
    double f, d;
    eval_p_dp(pp, en.lo, f, d);
    w = newton_bracketed(pp, dq, en.lo, en.hi, f) * p[11];
  }
  double B[4][9], Bp[3][3][5];
  {
    double* Bf = &B[0][0];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const double2 v = rec2[kRecB / 2 + i];
      Bf[2 * i] = v.x; Bf[2 * i + 1] = v.y;
    }
    double* Pf = &Bp[0][0][0];
#pragma unroll
    for (int i = 0; i < 22; ++i) {
      const double2 v = rec2[kRecBp / 2 + i];
      Pf[2 * i] = v.x; Pf[2 * i + 1] = v.y;
    }
    Pf[44] = rec[kRecBp + 44];
  }
  if (!essential_from_root(B, Bp, w, E)) return false;
  if (with_cheirality) {
    double q[5][2], qp[5][2];
    gather(q, qp);
    if (!pose_from_essential(E, q, qp, P)) return false;
  }
  return true;
}

}  // namespace tv5
