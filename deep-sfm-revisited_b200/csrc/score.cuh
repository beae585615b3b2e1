// score.cuh — Sampson inlier scoring of hypotheses against correspondences.
//
// Two scorers:
//   * sampson_exact(): float64, the reference's exact operation order
//     (RANSAC_FiveP/essential_matrix/kernel_functions.cu:231-264 as compiled by nvcc 12.9 for
//     sm_100a) — bit-identical inlier decisions.
//   * score_bounds kernel: float32 guard-band scorer on packed FFMA2 (fma.rn.f32x2), the
//     roofline kernel.  For each (hypothesis, point) it decides "surely an outlier under every
//     rounding" and accumulates out[m]; then exact count[m] <= n - out[m].  The two-sided
//     variant (tv5_score_bounds) also accumulates notin[m] = #{not surely an inlier}, giving
//     n - notin[m] <= exact count[m] as well.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tv5 {

// ------------------------------------------------------------------------------------------
// exact float64 decision (explicit _rn intrinsics: no re-association, no other contraction)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double sampson_err_exact(const double (&E)[9], double x1, double y1,
                                                    double x2, double y2) {
  const double Ex0 = __dadd_rn(__fma_rn(E[1], y1, __dmul_rn(E[0], x1)), E[2]);
  const double Ex1 = __dadd_rn(__fma_rn(E[4], y1, __dmul_rn(E[3], x1)), E[5]);
  const double Ex2 = __dadd_rn(__fma_rn(E[7], y1, __dmul_rn(E[6], x1)), E[8]);
  const double tE0 = __dadd_rn(__fma_rn(E[3], y2, __dmul_rn(E[0], x2)), E[6]);
  const double tE1 = __dadd_rn(__fma_rn(E[4], y2, __dmul_rn(E[1], x2)), E[7]);
  const double num = __dadd_rn(Ex2, __fma_rn(y2, Ex1, __dmul_rn(x2, Ex0)));
  const double den = __fma_rn(tE1, tE1, __fma_rn(tE0, tE0, __fma_rn(Ex0, Ex0, __dmul_rn(Ex1, Ex1))));
  return fabs(__ddiv_rn(num, __dsqrt_rn(den)));
}

__device__ __forceinline__ bool sampson_inlier_exact(const double (&E)[9], double x1, double y1,
                                                     double x2, double y2, double thr) {
  return sampson_err_exact(E, x1, y1, x2, y2) <= thr;  // NaN -> outlier, as the reference
}

// ------------------------------------------------------------------------------------------
// float32 guard-band scorer
// ------------------------------------------------------------------------------------------
// Hypothesis record: E^ = E/||E||_F;  rows 0,1 as is, row 2 scaled by s (g), plus the two unscaled
// row-2 entries needed by E^T x2.  s = sqrt(1-c)/thr is a per-image-pair constant (BandConst).
struct __align__(16) Hyp32 {
  float e00, e01, e02, e10, e11, e12, g0, g1, g2, e20, e21, pad;
};
static_assert(sizeof(Hyp32) == 48, "Hyp32 layout");

// Point-pair record (two consecutive correspondences p, q packed lane-wise for f32x2 math):
//   [x1p x1q y1p y1q] [x2p x2q y2p y2q] [s x2p  s x2q  s y2p  s y2q]
struct __align__(16) PointPair32 {
  float2 x1, y1, x2, y2, x2s, y2s;
};
static_assert(sizeof(PointPair32) == 48, "PointPair32 layout");

// Per image-pair constants of the bound (DESIGN.md "guard band").  With n' = num/thr, d = den
// (unit-norm E^), B the rounding bound on n' and sqrt(d):
//   surely outlier  <=  d + K - (1-c) n'^2 < 0        surely inlier  <=  d - K - (1+c) n'^2 >= 0
// The kernel evaluates n'' = sqrt(1-c) n' directly (the factor is folded into the point and
// hypothesis records), so the first test is  (d + K) - n''^2 < 0 : one FFMA2.
struct BandConst {
  float K;       // seed of the sum of squares
  float ratio;   // (1+c)/(1-c), two-sided variant only
  float two_K;   // 2K,          two-sided variant only
  float pad;
};

struct HypRegs {  // one hypothesis; the coefficients enter FFMA2 as scalar-broadcast operands
  float2 e00, e01, e02, e10, e11, e12, g0, g1, g2, e20, e21;
};

__device__ __forceinline__ float2 dup(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }

__device__ __forceinline__ void load_hyp(HypRegs& h, const Hyp32& s) {
  h.e00 = dup(s.e00); h.e01 = dup(s.e01); h.e02 = dup(s.e02);
  h.e10 = dup(s.e10); h.e11 = dup(s.e11); h.e12 = dup(s.e12);
  h.g0 = dup(s.g0); h.g1 = dup(s.g1); h.g2 = dup(s.g2);
  h.e20 = dup(s.e20); h.e21 = dup(s.e21);
}

// Two (hypothesis, point) evaluations.  One-sided form: 17 packed FP32 instructions (34 FP32
// lane-operations per evaluation pair... i.e. 17 per evaluation) + 2 sign-bit adds.
// TWO_SIDED adds the lower bound (3 more packed instructions, 2 more sign-bit adds).
template <bool TWO_SIDED>
__device__ __forceinline__ void eval_pair(const HypRegs& h, const PointPair32& p, float2 K,
                                          float2 neg_ratio, float2 neg_twoK, uint32_t& notin,
                                          uint32_t& out, bool both_lanes = true) {
  const float2 Ex0 = __ffma2_rn(h.e00, p.x1, __ffma2_rn(h.e01, p.y1, h.e02));
  const float2 Ex1 = __ffma2_rn(h.e10, p.x1, __ffma2_rn(h.e11, p.y1, h.e12));
  const float2 Ex2 = __ffma2_rn(h.g0, p.x1, __ffma2_rn(h.g1, p.y1, h.g2));
  const float2 tE0 = __ffma2_rn(h.e00, p.x2, __ffma2_rn(h.e10, p.y2, h.e20));
  const float2 tE1 = __ffma2_rn(h.e01, p.x2, __ffma2_rn(h.e11, p.y2, h.e21));
  const float2 n = __ffma2_rn(p.x2s, Ex0, __ffma2_rn(p.y2s, Ex1, Ex2));  // sqrt(1-c) num / thr
  float2 dK = __ffma2_rn(Ex0, Ex0, K);                                    // d + K
  dK = __ffma2_rn(Ex1, Ex1, dK);
  dK = __ffma2_rn(tE0, tE0, dK);
  dK = __ffma2_rn(tE1, tE1, dK);
  const float2 hi = __ffma2_rn(neg2(n), n, dK);                           // < 0: surely an outlier
  out += __float_as_uint(hi.x) >> 31;
  if (both_lanes) out += __float_as_uint(hi.y) >> 31;
  if (TWO_SIDED) {
    const float2 nn = __fmul2_rn(n, n);
    const float2 lo = __ffma2_rn(nn, neg_ratio, __fadd2_rn(dK, neg_twoK)); // < 0: not surely an inlier
    notin += __float_as_uint(lo.x) >> 31;
    if (both_lanes) notin += __float_as_uint(lo.y) >> 31;
  }
}

}  // namespace tv5
