// tv5.cu — kernels and C ABI of libtv5 (see include/tv5.h).  sm_100a only.
//
// Pipeline of one submission (B image pairs, everything stream-ordered, no host sync):
//   prep_norms      per-pair norm bounds; band_consts: guard-band constants per pair
//   prep_points     float64 [N,2] x2  ->  packed float32 point pairs
//   solve_sets      one minimal set per thread (solve5.cuh) -> E/P lists + float32 hypotheses
//   plan_tiles      per-pair guard-band constants, scoring tile table (single CTA)
//   score_bounds    float32 FFMA2 guard-band scorer (the roofline kernel), persistent CTAs,
//                   correspondence tiles staged in shared memory by cp.async.bulk (TMA)
//   pick_top        hypothesis with the largest upper bound  -> exact_counts (float64)
//   pick_rest       hypotheses whose upper bound reaches that exact count -> exact_counts
//   finalize        first-max selection (count desc, hypothesis id asc), E/P/mask output
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <new>
#include <vector>

#include "solve5.cuh"
#include "solve5_coop.cuh"
#include "solve5_split.cuh"
#include "polish.cuh"
#include "flow_points.cuh"
#include "plane_sweep.cuh"
#include "tv5_internal.h"

namespace tv5 {

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// prep_norms: grid (ceil(max_pp / 256), B) — max ||(x,1)||^2 per image pair, non-finite flag
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_norms(const PairDesc* __restrict__ desc,
                                                  PairState* __restrict__ state) {
  const PairDesc d = desc[blockIdx.y];
  const int npp = (d.n + 1) >> 1;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  float r1 = 0.f, r2 = 0.f;
  int bad = 0;
  if (g < npp) {
    const int p = 2 * g, q = min(2 * g + 1, d.n - 1);
    const double2 a1 = reinterpret_cast<const double2*>(d.x1)[p];
    const double2 b1 = reinterpret_cast<const double2*>(d.x1)[q];
    const double2 a2 = reinterpret_cast<const double2*>(d.x2)[p];
    const double2 b2 = reinterpret_cast<const double2*>(d.x2)[q];
    const double s1a = a1.x * a1.x + a1.y * a1.y, s1b = b1.x * b1.x + b1.y * b1.y;
    const double s2a = a2.x * a2.x + a2.y * a2.y, s2b = b2.x * b2.x + b2.y * b2.y;
    // !(x < big) catches NaN and Inf as well (fmax alone would drop a NaN operand)
    bad = !(s1a < 1e30) || !(s1b < 1e30) || !(s2a < 1e30) || !(s2b < 1e30);
    r1 = __double2float_ru(fmax(s1a, s1b) + 1.0);
    r2 = __double2float_ru(fmax(s2a, s2b) + 1.0);
  }
  r1 = warp_max(bad ? 0.f : r1);
  r2 = warp_max(bad ? 0.f : r2);
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&state[blockIdx.y].r1_bits, __float_as_uint(r1));
    atomicMax(&state[blockIdx.y].r2_bits, __float_as_uint(r2));
    if (bad) atomicOr(&state[blockIdx.y].nonfinite, 1);
  }
}

// ------------------------------------------------------------------------------------------
// band_consts: one thread per image pair.  Guard-band constants (DESIGN.md "guard band"):
//   computed n' = num/thr has |error| <= Bn = 8.2 u R1 R2 / thr   (u = 2^-24, ||E^||_F = 1,
//   R1 = max||(x1,1)||, R2 = max||(x2,1)||), sqrt(d) has |error| <= Bd = 18 u max(R1,R2).
//   With B = 1.02 (Bn + Bd):  sure-in  <=  (|n'|+B)^2 <= d,  sure-out  <=  (|n'|-B)^2 > d, and
//   2 B |n'| <= c n'^2 + B^2/c  gives the multiply-add forms used by eval_pair.
// ------------------------------------------------------------------------------------------
__global__ void band_consts(const PairDesc* __restrict__ desc, PairState* __restrict__ state, int B,
                            double thr, int allow_fast) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  PairState& s = state[b];
  const PairDesc& d = desc[b];
  const double u = 5.9604644775390625e-8;
  const double R1 = sqrt((double)__uint_as_float(s.r1_bits));
  const double R2 = sqrt((double)__uint_as_float(s.r2_bits));
  const double Bn = 8.2 * u * R1 * R2 / thr;
  const double Bd = 18.0 * u * fmax(R1, R2);
  const double Bt = 1.02 * (Bn + Bd);
  const int fast = allow_fast && !s.nonfinite && d.n_pre == d.n_full && (R1 < 1024.0) &&
                   (R2 < 1024.0) && (Bt < 0.125) && (Bt > 0.0);
  double c = fmin(0.25, fmax(3.0 * Bt, 1.0 / 1024.0));
  const double K = 1.02 * Bt * Bt * (1.0 + 1.0 / c);
  c += 16.0 * u;  // covers the rounding of the final evaluations themselves
  s.band.K = __double2float_ru(K);
  s.band.ratio = __double2float_ru((1.0 + c) / (1.0 - c));
  s.band.two_K = __double2float_ru(2.0 * K) * 1.000001f;
  s.band.pad = 0.f;
  s.s_scale = fast ? sqrt(1.0 - c) / thr : 1.0;
  s.fast = fast;
}

// ------------------------------------------------------------------------------------------
// prep_points: grid (ceil(max_pp / 256), B) — float32 point pairs for the scorer
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_points(const PairDesc* __restrict__ desc,
                                                   const PairState* __restrict__ state,
                                                   PointPair32* __restrict__ pp) {
  const PairDesc d = desc[blockIdx.y];
  if (!state[blockIdx.y].fast) return;
  const double sc = state[blockIdx.y].s_scale;
  const int npp = (d.n + 1) >> 1;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= npp) return;
  const int p = 2 * g, q = min(2 * g + 1, d.n - 1);
  const double2 a1 = reinterpret_cast<const double2*>(d.x1)[p];
  const double2 b1 = reinterpret_cast<const double2*>(d.x1)[q];
  const double2 a2 = reinterpret_cast<const double2*>(d.x2)[p];
  const double2 b2 = reinterpret_cast<const double2*>(d.x2)[q];
  PointPair32 o;
  o.x1 = make_float2((float)a1.x, (float)b1.x);
  o.y1 = make_float2((float)a1.y, (float)b1.y);
  o.x2 = make_float2((float)a2.x, (float)b2.x);
  o.y2 = make_float2((float)a2.y, (float)b2.y);
  o.x2s = make_float2((float)(a2.x * sc), (float)(b2.x * sc));
  o.y2s = make_float2((float)(a2.y * sc), (float)(b2.y * sc));
  pp[d.pp_off + g] = o;
}

// ------------------------------------------------------------------------------------------
// solve_sets: one thread per minimal set, grid (ceil(H/32), B), block 32
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void make_hyp32(const double* E, double sc, Hyp32& h) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += E[i] * E[i];
  const double f = 1.0 / sqrt(s);
  h.e00 = (float)(E[0] * f); h.e01 = (float)(E[1] * f); h.e02 = (float)(E[2] * f);
  h.e10 = (float)(E[3] * f); h.e11 = (float)(E[4] * f); h.e12 = (float)(E[5] * f);
  h.g0 = (float)(E[6] * f * sc); h.g1 = (float)(E[7] * f * sc);
  h.g2 = (float)(E[8] * f * sc);
  h.e20 = (float)(E[6] * f); h.e21 = (float)(E[7] * f);
  h.pad = 0.f;
}

// re-gathers the five point pairs of minimal set h (solve5_coop.cuh: Reload)
struct GatherSet {
  static constexpr bool kEnabled = true;
  const PairDesc& d;
  int h;
  __device__ __forceinline__ void operator()(double (&q)[5][2], double (&qp)[5][2]) const {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int idx = d.sets[5 * (size_t)h + i];
      idx = max(0, min(idx, d.n - 1));
      const double2 a = reinterpret_cast<const double2*>(d.x1)[idx];
      const double2 c = reinterpret_cast<const double2*>(d.x2)[idx];
      q[i][0] = a.x; q[i][1] = a.y;
      qp[i][0] = c.x; qp[i][1] = c.y;
    }
  }
};

__global__ void __launch_bounds__(32) solve_sets(const PairDesc* __restrict__ desc,
                                                 PairState* __restrict__ state, int H,
                                                 int with_cheirality,
                                                 double* __restrict__ E_list,
                                                 double* __restrict__ P_list,
                                                 int32_t* __restrict__ n_valid,
                                                 int32_t* __restrict__ n_roots,
                                                 Hyp32* __restrict__ hyp,
                                                 int32_t* __restrict__ hyp_id,
                                                 uint32_t* __restrict__ notin,
                                                 uint32_t* __restrict__ out) {
  __shared__ double sB[kCoopBasisDoubles][kCoopStride];
  __shared__ double sR[kCoopRowsDoubles][kCoopStride];
  __shared__ int sOk[32];
  const int h_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = h_raw < H;        // every lane takes part in the cooperative phase
  const int h = valid ? h_raw : H - 1;
  const int b = blockIdx.y;
  const PairDesc d = desc[b];
  const GatherSet gather{d, h};
  double q[5][2], qp[5][2];
  gather(q, qp);
  const size_t s = (size_t)b * H + h;
  double* E = E_list + s * 90;
  double* P = P_list ? P_list + s * 120 : nullptr;
  int nr = 0;
  // float32 hypothesis records are built while E is in registers; one slot per solution
  // (slot order is arbitrary: every record carries its id = set*16 + index)
  auto emit = [&](int j, const double (&Ereg)[9]) {
    if (!hyp) return;
    const int slot = atomicAdd(&state[b].M, 1);
    const size_t o = (size_t)b * H * 10 + slot;
    Hyp32 r;
    make_hyp32(Ereg, state[b].s_scale, r);
    hyp[o] = r;
    hyp_id[o] = h * 16 + j;
    notin[o] = 0u;
    out[o] = 0u;
  };
  const int nv = solve_minimal_set_coop(valid, q, qp, with_cheirality != 0, E, P, &nr, sB, sR, nullptr, sOk, emit, gather);
  if (!valid) return;
  n_valid[s] = nv;
  if (n_roots) n_roots[s] = nr;
}

// ------------------------------------------------------------------------------------------
// split solver (solve5_split.cuh): solve_front -> solve_roots -> solve_poses
// ------------------------------------------------------------------------------------------
// SPW = sets per warp (32, 16 or 8); shared memory is sized for SPW lanes (stride SPW + 1).
#ifndef TV5_FRONT_ALIAS
#define TV5_FRONT_ALIAS 1
#endif
#ifndef TV5_FRONT_MINB
#define TV5_FRONT_MINB 12   // <= 168 registers: three warps per SM sub-partition (16,384 registers each)
#endif
template <int SPW>
__global__ void __launch_bounds__(32, TV5_FRONT_MINB) solve_front(const PairDesc* __restrict__ desc, int H,
                                                  double* __restrict__ rec) {
  constexpr int S = SPW + 1;
#if TV5_FRONT_ALIAS
  __shared__ double sAll[kCoopRowsDoubles][S];   // rows 0..35: a set's basis, later its reduced rows (solve_front_warp)
  double (*sB)[S] = sAll;
  double (*sR)[S] = sAll;
#else
  __shared__ double sB[kCoopBasisDoubles][S];
  __shared__ double sR[kCoopRowsDoubles][S];
#endif
  __shared__ int sOk[32];
  const int first = blockIdx.x * SPW;
  const int h_raw = first + threadIdx.x;
  const bool valid = (int)threadIdx.x < SPW && h_raw < H;
  const int h = h_raw < H ? h_raw : H - 1;
  const int b = blockIdx.y;
  const PairDesc d = desc[b];
  const GatherSet gather{d, h};
  solve_front_warp<S>(valid, gather, rec + ((size_t)b * H + first) * kRecDoubles, min(SPW, H - first), sB, sR, sOk,
                      SPW);
}

static void launch_solve_front(int spw, int H, int nb, cudaStream_t st, const PairDesc* desc, double* rec) {
  const dim3 grid((H + spw - 1) / spw, nb);
  if (spw == 8) solve_front<8><<<grid, 32, 0, st>>>(desc, H, rec);
  else if (spw == 16) solve_front<16><<<grid, 32, 0, st>>>(desc, H, rec);
  else solve_front<32><<<grid, 32, 0, st>>>(desc, H, rec);
}

#ifndef TV5_ROOTS_MINB
#define TV5_ROOTS_MINB 6   // 168 registers, no spills: three warps per SM sub-partition
#endif
__global__ void __launch_bounds__(64, TV5_ROOTS_MINB) solve_roots(PairState* __restrict__ state, int H,
                                                  double* __restrict__ rec, RootEntry* __restrict__ entries,
                                                  int32_t* __restrict__ n_roots, int32_t* __restrict__ valid_mask) {
  const int h = blockIdx.x * 64 + threadIdx.x;
  if (h >= H) return;
  const int b = blockIdx.y;
  const size_t s = (size_t)b * H + h;
  // the set's entries go straight to their slots (reserved by one atomicAdd once the count is known)
  const int n = solve_roots_set(rec + s * kRecDoubles, h, [&](int cnt) {
    return entries + (size_t)b * H * 10 + atomicAdd(&state[b].n_entries, cnt);
  });
  if (n_roots) n_roots[s] = n;
  valid_mask[s] = 0;
}

#ifndef TV5_POSES_MINB
#define TV5_POSES_MINB 4
#endif
__global__ void __launch_bounds__(128, TV5_POSES_MINB) solve_poses(const PairDesc* __restrict__ desc, PairState* __restrict__ state,
                                                   int H, int with_cheirality, const double* __restrict__ rec,
                                                   const RootEntry* __restrict__ entries,
                                                   double* __restrict__ E_list, double* __restrict__ P_list,
                                                   int32_t* __restrict__ valid_mask, Hyp32* __restrict__ hyp,
                                                   int32_t* __restrict__ hyp_id, uint32_t* __restrict__ notin,
                                                   uint32_t* __restrict__ out) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= state[b].n_entries) return;
  const RootEntry en = entries[(size_t)b * H * 10 + e];
  const size_t s = (size_t)b * H + en.set;
  const int r = en.r_exact & 255;
  const PairDesc d = desc[b];
  const GatherSet gather{d, en.set};
  double E[9], P[12];
  if (!solve_pose_root(rec + s * kRecDoubles, en, with_cheirality != 0, gather, E, P)) return;
#pragma unroll
  for (int c = 0; c < 9; ++c) E_list[s * 90 + r * 9 + c] = E[c];
  if (P_list && with_cheirality) {
#pragma unroll
    for (int c = 0; c < 12; ++c) P_list[s * 120 + r * 12 + c] = P[c];
  }
  atomicOr(&valid_mask[s], 1 << r);
  if (hyp) {
    const int slot = atomicAdd(&state[b].M, 1);
    const size_t o = (size_t)b * H * 10 + slot;
    Hyp32 rcd;
    make_hyp32(E, state[b].s_scale, rcd);
    hyp[o] = rcd;
    hyp_id[o] = en.set * 16 + r;
    notin[o] = 0u;
    out[o] = 0u;
  }
}

// tv5_solve5: solutions stored at their root index -> lists compacted to the valid ones
__global__ void compact_solutions(int n_sets, double* __restrict__ E_list, double* __restrict__ P_list,
                                  int32_t* __restrict__ mask_to_count) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sets) return;
  const unsigned m = (unsigned)mask_to_count[s];
  int j = 0;
  for (int r = 0; r < 10; ++r) {
    if (!((m >> r) & 1u)) continue;
    if (j != r) {
      for (int c = 0; c < 9; ++c) { E_list[(size_t)s * 90 + j * 9 + c] = E_list[(size_t)s * 90 + r * 9 + c]; E_list[(size_t)s * 90 + r * 9 + c] = 0.0; }
      if (P_list)
        for (int c = 0; c < 12; ++c) { P_list[(size_t)s * 120 + j * 12 + c] = P_list[(size_t)s * 120 + r * 12 + c]; P_list[(size_t)s * 120 + r * 12 + c] = 0.0; }
    }
    ++j;
  }
  mask_to_count[s] = j;
}

// float32 hypothesis records for an arbitrary E list (tv5_score_bounds)
__global__ void hyps_from_list(const double* __restrict__ E_list, int M, PairState* state,
                               Hyp32* __restrict__ hyp, int32_t* hyp_id, uint32_t* notin,
                               uint32_t* out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0) state[0].M = M;
  if (m >= M) return;
  Hyp32 r;
  make_hyp32(E_list + 9 * (size_t)m, state[0].s_scale, r);
  hyp[m] = r;
  hyp_id[m] = m;
  notin[m] = 0u;
  out[m] = 0u;
}

// ------------------------------------------------------------------------------------------
// plan_tiles: one CTA of 1024 threads; tile table = exclusive prefix sum of tiles per pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) plan_tiles(const PairDesc* __restrict__ desc,
                                                   PairState* __restrict__ state,
                                                   Control* __restrict__ ctl, int B, int pp_per_tile,
                                                   int hyp_chunk) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int tiles = 0;
    if (b < B) {
      PairState& s = state[b];
      const int npp_all = (desc[b].n_full + 1) >> 1;
      if (!s.staged) { s.pp_lo = 0; s.pp_hi = npp_all; }      // no staging: the whole pair
      const int npp = max(0, min(s.pp_hi, npp_all) - s.pp_lo);
      s.n_hc = (s.M + hyp_chunk - 1) / hyp_chunk;
      s.n_pc = (npp + pp_per_tile - 1) / pp_per_tile;
      tiles = s.fast ? s.n_hc * s.n_pc : 0;
    }
    // inclusive scan inside the warp, then across warps
    int v = tiles;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) >= o) v += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      s_warp[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
    const int base = s_base;
    if (b < B) state[b].tile_start = base + warp_off + v - tiles;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ctl->n_tiles = s_base;
    ctl->tile_counter = 0;
  }
}

// ------------------------------------------------------------------------------------------
// score_bounds: persistent CTAs pulling (pair, hypothesis chunk, point chunk) tiles.
// ------------------------------------------------------------------------------------------
// HPT hypotheses per thread: kHypPerThread for the bulk of the work; 1 (256-slot chunks, more CTAs
// per SM) for the late early-exit stages, where a pair has only a few hundred hypotheses left.
// PARTIAL: tiles of a pair's last, partly filled hypothesis chunk evaluate only the register blocks
// that hold hypotheses (see below); chosen by the host when a pair has few chunks.
template <bool TWO_SIDED, int HPT = kHypPerThread, bool PARTIAL = false>
__global__ void __launch_bounds__(kScoreThreads, (HPT >= kHypPerThread ? TV5_SCORE_MINB : 4))
score_bounds(const PairDesc* __restrict__ desc, const PairState* __restrict__ state,
             Control* __restrict__ ctl, int B, int H, int pp_per_tile,
             const PointPair32* __restrict__ pp, const Hyp32* __restrict__ hyp,
             uint32_t* __restrict__ notin, uint32_t* __restrict__ out) {
  __shared__ __align__(128) PointPair32 tile[kMaxTilePairs];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_tile;
  const int tid = threadIdx.x;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  const int n_tiles = ctl->n_tiles;
  for (;;) {
    if (tid == 0) s_tile = atomicAdd(&ctl->tile_counter, 1);
    __syncthreads();
    const int t = s_tile;
    if (t >= n_tiles) break;
    // locate the image pair (tile_start is ascending; slow pairs contribute no tiles)
    int lo = 0, hi = B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (state[mid].tile_start <= t) lo = mid; else hi = mid - 1;
    }
    int b = lo;
    while (!state[b].fast || state[b].n_hc * state[b].n_pc == 0) --b;  // skip empty entries
    const PairState& s = state[b];
    const PairDesc& d = desc[b];
    const int lt = t - s.tile_start;
    const int hc = lt / s.n_pc, pc = lt - hc * s.n_pc;
    const int npp_all = (d.n_full + 1) >> 1;
    const int pp0 = s.pp_lo + pc * pp_per_tile;
    const int npp = min(pp_per_tile, min(s.pp_hi, npp_all) - pp0);
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)npp * (uint32_t)sizeof(PointPair32);
      mbar_expect_tx(&bar, bytes);
      bulk_load(tile, pp + d.pp_off + pp0, bytes, &bar);
    }
    // hypotheses of this thread (registers), loaded while the bulk copy is in flight
    const size_t hbase = (size_t)b * H * 10 + (size_t)hc * (kScoreThreads * HPT);
    HypRegs hr[HPT];
    bool live[HPT];
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
      const int m = hc * (kScoreThreads * HPT) + k * kScoreThreads + tid;
      live[k] = m < s.M;
      Hyp32 raw;
      if (live[k]) {
        const float4* src = reinterpret_cast<const float4*>(hyp + hbase + k * kScoreThreads + tid);
        float4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
        raw.e00 = v0.x; raw.e01 = v0.y; raw.e02 = v0.z; raw.e10 = v0.w;
        raw.e11 = v1.x; raw.e12 = v1.y; raw.g0 = v1.z; raw.g1 = v1.w;
        raw.g2 = v2.x; raw.e20 = v2.y; raw.e21 = v2.z; raw.pad = 0.f;
      } else {
        raw = Hyp32{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      }
      load_hyp(hr[k], raw);
    }
    const float2 bK = dup(s.band.K), nratio = dup(-s.band.ratio), ntwoK = dup(-s.band.two_K);
    uint32_t a[HPT], o[HPT];
#pragma unroll
    for (int k = 0; k < HPT; ++k) { a[k] = 0u; o[k] = 0u; }

    mbar_wait(&bar, phase);
    phase ^= 1u;
    // the very last point pair of an odd-sized image pair holds a duplicated point in lane .y
    const bool odd_tail = (pp0 + npp == npp_all) && (d.n_full & 1);
    const int nfull = odd_tail ? npp - 1 : npp;
    // A pair's last hypothesis chunk is usually not full (M mod 1024).  PARTIAL: its tiles evaluate only
    // as many hypotheses per thread as the chunk holds (CTA-uniform K in 1..HPT) instead of carrying
    // dead register blocks through the loop — 12 % of the evaluations of a 5,500-hypothesis shard
    // (configs[3] on 8 GPUs: 1.49 -> 1.39 ms).  For an 11,000-hypothesis pair it is 2 %, less than the
    // extra registers cost the main loop (15.36 -> 15.42 ms per 256 pairs), hence a separate instantiation.
    const int live_h = min(kScoreThreads * HPT, s.M - hc * (kScoreThreads * HPT));
    const int k_here = (live_h + kScoreThreads - 1) / kScoreThreads;
    auto run = [&](auto kc) {
      constexpr int K = decltype(kc)::value;
#pragma unroll kScoreUnroll
      for (int i = 0; i < nfull; ++i) {
        const PointPair32 p = tile[i];
#pragma unroll
        for (int k = 0; k < K; ++k) eval_pair<TWO_SIDED>(hr[k], p, bK, nratio, ntwoK, a[k], o[k]);
      }
      if (odd_tail) {
        const PointPair32 p = tile[nfull];
#pragma unroll
        for (int k = 0; k < K; ++k)
          eval_pair<TWO_SIDED>(hr[k], p, bK, nratio, ntwoK, a[k], o[k], false);
      }
    };
    if (!PARTIAL || HPT == 1 || k_here == HPT) run(std::integral_constant<int, HPT>());
    else if (k_here == 1) run(std::integral_constant<int, 1>());
    else if (k_here == 2) run(std::integral_constant<int, (HPT >= 2 ? 2 : 1)>());
    else run(std::integral_constant<int, (HPT >= 3 ? 3 : 1)>());
#pragma unroll
    for (int k = 0; k < HPT; ++k)
      if (live[k]) {
        const size_t slot = hbase + k * kScoreThreads + tid;
        if (TWO_SIDED && a[k]) atomicAdd(&notin[slot], a[k]);
        if (o[k]) atomicAdd(&out[slot], o[k]);
      }
    __syncthreads();  // everyone is done with `tile` and `s_tile` before the next round
  }
}

// ------------------------------------------------------------------------------------------
// Early exit (opt-in): the points are scored in stages; after a stage every hypothesis whose
// upper bound on the FULL count — points not yet seen counted as inliers — is below the exact
// count L of an actual hypothesis can no longer win and is dropped before the next stage.
//   set_stage      point-pair range of the next stage (fractions of each pair's points)
//   stage_leader   hypothesis with the fewest sure outliers so far -> cand[0] (exact_counts -> L)
//   prune_compact  survivors (out <= n - L) copied to the ping-pong arrays, M := survivor count
// Exactness: `out` counts evaluations that are outliers under every rounding (guard band), so
// n - out is a rigorous upper bound; a dropped hypothesis has count <= n - out < L <= winner.
// ------------------------------------------------------------------------------------------
__global__ void set_stage(const PairDesc* __restrict__ desc, PairState* __restrict__ state, int B,
                          float f_lo, float f_hi, int first, int pp_per_tile) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int npp = (desc[b].n_full + 1) >> 1;
  // stage boundaries on whole scoring tiles (64 point pairs for small inputs); the last stage
  // ends at the end
  const int q = npp >= 4 * pp_per_tile ? pp_per_tile : 64;
  const int lo = f_lo <= 0.f ? 0 : min(npp, (((int)(f_lo * npp) + q / 2) / q) * q);
  const int hi = f_hi >= 1.f ? npp : min(npp, (((int)(f_hi * npp) + q / 2) / q) * q);
  state[b].pp_lo = lo;
  state[b].pp_hi = max(hi, lo);
  state[b].staged = 1;
  if (first) state[b].M_total = state[b].M;
}

__global__ void __launch_bounds__(1024) stage_leader(const PairDesc* __restrict__ desc,
                                                    PairState* __restrict__ state, int H,
                                                    const uint32_t* __restrict__ out,
                                                    const int32_t* __restrict__ hyp_id,
                                                    int32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt) {
  const int b = blockIdx.x;
  PairState& s = state[b];
  const int M = s.M;
  const size_t base = (size_t)b * H * 10;
  if (threadIdx.x == 0) { cand_cnt[base] = 0; s.exact_from = 0; }
  if (!s.fast || M == 0) {  // float64 pairs are not staged: nothing to score, nothing will be pruned
    if (threadIdx.x == 0) s.n_cand = 0;
    return;
  }
  __shared__ unsigned long long s_key[32];
  unsigned long long key = 0ull;  // (fewest sure outliers, smallest id) as a maximum
#pragma unroll 4
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const unsigned long long k = ((unsigned long long)(0xFFFFFFFFu - out[base + m]) << 32) |
                                 (0xFFFFFFFFu - (uint32_t)hyp_id[base + m]);
    key = k > key ? k : key;
  }
  key = warp_max(key);
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
  __syncthreads();
  unsigned long long best = 0ull;
for (int w = 0; w < (int)(blockDim.x >> 5); ++w) best = s_key[w] > best ? s_key[w] : best;
  const int best_id = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
  if (threadIdx.x == 0) s.n_cand = 1;
  for (int m = threadIdx.x; m < M; m += blockDim.x)
    if (hyp_id[base + m] == best_id) cand[base] = m;
}

__global__ void __launch_bounds__(256) prune_compact(const PairDesc* __restrict__ desc,
                                                     PairState* __restrict__ state, int H,
                                                     const Hyp32* __restrict__ hyp, const int32_t* __restrict__ hyp_id,
                                                     const uint32_t* __restrict__ out,
                                                     const int32_t* __restrict__ cand_cnt,
                                                     Hyp32* __restrict__ hyp_dst, int32_t* __restrict__ id_dst,
                                                     uint32_t* __restrict__ out_dst) {
  const int b = blockIdx.x;
  PairState& s = state[b];
  const size_t base = (size_t)b * H * 10;
  const int M = s.M;
  const int n = desc[b].n_full;
  const int L = cand_cnt[base];                       // exact full count of the stage leader
  const uint32_t max_out = (uint32_t)max(n - L, 0);  // survive iff n - out >= L
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int m0 = 0; m0 < M; m0 += blockDim.x) {
    const int m = m0 + threadIdx.x;
    const bool keep = m < M && (!s.fast || out[base + m] <= max_out);
    // block-wide compaction: ballot per warp, one shared-memory atomic per warp
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31;
    int wbase = 0;
    if (lane == 0 && bal) wbase = atomicAdd(&s_count, __popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (keep) {
      const int dst = wbase + __popc(bal & ((1u << lane) - 1u));
      hyp_dst[base + dst] = hyp[base + m];
      id_dst[base + dst] = hyp_id[base + m];
      out_dst[base + dst] = out[base + m];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s.M = s_count;
}

// ------------------------------------------------------------------------------------------
// Candidate selection from the upper bounds hi[m] = n - out[m]  (exact count <= hi), one CTA per
// image pair, two rounds:
//   pick_top   the hypothesis with the largest hi (first by id)       -> exact count L
//   pick_rest  every other hypothesis with hi >= max(L, 1)            -> exact counts
// The true winner w has exact(w) >= L and hi(w) >= exact(w), so it is in the second set; ties
// are all included, so the first-maximum rule can be applied exactly by `finalize`.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pick_top(const PairDesc* __restrict__ desc,
                                                PairState* __restrict__ state, int H,
                                                const uint32_t* __restrict__ out,
                                                const int32_t* __restrict__ hyp_id,
                                                int32_t* __restrict__ cand,
                                                int32_t* __restrict__ cand_cnt) {
  const int b = blockIdx.x;
  PairState& s = state[b];
  const int M = s.M;
  const size_t base = (size_t)b * H * 10;
  if (!s.fast) {  // float64 path: every hypothesis is a candidate
    for (int m = threadIdx.x; m < M; m += blockDim.x) { cand[base + m] = m; cand_cnt[base + m] = 0; }
    if (threadIdx.x == 0) { s.n_cand = M; s.exact_from = 0; }
    return;
  }
  __shared__ unsigned long long s_key[32];
  const int n = desc[b].n_full;
  unsigned long long key = 0ull;  // (hi << 32 | ~id) << 0, slot recovered by a second pass
#pragma unroll 4
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const uint32_t hi = (uint32_t)(n - (int)out[base + m]);
    const unsigned long long k = ((unsigned long long)hi << 32) | (0xFFFFFFFFu - (uint32_t)hyp_id[base + m]);
    key = k > key ? k : key;
  }
  key = warp_max(key);
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
  __syncthreads();
  unsigned long long best = 0ull;
for (int w = 0; w < (int)(blockDim.x >> 5); ++w) best = s_key[w] > best ? s_key[w] : best;
  const int best_id = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
  if (threadIdx.x == 0) { s.n_cand = M > 0 ? 1 : 0; s.exact_from = 0; }
  for (int m = threadIdx.x; m < M; m += blockDim.x)
    if (hyp_id[base + m] == best_id) { cand[base] = m; cand_cnt[base] = 0; }
}

__global__ void __launch_bounds__(1024) pick_rest(const PairDesc* __restrict__ desc,
                                                 PairState* __restrict__ state, int H,
                                                 const uint32_t* __restrict__ out,
                                                 int32_t* __restrict__ cand,
                                                 int32_t* __restrict__ cand_cnt) {
  const int b = blockIdx.x;
  PairState& s = state[b];
  const size_t base = (size_t)b * H * 10;
  if (!s.fast || s.M == 0) {
    if (threadIdx.x == 0) s.exact_from = s.n_cand;  // nothing left to re-score
    return;
  }
  const int M = s.M;
  const int n = desc[b].n_full;
  const int top = cand[base];
  const int L = max(cand_cnt[base], 1);  // a hypothesis that cannot have a single inlier never wins
  __syncthreads();
  if (threadIdx.x == 0) s.exact_from = 1;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    if (m == top) continue;
    const int hi = n - (int)out[base + m];
    if (hi >= L) {
      const int i = atomicAdd(&s.n_cand, 1);
      cand[base + i] = m;
      cand_cnt[base + i] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// exact_counts: float64 re-score of the candidates.  grid (X, B); work item = (candidate,
// chunk of kExactChunk points), strided over X CTAs.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exact_counts(const PairDesc* __restrict__ desc,
                                                    const PairState* __restrict__ state, int H,
                                                    double thr, const double* __restrict__ E_list,
                                                    const int32_t* __restrict__ hyp_id,
                                                    const int32_t* __restrict__ cand,
                                                    int32_t* __restrict__ cand_cnt) {
  const int b = blockIdx.y;
  const PairDesc d = desc[b];
  const int first = state[b].exact_from;
  const int n_cand = state[b].n_cand - first;
  const int n = d.n_full;
  const int n_chunks = (n + kExactChunk - 1) / kExactChunk;
  const size_t base = (size_t)b * H * 10;
  __shared__ int s_part[8];
  for (int item = blockIdx.x; item < n_cand * n_chunks; item += gridDim.x) {
    const int ci = first + item / n_chunks, ch = item % n_chunks;
    const int id = hyp_id[base + cand[base + ci]];
    const double* Eg = E_list + ((size_t)b * H + (id >> 4)) * 90 + (id & 15) * 9;
    double E[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) E[i] = Eg[i];
    int c = 0;
    const int k1 = min(n, (ch + 1) * kExactChunk);
    for (int k = ch * kExactChunk + threadIdx.x; k < k1; k += blockDim.x) {
      const double2 p1 = reinterpret_cast<const double2*>(d.x1)[k];
      const double2 p2 = reinterpret_cast<const double2*>(d.x2)[k];
      c += sampson_inlier_exact(E, p1.x, p1.y, p2.x, p2.y, thr) ? 1 : 0;
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      if (t) atomicAdd(&cand_cnt[base + ci], t);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// finalize: first-max selection + outputs.  One CTA per image pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize(const PairDesc* __restrict__ desc,
                                                const PairState* __restrict__ state, int H,
                                                double thr, const double* __restrict__ E_list,
                                                const double* __restrict__ P_list,
                                                const int32_t* __restrict__ hyp_id,
                                                const int32_t* __restrict__ cand,
                                                const int32_t* __restrict__ cand_cnt,
                                                const int32_t* __restrict__ valid_mask) {
  const int b = blockIdx.x;
  const PairDesc d = desc[b];
  const PairState& s = state[b];
  const size_t base = (size_t)b * H * 10;
  __shared__ unsigned long long s_key[8];
  __shared__ unsigned long long s_win;
  unsigned long long key = 0ull;
  for (int i = threadIdx.x; i < s.n_cand; i += blockDim.x) {
    const uint32_t cnt = (uint32_t)cand_cnt[base + i];
    const uint32_t id = (uint32_t)hyp_id[base + cand[base + i]];
    const unsigned long long k = ((unsigned long long)cnt << 32) | (0xFFFFFFFFu - id);
    key = k > key ? k : key;
  }
  key = warp_max(key);
  if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long w = 0ull;
    for (int i = 0; i < 8; ++i) w = s_key[i] > w ? s_key[i] : w;
    s_win = w;
  }
  __syncthreads();
  const unsigned long long win = s_win;
  const int count = (int)(win >> 32);
  const int id = (int)(0xFFFFFFFFu - (uint32_t)(win & 0xFFFFFFFFull));
  const bool have = count > 0;
  const int sid = have ? id : 0;
  const double* Eg = E_list + ((size_t)b * H + (sid >> 4)) * 90 + (sid & 15) * 9;
  if (threadIdx.x < 9) d.E_out[threadIdx.x] = have ? Eg[threadIdx.x] : 0.0;
  if (d.P_out && threadIdx.x < 12) {
    const double* Pg = P_list ? P_list + ((size_t)b * H + (sid >> 4)) * 120 + (sid & 15) * 12 : nullptr;
    d.P_out[threadIdx.x] = (have && Pg) ? Pg[threadIdx.x] : 0.0;
  }
  if (threadIdx.x == 0) {
    tv5_result r;
    r.count = count;
    r.best_set = have ? (id >> 4) : -1;
    // split solver: ids carry the root index; the compacted index counts the valid roots below it
    int root = id & 15;
    if (have && valid_mask) root = __popc((unsigned)valid_mask[(size_t)b * H + (id >> 4)] & ((1u << root) - 1u));
    r.best_root = have ? root : -1;
    r.n_hypotheses = s.M_total > 0 ? s.M_total : s.M;
    r.n_candidates = s.n_cand;
    r.fast_path = s.fast;
    r.reserved[0] = r.reserved[1] = 0;
    *d.result = r;
  }
  if (d.mask_out) {
    double E[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) E[i] = have ? Eg[i] : 0.0;
    for (int k = threadIdx.x; k < d.n_full; k += blockDim.x) {
      const double2 p1 = reinterpret_cast<const double2*>(d.x1)[k];
      const double2 p2 = reinterpret_cast<const double2*>(d.x2)[k];
      d.mask_out[k] = (have && sampson_inlier_exact(E, p1.x, p1.y, p2.x, p2.y, thr)) ? 1 : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// general two-stage selection (n_pre != n_full): float64 only.
//   stage A  counts of every hypothesis on the first n_pre points     (exact_counts)
//   per set  best root = first maximum                                (set_winners)
//   stage B  counts of the per-set winners on the first n_full points (exact_counts)
//   finalize first maximum over sets
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) set_winners(PairDesc* __restrict__ desc,
                                                   PairState* __restrict__ state, int H,
                                                   const int32_t* __restrict__ hyp_id,
                                                   int32_t* __restrict__ cand,
                                                   int32_t* __restrict__ cand_cnt,
                                                   int32_t* __restrict__ scratch) {
  // Stage A left cand[i] = i (all hypotheses) and cand_cnt[i] = count on n_pre points.  Per
  // minimal set the reference keeps the FIRST maximum over its roots (kernel_functions.cu:
  // 186-202, strict >).  Slots are handed out by atomicAdd in arbitrary order, so the roots of a
  // set are found by set id, never by slot adjacency: every hypothesis folds the key
  // ((count+1) << 4 | 15 - root) into its set's cell with atomicMax; the hypothesis whose key
  // equals the cell is the set's winner (unique: the root index is part of the key).
  const int b = blockIdx.x;
  PairState& s = state[b];
  const size_t base = (size_t)b * H * 10;
  const size_t sbase = (size_t)b * H * 20;  // scratch: [H] per-set keys, then [H*10] winner slots
  int32_t* set_key = scratch + sbase;
  int32_t* list = scratch + sbase + (size_t)H * 10;
  const int M = s.M;
  for (int h = threadIdx.x; h < H; h += blockDim.x) set_key[h] = 0;
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int id = hyp_id[base + m];
    atomicMax(&set_key[id >> 4], ((cand_cnt[base + m] + 1) << 4) | (15 - (id & 15)));
  }
  __syncthreads();
  if (threadIdx.x == 0) { s.n_cand = 0; s.exact_from = 0; }
  __syncthreads();
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int id = hyp_id[base + m];
    if (set_key[id >> 4] == (((cand_cnt[base + m] + 1) << 4) | (15 - (id & 15)))) list[atomicAdd(&s.n_cand, 1)] = m;
  }
  __syncthreads();
  const int nc = s.n_cand;
  for (int i = threadIdx.x; i < nc; i += blockDim.x) {
    cand[base + i] = list[i];
    cand_cnt[base + i] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// stand-alone exact scorer for tv5_score: grid (ceil(n/1024), M), block 256
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_exact_list(const double* __restrict__ x1,
                                                        const double* __restrict__ x2, int n,
                                                        const double* __restrict__ E_list,
                                                        double thr, int32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ masks) {
  const int m = blockIdx.y;
  double E[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) E[i] = E_list[9 * (size_t)m + i];
  const int words = (n + 31) >> 5;
  int c = 0;
  const int k0 = blockIdx.x * 1024;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = k0 + r * 256 + threadIdx.x;
    bool in = false;
    if (k < n) {
      const double2 p1 = reinterpret_cast<const double2*>(x1)[k];
      const double2 p2 = reinterpret_cast<const double2*>(x2)[k];
      in = sampson_inlier_exact(E, p1.x, p1.y, p2.x, p2.y, thr);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0) {
      c += __popc(bal);
      if (masks && (k >> 5) < words) masks[(size_t)m * words + (k >> 5)] = bal;
    }
  }
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counts[m], c);
}

__global__ void bounds_to_lo_hi(const uint32_t* notin, const uint32_t* out, int M, int n,
                                int32_t* lo, int32_t* hi) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) { lo[m] = n - (int)notin[m]; hi[m] = n - (int)out[m]; }
}

// ------------------------------------------------------------------------------------------
// reference RNG index table
// ------------------------------------------------------------------------------------------
__global__ void ref_rng_kernel(int N, int iters, int32_t* __restrict__ sets) {
  const int gi = threadIdx.x + blockDim.x * blockIdx.x;
  curandState st;
  curand_init(1234ULL, gi, 0, &st);
  for (int it = 0; it < iters; ++it)
    for (int i = 0; i < 5; ++i) {
      float r = curand_uniform(&st);
      r *= ((N - 1) - 0 + 0.999999f);
      r += 0;
      int idx = (int)truncf(r);
      sets[((size_t)gi * iters + it) * 5 + i] = min(idx, N - 1);
    }
}

// ------------------------------------------------------------------------------------------
// hypothesis-sharded single pair: 192-byte winner record per GPU, first-maximum pick over G records
// ------------------------------------------------------------------------------------------
struct WinnerRecord {
  unsigned long long key;   // count << 32 | ~(global set * 16 + root); 0 = no hypothesis with an inlier
  double E[9];
  double P[12];
  int32_t n_hyp, n_cand, fast, pad;
};
static_assert(sizeof(WinnerRecord) == TV5_WINNER_RECORD_BYTES, "record layout");

__global__ void winner_record(const double* __restrict__ E, const double* __restrict__ P,
                              const tv5_result* __restrict__ res, int set_offset, WinnerRecord* __restrict__ out) {
  const int t = threadIdx.x;
  const tv5_result r = *res;
  const bool have = r.count > 0 && r.best_set >= 0;
  if (t < 9) out->E[t] = E[t];
  if (t < 12) out->P[t] = P ? P[t] : 0.0;
  if (t == 0) {
    const uint32_t id = (uint32_t)(r.best_set + set_offset) * 16u + (uint32_t)r.best_root;
    out->key = have ? (((unsigned long long)(uint32_t)r.count << 32) | (0xFFFFFFFFu - id)) : 0ull;
    out->n_hyp = r.n_hypotheses;
    out->n_cand = r.n_candidates;
    out->fast = r.fast_path;
    out->pad = 0;
  }
}

__global__ void winner_pick(const WinnerRecord* __restrict__ recs, int G, double* __restrict__ E_out,
                            double* __restrict__ P_out, tv5_result* __restrict__ res_out) {
  // one warp: lane g folds records g, g + 32, ...; ids are unique across ranks, so the maximum key is too
  const int lane = threadIdx.x;
  unsigned long long key = 0ull;
  int n_hyp = 0, n_cand = 0, fast = 1;
  for (int g = lane; g < G; g += 32) {
    key = recs[g].key > key ? recs[g].key : key;
    n_hyp += recs[g].n_hyp;
    n_cand += recs[g].n_cand;
    fast &= recs[g].fast;
  }
  const unsigned long long best = warp_max(key);
  n_hyp = warp_sum(n_hyp);
  n_cand = warp_sum(n_cand);
  fast = __all_sync(0xffffffffu, fast);
  int owner = -1;
  for (int g = lane; g < G; g += 32)
    if (best != 0ull && recs[g].key == best) owner = g;
  owner = warp_max(owner);
  if (lane < 9) E_out[lane] = owner >= 0 ? recs[owner].E[lane] : 0.0;
  if (P_out && lane < 12) P_out[lane] = owner >= 0 ? recs[owner].P[lane] : 0.0;
  if (lane == 0) {
    tv5_result r;
    const uint32_t id = 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull);
    r.count = (int32_t)(best >> 32);
    r.best_set = owner >= 0 ? (int32_t)(id >> 4) : -1;
    r.best_root = owner >= 0 ? (int32_t)(id & 15u) : -1;
    r.n_hypotheses = n_hyp;
    r.n_candidates = n_cand;
    r.fast_path = fast;
    r.reserved[0] = r.reserved[1] = 0;
    *res_out = r;
  }
}

// The curand_uniform stream of the reference does not depend on N (kernel_functions.cu:269-278 scales
// the draw afterwards), so the draws are generated once per context — draw-major, u[d*512 + tid] =
// d-th draw of reference thread tid, which makes the table for `iters` a prefix of every longer
// one — and each submission only applies the reference's float32 scaling for its own N on the
// stream: no allocation, no synchronisation, no curand_init per call.
__global__ void ref_rng_uniform_kernel(int draws, float* __restrict__ u) {
  const int gi = threadIdx.x + blockDim.x * blockIdx.x;
  curandState st;
  curand_init(1234ULL, gi, 0, &st);
  for (int d = 0; d < draws; ++d) u[(size_t)d * TV5_REF_THREADS + gi] = curand_uniform(&st);
}

// grid (ceil(H*5/256), B): sets[b][h][k] for h = tid*iters + it  <-  draw it*5+k of thread tid
__global__ void __launch_bounds__(256) rng_scale_sets(const PairDesc* __restrict__ desc, int H, int iters,
                                                      const float* __restrict__ u, int32_t* __restrict__ sets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * 5) return;
  const int N = desc[blockIdx.y].n;
  const int h = i / 5, k = i - 5 * h, tid = h / iters, it = h - tid * iters;
  float r = u[(size_t)(it * 5 + k) * TV5_REF_THREADS + tid];
  r *= ((N - 1) - 0 + 0.999999f);
  r += 0;
  sets[(size_t)blockIdx.y * H * 5 + i] = min((int)truncf(r), N - 1);
}

// ------------------------------------------------------------------------------------------
// FP32 peak microbenchmark: 8 independent dependent-chains per thread
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* sink, int iters, float a, float b) {
  if (MODE == 0) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) sink[0] = s;
  } else {
    float2 v[8];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2((float)(threadIdx.x + i), (float)i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) sink[0] = s;
  }
}

}  // namespace tv5

// ==========================================================================================
// host side
// ==========================================================================================
using namespace tv5;

// Every entry point that touches the context's workspace or settings holds the context's (recursive)
// mutex for the duration of the call: two host threads sharing one context are serialised instead of
// racing on the workspace (entry points call each other, hence recursive; a null context locks nothing).
struct CtxLock {
  std::recursive_mutex* m;
  explicit CtxLock(tv5_ctx* c) : m(c ? &c->mu : nullptr) { if (m) m->lock(); }
  ~CtxLock() { if (m) m->unlock(); }
  CtxLock(const CtxLock&) = delete;
  CtxLock& operator=(const CtxLock&) = delete;
};
#define TV5_LOCK(ctx) CtxLock lock__(ctx)

#define TV5_CUDA(ctx, call)                         \
  do {                                              \
    cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) {                       \
      (ctx)->last_cuda = (int)e__;                  \
      return TV5_ERR_CUDA;                          \
    }                                               \
  } while (0)

// ------------------------------------------------------------------------------------------
// Workspace allocation.  In guard mode (tv5_debug_guard; a testing aid standing in for
// compute-sanitizer's memcheck / initcheck, which are closed on the B200 pool) every buffer gets a
// 256-byte zone of 0xA5 on either side, checked by tv5_debug_check_guards, and its payload is filled
// with a caller-chosen poison byte — at allocation and again on tv5_debug_poison — so that a kernel
// that reads workspace it did not write produces poison-dependent results (the tests run every
// scenario under several poisons and require bit-identical outputs).
// ------------------------------------------------------------------------------------------
constexpr size_t kGuardBytes = 256;
static cudaError_t ws_malloc_bytes(tv5_ctx* ctx, void** p, size_t bytes) {
  if (!ctx->guard) return cudaMalloc(p, bytes);
  const size_t payload = (bytes + 255) & ~(size_t)255;     // the back zone starts on an aligned address
  char* base = nullptr;
  cudaError_t e = cudaMalloc((void**)&base, payload + 2 * kGuardBytes);
  if (e != cudaSuccess) return e;
  cudaMemset(base, 0xA5, payload + 2 * kGuardBytes);
  cudaMemset(base + kGuardBytes, ctx->poison, bytes);
  ctx->guards.push_back({base + kGuardBytes, base, bytes, payload});
  *p = base + kGuardBytes;
  return cudaSuccess;
}
template <typename T>
static cudaError_t ws_malloc(tv5_ctx* ctx, T** p, size_t bytes) { return ws_malloc_bytes(ctx, (void**)p, bytes); }
static void ws_free(tv5_ctx* ctx, void* p) {
  if (!p) return;
  for (size_t i = 0; i < ctx->guards.size(); ++i)
    if (ctx->guards[i].user == p) {
      cudaFree(ctx->guards[i].base);
      ctx->guards.erase(ctx->guards.begin() + i);
      return;
    }
  cudaFree(p);
}

template <typename T>
static int grow(tv5_ctx* ctx, T*& p, size_t& cap, size_t need) {
  if (need <= cap && p) return TV5_OK;
  if (p) ws_free(ctx, p);
  p = nullptr;
  size_t n = std::max(need, cap + cap / 2);
  if (ws_malloc(ctx, &p, n * sizeof(T)) != cudaSuccess) { p = nullptr; cap = 0; ctx->last_cuda = (int)cudaGetLastError(); return TV5_ERR_NOMEM; }
  cap = n;
  return TV5_OK;
}
template <typename T>
static int grow_same(tv5_ctx* ctx, T*& p, size_t old_cap, size_t new_cap) {
  if (new_cap == old_cap && p) return TV5_OK;
  if (p) ws_free(ctx, p);
  p = nullptr;
  if (ws_malloc(ctx, &p, new_cap * sizeof(T)) != cudaSuccess) { p = nullptr; ctx->last_cuda = (int)cudaGetLastError(); return TV5_ERR_NOMEM; }
  return TV5_OK;
}

static int ensure_workspace_impl(tv5_ctx* ctx, int B, size_t total_pp, size_t total_sets);
static int ensure_workspace(tv5_ctx* ctx, int B, size_t total_pp, size_t total_sets) {
  Workspace& w = ctx->ws;
  const void* before[4] = {w.desc, w.pp, w.E_list, w.hyp};
  const int rc = ensure_workspace_impl(ctx, B, total_pp, total_sets);
  const void* after[4] = {w.desc, w.pp, w.E_list, w.hyp};
  if (memcmp(before, after, sizeof(before)) != 0) {   // captured graphs hold the old addresses
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
  }
  return rc;
}

static int ensure_workspace_impl(tv5_ctx* ctx, int B, size_t total_pp, size_t total_sets) {
  Workspace& w = ctx->ws;
  int rc;
  if ((size_t)B > w.desc_cap || !w.desc) {
    const size_t oc = w.desc_cap;
    const size_t nc = std::max((size_t)B, oc * 2);
    w.desc_cap = 0;
    if ((rc = grow_same(ctx, w.desc, oc, nc))) return rc;
    if ((rc = grow_same(ctx, w.state, oc, nc))) return rc;
    w.desc_cap = nc;
  }
  if (!w.ctl && ws_malloc(ctx, &w.ctl, sizeof(Control) * kPipeChunks) != cudaSuccess) { w.ctl = nullptr; return TV5_ERR_NOMEM; }
  if ((rc = grow(ctx, w.pp, w.pp_cap, total_pp))) return rc;
  if (total_sets > w.sets_cap || !w.E_list) {
    const size_t oc = w.sets_cap;
    const size_t nc = std::max(total_sets, oc + oc / 2);
    // All per-set buffers change size together.  If any allocation fails, the capacity is reset to
    // zero so that the next call reallocates every one of them (grow_same frees what is still
    // held): a later, smaller submission can never run on a mix of old- and new-sized buffers.
    w.sets_cap = 0;
    w.hyp_cap = 0;
    RootEntry* en = (RootEntry*)w.entries;
    rc = grow_same(ctx, w.E_list, oc * 90, nc * 90);
    if (!rc) rc = grow_same(ctx, w.P_list, oc * 120, nc * 120);
    if (!rc) rc = grow_same(ctx, w.n_valid, oc, nc);
    if (!rc) rc = grow_same(ctx, w.n_roots, oc, nc);
    if (!rc) rc = grow_same(ctx, w.rec, oc * kRecDoubles, nc * kRecDoubles);
    if (!rc) { rc = grow_same(ctx, en, oc * 10, nc * 10); w.entries = en; }
    if (!rc) rc = grow_same(ctx, w.hyp, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.hyp_id, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.notin, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.out, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.hyp2, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.hyp_id2, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.out2, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.cand, oc * 30, nc * 30);  // + 2x set_winners scratch
    if (!rc) rc = grow_same(ctx, w.cand_cnt, oc * 10, nc * 10);
    if (!rc) rc = grow_same(ctx, w.rng_sets, oc * 5, nc * 5);
    if (rc) return rc;
    w.sets_cap = nc;
    w.hyp_cap = nc * 10;
  }
  return TV5_OK;
}

// Sets per warp of solve_front: fewer for small submissions (shorter critical path, more warps), 32 for
// throughput.  Measured round 2 (tools/batch_sweep.py, solver ms at 2 / 4 / 8 / 16 / 32 pairs of 4,096 sets):
// this rule 0.081 / 0.101 / 0.148 / 0.272 / 0.453; always 32: 0.094 / 0.105 / 0.148 / 0.272 / 0.453; a rule
// that minimises (waves x per-warp cost) picked 16 up to 32 pairs and lost: 0.183 / 0.316 / 0.487 at 8 / 16 / 32.
static int front_sets_per_warp(int64_t total_sets, int /*sm_count*/) {
  if (const char* e = getenv("TV5_FRONT_SPW")) return atoi(e) == 8 ? 8 : (atoi(e) == 16 ? 16 : 32);   // dev knob
  return total_sets <= 8192 ? 8 : (total_sets <= 16384 ? 16 : 32);
}

static void stage_mark(tv5_ctx* ctx, cudaStream_t st, int i) {
  if (ctx->profiling) cudaEventRecord(ctx->ev[i], st);
}

// profiling events of the pose pipeline: kProfPerChunk per chunk
//   0 prep start, 1 solve start, 2 solve end (front stream)
//   3 plan start, 4 score start, 5 score end, 6 finalize start, 7 finalize end (back stream)
static int ensure_prof_events(tv5_ctx* ctx, int n_chunks) {
  while ((int)ctx->prof_ev.size() < n_chunks * kProfPerChunk) {
    cudaEvent_t e;
    TV5_CUDA(ctx, cudaEventCreate(&e));
    ctx->prof_ev.push_back(e);
  }
  return TV5_OK;
}

static int ensure_pipe_streams(tv5_ctx* ctx) {
  if (ctx->front_stream) return TV5_OK;
  int lo = 0, hi = 0;
  TV5_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least priority (numerically largest)
  TV5_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->back_stream, cudaStreamNonBlocking, hi));
  TV5_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->front_stream, cudaStreamNonBlocking, lo));
  TV5_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_entry, cudaEventDisableTiming));
  TV5_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_done, cudaEventDisableTiming));
  for (auto& e : ctx->pipe_solved) TV5_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return TV5_OK;
}

static void profile_collect(tv5_ctx* ctx) {
  if (!ctx->profiling || !ctx->ev_pending) return;
  ctx->ev_pending = false;
  if (ctx->prof_chunks > 0) {  // pose pipeline
    static const int kFrom[TV5_N_STAGES] = {0, 1, 3, 4, 5, 6};
    static const int kTo[TV5_N_STAGES] = {1, 2, 4, 5, 6, 7};
    cudaEventSynchronize(ctx->prof_ev[(ctx->prof_chunks - 1) * kProfPerChunk + 7]);
    for (int i = 0; i < TV5_N_STAGES; ++i) {
      double sum = 0.0;
      bool ok = true;
      for (int c = 0; c < ctx->prof_chunks; ++c) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_ev[c * kProfPerChunk + kFrom[i]],
                                 ctx->prof_ev[c * kProfPerChunk + kTo[i]]) != cudaSuccess) { ok = false; break; }
        sum += ms;
      }
      if (ok) { ctx->stage_ms[i] += sum; ctx->stage_launches[i] += 1; }
    }
    ctx->prof_chunks = 0;
    return;
  }
  cudaEventSynchronize(ctx->ev[TV5_N_STAGES]);
  for (int i = 0; i < TV5_N_STAGES; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) {
      ctx->stage_ms[i] += ms;
      ctx->stage_launches[i] += 1;
    }
  }
}

extern "C" {

int tv5_version(void) { return TV5_VERSION; }

const char* tv5_strerror(int code) {
  switch (code) {
    case TV5_OK: return "ok";
    case TV5_ERR_INVALID: return "invalid argument";
    case TV5_ERR_CUDA: return "CUDA runtime error";
    case TV5_ERR_NOMEM: return "out of device memory";
    case TV5_ERR_NO_DEVICE: return "no usable CUDA device";
    default: return "unknown error";
  }
}

int tv5_create(int device, tv5_ctx** out) {
  if (!out) return TV5_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return TV5_ERR_NO_DEVICE;
  }
  tv5_ctx* ctx = new (std::nothrow) tv5_ctx();
  if (!ctx) return TV5_ERR_NOMEM;
  ctx->device = device;
  int rc = TV5_OK;
  do {
    if (cudaSetDevice(device) != cudaSuccess) { rc = TV5_ERR_CUDA; break; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { rc = TV5_ERR_CUDA; break; }
    if (prop.major != 10) { rc = TV5_ERR_NO_DEVICE; break; }  // sm_100a cubin only
    ctx->sm_count = prop.multiProcessorCount;
    // completion rows of the null-space system: the recurrence the reference runs per call
    // (essential_matrix_5pt.cu:639-649); plain IEEE double arithmetic, so host == device.
    double fill[4][9];
    double ran = 3.18730379;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 9; ++j) {
        ran *= 3.18730379;
        ran = 2.0 * (ran - floor(ran)) - 1.0;
        fill[i][j] = ran;
      }
    if (cudaMemcpyToSymbol(c_completion, fill, sizeof(fill)) != cudaSuccess) { rc = TV5_ERR_CUDA; break; }
    for (int i = 0; i <= TV5_N_STAGES; ++i)
      if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { rc = TV5_ERR_CUDA; break; }
  } while (0);
  if (rc != TV5_OK) {
    ctx->last_cuda = (int)cudaGetLastError();
    delete ctx;
    return rc;
  }
  *out = ctx;
  return TV5_OK;
}

int tv5_destroy(tv5_ctx* ctx) {
  if (!ctx) return TV5_OK;
  cudaSetDevice(ctx->device);
  Workspace& w = ctx->ws;
  void* ptrs[] = {w.desc, w.state, w.ctl, w.pp, w.E_list, w.P_list, w.n_valid, w.n_roots, w.rec, w.entries, w.hyp,
                  w.hyp_id, w.notin, w.out, w.hyp2, w.hyp_id2, w.out2, w.cand, w.cand_cnt, w.rng_sets, w.h2d_x, w.h2d_sets, w.out_E,
                  w.out_P, w.out_res, w.polish_jobs, w.polish_partial, w.polish_barrier,
                  w.polish_x, w.polish_E, w.flow_jobs, w.flow_x, w.flow_EP};
  for (void* p : ptrs)
    if (p) ws_free(ctx, p);
  if (ctx->rng_u) ws_free(ctx, ctx->rng_u);
  if (ctx->last_done) cudaEventDestroy(ctx->last_done);
  if (ctx->copy_stream) {
    cudaStreamDestroy(ctx->copy_stream);
    for (auto& e : ctx->chunk_ev) cudaEventDestroy(e);
    cudaEventDestroy(ctx->start_ev);
  }
  for (int i = 0; i <= TV5_N_STAGES; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (auto e : ctx->prof_ev) cudaEventDestroy(e);
  for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
  if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
  if (ctx->cap_stream2) cudaStreamDestroy(ctx->cap_stream2);
  if (ctx->cap_fork) cudaEventDestroy(ctx->cap_fork);
  if (ctx->cap_join) cudaEventDestroy(ctx->cap_join);
  if (ctx->front_stream) {
    cudaStreamDestroy(ctx->front_stream);
    cudaStreamDestroy(ctx->back_stream);
    cudaEventDestroy(ctx->pipe_entry);
    cudaEventDestroy(ctx->pipe_done);
    for (auto& e : ctx->pipe_solved) cudaEventDestroy(e);
  }
  delete ctx;
  return TV5_OK;
}

int tv5_last_cuda_error(const tv5_ctx* ctx) { return ctx ? ctx->last_cuda : 0; }
int tv5_device_sm_count(const tv5_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int tv5_ref_rng_sets(tv5_ctx* ctx, void* stream, int N, int iters, int32_t* sets_out) {
  TV5_LOCK(ctx);
  if (!ctx || N < 1 || iters < 1 || !sets_out) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  ref_rng_kernel<<<8, 64, 0, (cudaStream_t)stream>>>(N, iters, sets_out);
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

// uniform draws of the reference RNG for at least `iters` iterations (see ref_rng_uniform_kernel).
// Allocates and synchronises only when a longer table than any seen before is needed.
static int ensure_rng_uniform(tv5_ctx* ctx, cudaStream_t st, int iters, const float** out) {
  if (ctx->rng_u && ctx->rng_u_iters >= iters) { *out = ctx->rng_u; return TV5_OK; }
  if (ctx->rng_u) {
    // the old table may be referenced by queued work and by captured graphs
    TV5_CUDA(ctx, cudaDeviceSynchronize());
    ws_free(ctx, ctx->rng_u);
    ctx->rng_u = nullptr;
    ctx->rng_u_iters = 0;
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
  }
  const int cap = std::max(iters, 8);
  if (ws_malloc(ctx, &ctx->rng_u, (size_t)TV5_REF_THREADS * cap * 5 * sizeof(float)) != cudaSuccess) {
    ctx->rng_u = nullptr;
    ctx->last_cuda = (int)cudaGetLastError();
    return TV5_ERR_NOMEM;
  }
  ref_rng_uniform_kernel<<<8, 64, 0, st>>>(cap * 5, ctx->rng_u);
  TV5_CUDA(ctx, cudaGetLastError());
  TV5_CUDA(ctx, cudaStreamSynchronize(st));   // visible to work later submitted on other streams
  ctx->rng_u_iters = cap;
  *out = ctx->rng_u;
  return TV5_OK;
}

// One context owns one workspace: submissions on different streams are ordered one after the
// other (the new stream waits for the previous submission's completion event), so two streams — or
// two host threads taking turns — never race on it.  Concurrent calls from two host threads on the
// SAME context remain the caller's responsibility (include/tv5.h).
static int submission_enter(tv5_ctx* ctx, cudaStream_t st) {
  if (!ctx->last_done) TV5_CUDA(ctx, cudaEventCreateWithFlags(&ctx->last_done, cudaEventDisableTiming));
  if (ctx->has_last && ctx->last_stream != st) TV5_CUDA(ctx, cudaStreamWaitEvent(st, ctx->last_done, 0));
  return TV5_OK;
}
static int submission_leave(tv5_ctx* ctx, cudaStream_t st) {
  TV5_CUDA(ctx, cudaEventRecord(ctx->last_done, st));
  ctx->last_stream = st;
  ctx->has_last = true;
  return TV5_OK;
}

// ready_ev / ready_first (optional): input chunk k = pairs [ready_first[k], ready_first[k+1]) is
// complete in device memory once ready_ev[k] has fired (host-buffer entry point).
static int pose_batch_impl(tv5_ctx* ctx, void* stream, int B, const double* x1, const double* x2,
                           const int64_t* pt_offsets, const int32_t* sets, int iters, int n_pre,
                           int n_full, double thr, int with_cheirality, double* E_out,
                           double* P_out, tv5_result* result, uint8_t* mask_out,
                           const cudaEvent_t* ready_ev, const int* ready_first, int n_ready) {
  if (!ctx || B < 1 || !x1 || !x2 || !pt_offsets || iters < 1 || !E_out || !result) return TV5_ERR_INVALID;
  if (!(thr > 0.0) || !(thr < 1e300)) return TV5_ERR_INVALID;
  if (B > 65535 || iters > (1 << 20) / TV5_REF_THREADS) return TV5_ERR_INVALID;   // grid.y / index ranges
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  profile_collect(ctx);
  const int H = TV5_REF_THREADS * iters;
  std::vector<PairDesc> hd((size_t)B);
  size_t total_pp = 0;
  int max_pp = 0, max_n = 0;
  bool two_stage = false;
  for (int b = 0; b < B; ++b) {
    const int64_t n64 = pt_offsets[b + 1] - pt_offsets[b];
    if (n64 < 1 || n64 > 0x3fffffff) return TV5_ERR_INVALID;
    const int n = (int)n64;
    PairDesc& d = hd[b];
    d.x1 = x1 + 2 * pt_offsets[b];
    d.x2 = x2 + 2 * pt_offsets[b];
    d.n = n;
    d.n_pre = n_pre <= 0 ? n : std::min(n_pre, n);
    d.n_full = n_full <= 0 ? n : std::min(n_full, n);
    if (d.n_pre != d.n_full) two_stage = true;
    d.sets = sets ? sets + (size_t)b * H * 5 : nullptr;   // reference RNG table: filled in below
    d.E_out = E_out + 9 * (size_t)b;
    d.P_out = P_out ? P_out + 12 * (size_t)b : nullptr;
    d.result = result + b;
    d.mask_out = mask_out ? mask_out + pt_offsets[b] : nullptr;
    d.pp_off = (int64_t)total_pp;
    d.pad = 0;
    const int npp = (n + 1) / 2;
    total_pp += (size_t)npp;
    max_pp = std::max(max_pp, npp);
    max_n = std::max(max_n, n);
  }
  int rc = ensure_workspace(ctx, B, total_pp, (size_t)B * H);
  if (rc) return rc;
  Workspace& w = ctx->ws;
  const float* rng_u = nullptr;
  if (!sets) {
    if ((rc = ensure_rng_uniform(ctx, st, iters, &rng_u))) return rc;
    for (int b = 0; b < B; ++b) hd[b].sets = w.rng_sets + (size_t)b * H * 5;
  }
  if ((rc = submission_enter(ctx, st))) return rc;
  TV5_CUDA(ctx, cudaMemcpyAsync(w.desc, hd.data(), sizeof(PairDesc) * B, cudaMemcpyHostToDevice, st));
  // ---- single pair: the fixed launch sequence is captured once per (N, iterations, flags) into a
  //      CUDA graph and replayed (one graph launch instead of a memset and 13 kernel launches);
  //      the descriptor with the caller's pointers is uploaded outside the graph, so one graph
  //      serves every call of that shape.
  const bool want_graph = ctx->use_graphs && B == 1 && !ready_ev && !ctx->profiling && !ctx->overlap && !two_stage;
  // Graphs are keyed on a BUCKET of N (3 significant bits: 8 buckets per octave), not on N itself:
  // every kernel reads the true point count from the descriptor on the device and only the grid
  // sizes and the tile length derive from the bucket, so SFMnet's keypoint path — whose N changes
  // from pair to pair — replays a handful of graphs instead of issuing 15 launches per call.
  auto bucket_n = [](int n) {
    if (n <= 1024) return (n + 127) & ~127;
    int step = 128;
    while (step * 16 < n) step <<= 1;        // n in (8*step, 16*step]
    return (n + step - 1) / step * step;
  };
  const int grid_n0 = want_graph ? bucket_n(hd[0].n) : hd[0].n;
  GraphKey gkey{grid_n0, iters, with_cheirality, (int)ctx->split_solver | ((int)ctx->early_exit << 1) |
                                                      ((int)ctx->force_exact << 2) | ((P_out != nullptr) << 3) | ((sets == nullptr) << 4), thr};
  bool build_graph = false;
  if (want_graph) {
    for (auto& g : ctx->graphs)
      if (g.key == gkey) {
        TV5_CUDA(ctx, cudaGraphLaunch(g.exec, st));
        return submission_leave(ctx, st);
      }
    // capturing costs about a millisecond: only shapes that keep coming back are captured (SFMnet's
    // keypoint counts change from pair to pair; its dense crop does not)
    bool seen = false;
    for (auto& e : ctx->graph_seen)
      if (e.key == gkey) { seen = true; build_graph = ++e.count >= 3; break; }
    if (!seen) {
      if (ctx->graph_seen.size() >= 64) ctx->graph_seen.erase(ctx->graph_seen.begin());
      ctx->graph_seen.push_back({gkey, 1});
    }
  }
  cudaStream_t sx = st;   // the stream the work is enqueued on (the capture stream while a graph is built)
  bool capturing = false;
  auto enqueue_all = [&]() -> int {
  TV5_CUDA(ctx, cudaMemsetAsync(w.state, 0, sizeof(PairState) * B, sx));
    const int allow_fast = (two_stage || ctx->force_exact) ? 0 : 1;
  
    // A submission may be cut into chunks of pairs, each with a front (prep + five-point solve:
    // float64, latency bound) and a back (scoring: float32 pipe bound, + selection).
    //  * host-buffer entry point: chunk c runs as soon as its copies have landed (ready_ev), in
    //    plain order front(c), back(c) on the caller's stream;
    //  * tv5_set_overlap(1): fronts on a low-priority and backs on a high-priority internal stream,
    //    so the solver of chunk c+1 runs in the shadow of the scorer of chunk c.  Measured on B200
    //    (DESIGN.md section 4.4): no gain — solver warps make almost no progress next to the
    //    FFMA2-saturating scorer — hence off by default.
    int n_chunks = 1;
    if (ready_ev && !two_stage) n_chunks = n_ready;        // compute chunks = copy chunks
    else if (ctx->overlap && !two_stage) n_chunks = std::max(1, std::min(kPipeChunks, B / kPipeMinPairs));
    const bool two_streams = ctx->overlap && n_chunks > 1;
    cudaStream_t s_front = sx, s_back = sx;
    if (two_streams) {
      if ((rc = ensure_pipe_streams(ctx))) return rc;
      s_front = ctx->front_stream;
      s_back = ctx->back_stream;
      TV5_CUDA(ctx, cudaEventRecord(ctx->pipe_entry, sx));
      TV5_CUDA(ctx, cudaStreamWaitEvent(s_front, ctx->pipe_entry, 0));
      TV5_CUDA(ctx, cudaStreamWaitEvent(s_back, ctx->pipe_entry, 0));
    }
    std::vector<int> first((size_t)n_chunks + 1);
    for (int c = 0; c <= n_chunks; ++c)
      first[c] = (ready_ev && n_chunks == n_ready) ? ready_first[c] : (int)((int64_t)B * c / n_chunks);
    const int slots = TV5_SCORE_MINB * ctx->sm_count;
    const bool prof = ctx->profiling;
    if (prof && (rc = ensure_prof_events(ctx, n_chunks))) return rc;
    ctx->prof_chunks = prof ? n_chunks : 0;
  
    // ---- front: prep + solve of one chunk
    auto front = [&](int c) -> int {
      const int b0 = first[c], nb = first[c + 1] - b0;
      int cmax_pp = 0;
      for (int b = b0; b < b0 + nb; ++b) cmax_pp = std::max(cmax_pp, ((B == 1 ? grid_n0 : hd[b].n) + 1) / 2);
      const PairDesc* desc = w.desc + b0;
      PairState* state = w.state + b0;
      const size_t so = (size_t)b0 * H;
      if (ready_ev) {  // inputs of this chunk: every host chunk up to the one holding its last pair
        int k = 0;
        while (k + 1 < n_ready && ready_first[k + 1] < b0 + nb) ++k;
        for (int j = (c == 0 ? 0 : k); j <= k; ++j) TV5_CUDA(ctx, cudaStreamWaitEvent(s_front, ready_ev[j], 0));
      }
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 0], s_front);
      // While a single-pair graph is being captured the point preparation becomes a branch of its
      // own: nothing before solve_poses (which needs the pair's scale from band_consts) depends on
      // it, so in the replayed graph it runs next to the table scaling, solve_front and solve_roots
      // instead of in front of them (about 10 us of a 170 us call).
      const bool fork = capturing && ctx->split_solver && ctx->cap_stream2;
      cudaStream_t s_prep = s_front;
      if (fork) {
        TV5_CUDA(ctx, cudaEventRecord(ctx->cap_fork, s_front));
        TV5_CUDA(ctx, cudaStreamWaitEvent(ctx->cap_stream2, ctx->cap_fork, 0));
        s_prep = ctx->cap_stream2;
      }
      prep_norms<<<dim3((cmax_pp + 255) / 256, nb), 256, 0, s_prep>>>(desc, state);
      band_consts<<<(nb + 127) / 128, 128, 0, s_prep>>>(desc, state, nb, thr, allow_fast);
      if (allow_fast) prep_points<<<dim3((cmax_pp + 255) / 256, nb), 256, 0, s_prep>>>(desc, state, w.pp);
      if (fork) TV5_CUDA(ctx, cudaEventRecord(ctx->cap_join, s_prep));
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 1], s_front);
      if (rng_u) rng_scale_sets<<<dim3((H * 5 + 255) / 256, nb), 256, 0, s_front>>>(desc, H, iters, rng_u,
                                                                                   w.rng_sets + so * 5);
      if (ctx->split_solver) {
        const int spw = front_sets_per_warp((int64_t)nb * H, ctx->sm_count);
        launch_solve_front(spw, H, nb, s_front, desc, w.rec + so * kRecDoubles);
        solve_roots<<<dim3((H + 63) / 64, nb), 64, 0, s_front>>>(state, H, w.rec + so * kRecDoubles,
                                                                 (RootEntry*)w.entries + so * 10, w.n_roots + so,
                                                                 w.n_valid + so);
        if (fork) TV5_CUDA(ctx, cudaStreamWaitEvent(s_front, ctx->cap_join, 0));
        solve_poses<<<dim3((H * 10 + 127) / 128, nb), 128, 0, s_front>>>(
            desc, state, H, with_cheirality, w.rec + so * kRecDoubles, (const RootEntry*)w.entries + so * 10,
            w.E_list + so * 90, with_cheirality ? w.P_list + so * 120 : nullptr, w.n_valid + so, w.hyp + so * 10,
            w.hyp_id + so * 10, w.notin + so * 10, w.out + so * 10);
      } else {
        solve_sets<<<dim3((H + 31) / 32, nb), 32, 0, s_front>>>(
            desc, state, H, with_cheirality, w.E_list + so * 90, with_cheirality ? w.P_list + so * 120 : nullptr,
            w.n_valid + so, w.n_roots + so, w.hyp + so * 10, w.hyp_id + so * 10, w.notin + so * 10, w.out + so * 10);
      }
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 2], s_front);
      if (two_streams) TV5_CUDA(ctx, cudaEventRecord(ctx->pipe_solved[c], s_front));
      return TV5_OK;
    };
    // ---- back: scoring + selection of one chunk
    auto back = [&](int c) -> int {
      const int b0 = first[c], nb = first[c + 1] - b0;
      int cmax_pp = 0;
      for (int b = b0; b < b0 + nb; ++b) cmax_pp = std::max(cmax_pp, ((B == 1 ? grid_n0 : hd[b].n) + 1) / 2);
      PairDesc* desc = w.desc + b0;
      PairState* state = w.state + b0;
      Control* ctl = w.ctl + c;
      const size_t so = (size_t)b0 * H;
      const double* E_list = w.E_list + so * 90;
      const double* P_list = with_cheirality ? w.P_list + so * 120 : nullptr;
      const Hyp32* hyp = w.hyp + so * 10;
      int32_t* hyp_id = w.hyp_id + so * 10;
      uint32_t* notin = w.notin + so * 10;
      uint32_t* out = w.out + so * 10;
      int32_t* cand = w.cand + so * 10;
      int32_t* cand_cnt = w.cand_cnt + so * 10;
      if (two_streams) TV5_CUDA(ctx, cudaStreamWaitEvent(s_back, ctx->pipe_solved[c], 0));
      // tile size: big tiles for batches, enough tiles to balance 2 CTAs/SM for a single pair
      int pp_per_tile = kMaxTilePairs;
      {
        const double est_M = (double)H * (with_cheirality ? 3.0 : 4.5);
        const double n_hc = std::max(1.0, est_M / kHypChunk);
        const double want_pc = 6.0 * slots / (n_hc * nb);
        if (want_pc > 1.0) {
          int t = (int)((double)cmax_pp / want_pc);
          t = (t + 7) & ~7;
          pp_per_tile = std::max(64, std::min(kMaxTilePairs, t));
        }
      }
      const int X = std::max(1, std::min(4096, (4 * ctx->sm_count + nb - 1) / nb));
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 3], s_back);
      // staging only pays when the scoring work dwarfs the extra launches (~10 small kernels)
      const bool staged = ctx->early_exit && allow_fast && !two_stage &&
                          (double)nb * H * 3.0 * (2.0 * cmax_pp) >= 2.0e8;
      if (staged) {
        // staged scoring with exact pruning (see set_stage / stage_leader / prune_compact)
        if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 4], s_back);
        const Hyp32* alt_hyp = w.hyp2 + so * 10;
        const int32_t* alt_id = w.hyp_id2 + so * 10;
        const uint32_t* alt_out = w.out2 + so * 10;
        const int n_stages = ctx->early_stages;
        for (int stg = 0; stg < n_stages; ++stg) {
          set_stage<<<(nb + 127) / 128, 128, 0, s_back>>>(desc, state, nb, ctx->early_frac[stg], ctx->early_frac[stg + 1], stg == 0,
                                                          pp_per_tile);
          if (stg == 0) {
            plan_tiles<<<1, 1024, 0, s_back>>>(desc, state, ctl, nb, pp_per_tile, kHypChunk);
            score_bounds<false><<<slots, kScoreThreads, 0, s_back>>>(desc, state, ctl, nb, H, pp_per_tile, w.pp, hyp,
                                                                    notin, out);
          } else {  // few survivors per pair: 256-slot hypothesis chunks
            plan_tiles<<<1, 1024, 0, s_back>>>(desc, state, ctl, nb, pp_per_tile, kScoreThreads);
            score_bounds<false, 1><<<4 * ctx->sm_count, kScoreThreads, 0, s_back>>>(desc, state, ctl, nb, H, pp_per_tile,
                                                                                   w.pp, hyp, notin, out);
          }
          if (stg + 1 < n_stages) {
            stage_leader<<<nb, 1024, 0, s_back>>>(desc, state, H, out, hyp_id, cand, cand_cnt);
            exact_counts<<<dim3(X, nb), 256, 0, s_back>>>(desc, state, H, thr, E_list, hyp_id, cand, cand_cnt);
            prune_compact<<<nb, 256, 0, s_back>>>(desc, state, H, hyp, hyp_id, out, cand_cnt, (Hyp32*)alt_hyp,
                                                  (int32_t*)alt_id, (uint32_t*)alt_out);
            std::swap(hyp, alt_hyp);
            { const int32_t* t = hyp_id; hyp_id = (int32_t*)alt_id; alt_id = t; }
            { const uint32_t* t = out; out = (uint32_t*)alt_out; alt_out = t; }
          }
        }
        if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 5], s_back);
      } else {
        plan_tiles<<<1, 1024, 0, s_back>>>(desc, state, ctl, nb, pp_per_tile, kHypChunk);
        if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 4], s_back);
        if (allow_fast) {
          // few hypothesis chunks per pair (a hypothesis shard of one large pair, a small budget): the
          // partly filled last chunk is a noticeable share -> the PARTIAL instantiation
          const bool few_chunks = (double)H * (with_cheirality ? 3.0 : 4.5) <= 8.0 * kHypChunk;
          if (few_chunks)
            score_bounds<false, kHypPerThread, true><<<slots, kScoreThreads, 0, s_back>>>(desc, state, ctl, nb, H, pp_per_tile,
                                                                                         w.pp, hyp, notin, out);
          else
            score_bounds<false><<<slots, kScoreThreads, 0, s_back>>>(desc, state, ctl, nb, H, pp_per_tile, w.pp, hyp,
                                                                    notin, out);
        }
        if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 5], s_back);
      }
      pick_top<<<nb, 1024, 0, s_back>>>(desc, state, H, out, hyp_id, cand, cand_cnt);
      if (two_stage) {
        // (single chunk) stage A runs on n_pre points: exact_counts reads n_full, so the descriptors
        // are re-uploaded with n_full := n_pre for this stage and restored afterwards.
        std::vector<PairDesc> ha = hd;
        for (auto& d : ha) d.n_full = d.n_pre;
        TV5_CUDA(ctx, cudaMemcpyAsync(w.desc, ha.data(), sizeof(PairDesc) * B, cudaMemcpyHostToDevice, s_back));
        exact_counts<<<dim3(X, nb), 256, 0, s_back>>>(desc, state, H, thr, E_list, hyp_id, cand, cand_cnt);
        set_winners<<<nb, 256, 0, s_back>>>(desc, state, H, hyp_id, cand, cand_cnt, w.cand + w.sets_cap * 10);
        TV5_CUDA(ctx, cudaMemcpyAsync(w.desc, hd.data(), sizeof(PairDesc) * B, cudaMemcpyHostToDevice, s_back));
        exact_counts<<<dim3(X, nb), 256, 0, s_back>>>(desc, state, H, thr, E_list, hyp_id, cand, cand_cnt);
      } else {
        exact_counts<<<dim3(X, nb), 256, 0, s_back>>>(desc, state, H, thr, E_list, hyp_id, cand, cand_cnt);
        if (allow_fast) {
          pick_rest<<<nb, 1024, 0, s_back>>>(desc, state, H, out, cand, cand_cnt);
          exact_counts<<<dim3(X, nb), 256, 0, s_back>>>(desc, state, H, thr, E_list, hyp_id, cand, cand_cnt);
        }
      }
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 6], s_back);
      finalize<<<nb, 256, 0, s_back>>>(desc, state, H, thr, E_list, P_list, hyp_id, cand, cand_cnt,
                                       ctx->split_solver ? w.n_valid + so : nullptr);
      if (prof) cudaEventRecord(ctx->prof_ev[c * kProfPerChunk + 7], s_back);
      return TV5_OK;
    };
    if (two_streams) {
      for (int c = 0; c < n_chunks; ++c)
        if ((rc = front(c))) return rc;
      for (int c = 0; c < n_chunks; ++c)
        if ((rc = back(c))) return rc;
    } else {
      for (int c = 0; c < n_chunks; ++c) {
        if ((rc = front(c))) return rc;
        if ((rc = back(c))) return rc;
      }
    }
    if (two_streams) {
      TV5_CUDA(ctx, cudaEventRecord(ctx->pipe_done, s_back));
      TV5_CUDA(ctx, cudaStreamWaitEvent(sx, ctx->pipe_done, 0));
    }
  if (prof) ctx->ev_pending = true;
    TV5_CUDA(ctx, cudaGetLastError());
    return TV5_OK;
  };
  if (want_graph && build_graph) {
    if (!ctx->cap_stream && cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      ctx->use_graphs = false;
    } else if (cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      if (!ctx->cap_stream2 && (cudaStreamCreateWithFlags(&ctx->cap_stream2, cudaStreamNonBlocking) != cudaSuccess ||
                                cudaEventCreateWithFlags(&ctx->cap_fork, cudaEventDisableTiming) != cudaSuccess ||
                                cudaEventCreateWithFlags(&ctx->cap_join, cudaEventDisableTiming) != cudaSuccess)) {
        cudaGetLastError();
        ctx->cap_stream2 = nullptr;   // no fork: the graph is captured as one chain
      }
      sx = ctx->cap_stream;
      capturing = true;
      const int rc_cap = enqueue_all();
      capturing = false;
      cudaGraph_t graph = nullptr;
      const cudaError_t e_end = cudaStreamEndCapture(ctx->cap_stream, &graph);
      cudaGraphExec_t exec = nullptr;
      if (rc_cap == TV5_OK && e_end == cudaSuccess && graph &&
          cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphDestroy(graph);
        if (ctx->graphs.size() >= 64) { cudaGraphExecDestroy(ctx->graphs.front().exec); ctx->graphs.erase(ctx->graphs.begin()); }
        ctx->graphs.push_back({gkey, exec});
        TV5_CUDA(ctx, cudaGraphLaunch(exec, st));
        return submission_leave(ctx, st);
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->use_graphs = false;   // capture is not possible here: plain launches from now on
      sx = st;
    } else {
      cudaGetLastError();
      ctx->use_graphs = false;
    }
  }
  if ((rc = enqueue_all())) return rc;
  return submission_leave(ctx, st);
}

int tv5_compute_pose_batch(tv5_ctx* ctx, void* stream, int B, const double* x1, const double* x2,
                           const int64_t* pt_offsets, const int32_t* sets, int iters, int n_pre,
                           int n_full, double thr, int with_cheirality, double* E_out,
                           double* P_out, tv5_result* result, uint8_t* mask_out) {
  TV5_LOCK(ctx);
  return pose_batch_impl(ctx, stream, B, x1, x2, pt_offsets, sets, iters, n_pre, n_full, thr, with_cheirality,
                         E_out, P_out, result, mask_out, nullptr, nullptr, 0);
}

int tv5_compute_pose(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                     const int32_t* sets, int iters, int n_pre, int n_full, double thr,
                     int with_cheirality, double* E_out, double* P_out, tv5_result* result,
                     uint8_t* mask_out) {
  if (N < 1) return TV5_ERR_INVALID;
  const int64_t off[2] = {0, N};
  return tv5_compute_pose_batch(ctx, stream, 1, x1, x2, off, sets, iters, n_pre, n_full, thr,
                                with_cheirality, E_out, P_out, result, mask_out);
}

int tv5_compute_pose_batch_host(tv5_ctx* ctx, void* stream, int B, const double* x1,
                                const double* x2, const int64_t* pt_offsets, const int32_t* sets,
                                int iters, int n_pre, int n_full, double thr, int with_cheirality,
                                double* E_out, double* P_out, tv5_result* result) {
  TV5_LOCK(ctx);
  if (!ctx || B < 1 || !x1 || !x2 || !pt_offsets || !E_out || !result || iters < 1) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t total = (size_t)(pt_offsets[B] - pt_offsets[0]);
  const int H = TV5_REF_THREADS * iters;
  Workspace& w = ctx->ws;
  int rc;
  if ((rc = grow(ctx, w.h2d_x, w.h2d_cap, total * 4))) return rc;
  if (sets && (rc = grow(ctx, w.h2d_sets, w.h2d_sets_cap, (size_t)B * H * 5))) return rc;
  if ((size_t)B > w.out_cap || !w.out_E) {
    if ((rc = grow_same(ctx, w.out_E, w.out_cap * 9, (size_t)B * 9))) return rc;
    if ((rc = grow_same(ctx, w.out_P, w.out_cap * 12, (size_t)B * 12))) return rc;
    if ((rc = grow_same(ctx, w.out_res, w.out_cap, (size_t)B))) return rc;
    w.out_cap = (size_t)B;
  }
  if (!ctx->copy_stream) {
    TV5_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->chunk_ev) TV5_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    TV5_CUDA(ctx, cudaEventCreateWithFlags(&ctx->start_ev, cudaEventDisableTiming));
  }
  // Pipeline: the batch is cut into chunks of pairs; the host->device copies of chunk k+1 (own
  // stream) overlap the kernels of chunk k (caller's stream).  The copies are ~10x faster than the
  // kernels, so the chunks grow geometrically (1/32, 7/32, 3/4 of the batch): the first kernels
  // start after 1/32 of the copy time and only three launches of each kernel are made (measured
  // round 2, 256 pairs: 19.19 ms against 19.40 for 1/16, 3/16, 3/4; a single early chunk stalls on PCIe).
  std::vector<int> first;
  first.push_back(0);
  if (const char* e = getenv("TV5_HOST_CHUNKS")) {   // dev knob: interior chunk boundaries as fractions, e.g. "0.03,0.2"
    const char* q = e;
    while (*q) {
      char* end = nullptr;
      const double f = strtod(q, &end);
      if (end == q) break;
      const int b = (int)(f * B);
      if (b > first.back() && b < B) first.push_back(b);
      q = (*end == ',') ? end + 1 : end;
    }
  } else if (B >= 64) { first.push_back(B / 32); first.push_back(B / 4); }
  else if (B >= 2) first.push_back(B / 2);
  first.push_back(B);
  const int n_chunks = (int)first.size() - 1;
  double* dx1 = w.h2d_x;
  double* dx2 = w.h2d_x + total * 2;
  TV5_CUDA(ctx, cudaEventRecord(ctx->start_ev, st));
  TV5_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->start_ev, 0));
  for (int k = 0; k < n_chunks; ++k) {
    const int b0 = first[k], b1 = first[k + 1];
    const size_t p0 = (size_t)(pt_offsets[b0] - pt_offsets[0]), np = (size_t)(pt_offsets[b1] - pt_offsets[b0]);
    TV5_CUDA(ctx, cudaMemcpyAsync(dx1 + 2 * p0, x1 + 2 * pt_offsets[b0], np * 2 * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
    TV5_CUDA(ctx, cudaMemcpyAsync(dx2 + 2 * p0, x2 + 2 * pt_offsets[b0], np * 2 * sizeof(double),
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
    if (sets)
      TV5_CUDA(ctx, cudaMemcpyAsync(w.h2d_sets + (size_t)b0 * H * 5, sets + (size_t)b0 * H * 5,
                                    (size_t)(b1 - b0) * H * 5 * sizeof(int32_t), cudaMemcpyHostToDevice,
                                    ctx->copy_stream));
    TV5_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[k], ctx->copy_stream));
  }
  // one submission; each chunk of pairs starts as soon as its copies have landed
  std::vector<int64_t> off((size_t)B + 1);
  for (int b = 0; b <= B; ++b) off[b] = pt_offsets[b] - pt_offsets[0];
  rc = pose_batch_impl(ctx, stream, B, dx1, dx2, off.data(), sets ? w.h2d_sets : nullptr, iters, n_pre, n_full, thr,
                       with_cheirality, w.out_E, w.out_P, w.out_res, nullptr, ctx->chunk_ev, first.data(), n_chunks);
  if (rc) return rc;
  TV5_CUDA(ctx, cudaMemcpyAsync(E_out, w.out_E, (size_t)B * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (P_out) TV5_CUDA(ctx, cudaMemcpyAsync(P_out, w.out_P, (size_t)B * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TV5_CUDA(ctx, cudaMemcpyAsync(result, w.out_res, (size_t)B * sizeof(tv5_result), cudaMemcpyDeviceToHost, st));
  TV5_CUDA(ctx, cudaStreamSynchronize(st));
  return TV5_OK;
}

int tv5_solve5(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
               const int32_t* sets, int H, int with_cheirality, double* E_list, double* P_list,
               int32_t* n_roots, int32_t* n_valid) {
  TV5_LOCK(ctx);
  if (!ctx || !x1 || !x2 || N < 1 || !sets || H < 1 || !E_list || !n_valid) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_workspace(ctx, 1, 1, ctx->split_solver ? (size_t)H : 1);
  if (rc) return rc;
  if ((rc = submission_enter(ctx, st))) return rc;
  PairDesc d;
  memset(&d, 0, sizeof(d));
  d.x1 = x1; d.x2 = x2; d.sets = sets; d.n = N; d.n_pre = d.n_full = N;
  TV5_CUDA(ctx, cudaMemcpyAsync(ctx->ws.desc, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  TV5_CUDA(ctx, cudaMemsetAsync(E_list, 0, (size_t)H * 90 * sizeof(double), st));
  if (P_list) TV5_CUDA(ctx, cudaMemsetAsync(P_list, 0, (size_t)H * 120 * sizeof(double), st));
  if (ctx->split_solver) {
    Workspace& w = ctx->ws;
    TV5_CUDA(ctx, cudaMemsetAsync(w.state, 0, sizeof(PairState), st));
    const int spw = front_sets_per_warp(H, ctx->sm_count);
    launch_solve_front(spw, H, 1, st, w.desc, w.rec);
    solve_roots<<<dim3((H + 63) / 64, 1), 64, 0, st>>>(w.state, H, w.rec, (RootEntry*)w.entries, n_roots, n_valid);
    solve_poses<<<dim3((H * 10 + 127) / 128, 1), 128, 0, st>>>(w.desc, w.state, H, with_cheirality, w.rec,
                                                              (const RootEntry*)w.entries, E_list,
                                                              with_cheirality ? P_list : nullptr, n_valid, nullptr,
                                                              nullptr, nullptr, nullptr);
    compact_solutions<<<(H + 127) / 128, 128, 0, st>>>(H, E_list, with_cheirality ? P_list : nullptr, n_valid);
  } else {
    solve_sets<<<dim3((H + 31) / 32, 1), 32, 0, st>>>(ctx->ws.desc, ctx->ws.state, H, with_cheirality,
                                                     E_list, with_cheirality ? P_list : nullptr, n_valid,
                                                     n_roots, nullptr, nullptr, nullptr, nullptr);
  }
  TV5_CUDA(ctx, cudaGetLastError());
  return submission_leave(ctx, st);
}

int tv5_score(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int n_test,
              const double* E_list, int M, double thr, int32_t* counts, uint32_t* masks) {
  TV5_LOCK(ctx);
  if (!ctx || n_test < 0 || M < 0) return TV5_ERR_INVALID;
  if (M == 0) return TV5_OK;
  if (!E_list || !counts || (n_test > 0 && (!x1 || !x2))) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  TV5_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)M * sizeof(int32_t), st));
  if (n_test == 0) return TV5_OK;
  for (int m0 = 0; m0 < M; m0 += 65535) {
    const int mc = std::min(65535, M - m0);
    score_exact_list<<<dim3((n_test + 1023) / 1024, mc), 256, 0, st>>>(
        x1, x2, n_test, E_list + 9 * (size_t)m0, thr, counts + m0,
        masks ? masks + (size_t)m0 * ((n_test + 31) / 32) : nullptr);
  }
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

int tv5_score_bounds(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int n_test,
                     const double* E_list, int M, double thr, int32_t* lo, int32_t* hi) {
  TV5_LOCK(ctx);
  if (!ctx || !x1 || !x2 || n_test < 1 || !E_list || M < 1 || !lo || !hi || !(thr > 0.0)) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  profile_collect(ctx);
  const int H = (M + 9) / 10;  // workspace is sized in sets of 10 hypotheses
  int rc = ensure_workspace(ctx, 1, (size_t)(n_test + 1) / 2, (size_t)H);
  if (rc) return rc;
  Workspace& w = ctx->ws;
  if ((rc = submission_enter(ctx, st))) return rc;
  PairDesc d;
  memset(&d, 0, sizeof(d));
  d.x1 = x1; d.x2 = x2; d.n = n_test; d.n_pre = d.n_full = n_test; d.pp_off = 0;
  TV5_CUDA(ctx, cudaMemcpyAsync(w.desc, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  TV5_CUDA(ctx, cudaMemsetAsync(w.state, 0, sizeof(PairState), st));
  const int npp = (n_test + 1) / 2;
  stage_mark(ctx, st, 0);
  prep_norms<<<dim3((npp + 255) / 256, 1), 256, 0, st>>>(w.desc, w.state);
  band_consts<<<1, 32, 0, st>>>(w.desc, w.state, 1, thr, 1);
  prep_points<<<dim3((npp + 255) / 256, 1), 256, 0, st>>>(w.desc, w.state, w.pp);
  stage_mark(ctx, st, 1);
  hyps_from_list<<<(M + 255) / 256, 256, 0, st>>>(E_list, M, w.state, w.hyp, w.hyp_id, w.notin, w.out);
  stage_mark(ctx, st, 2);
  const int slots = TV5_SCORE_MINB * ctx->sm_count;
  int pp_per_tile = kMaxTilePairs;
  {
    const double n_hc = std::max(1.0, (double)M / kHypChunk);
    const double want_pc = 6.0 * slots / n_hc;
    if (want_pc > 1.0) {
      int t = (int)((double)npp / want_pc);
      t = (t + 7) & ~7;
      pp_per_tile = std::max(64, std::min(kMaxTilePairs, t));
    }
  }
  plan_tiles<<<1, 1024, 0, st>>>(w.desc, w.state, w.ctl, 1, pp_per_tile, kHypChunk);
  stage_mark(ctx, st, 3);
  // H*10 is the stride between pairs inside the kernel; with one pair it is irrelevant
  score_bounds<true><<<slots, kScoreThreads, 0, st>>>(w.desc, w.state, w.ctl, 1, H, pp_per_tile, w.pp, w.hyp,
                                                      w.notin, w.out);
  stage_mark(ctx, st, 4);
  bounds_to_lo_hi<<<(M + 255) / 256, 256, 0, st>>>(w.notin, w.out, M, n_test, lo, hi);
  stage_mark(ctx, st, 5);
  stage_mark(ctx, st, 6);
  if (ctx->profiling) ctx->ev_pending = true;
  TV5_CUDA(ctx, cudaGetLastError());
  // the fast path may have been refused on the device (non-finite / huge coordinates)
  PairState hs;
  TV5_CUDA(ctx, cudaMemcpyAsync(&hs, w.state, sizeof(hs), cudaMemcpyDeviceToHost, st));
  TV5_CUDA(ctx, cudaStreamSynchronize(st));
  if ((rc = submission_leave(ctx, st))) return rc;
  return hs.fast ? TV5_OK : TV5_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------
// decomposition and refinement (polish.cuh)
// ------------------------------------------------------------------------------------------
int tv5_decompose(const double* E, double* angles) {
  if (!E || !angles) return TV5_ERR_INVALID;
  givens_to_angles(givens_decompose(E), angles);
  return TV5_OK;
}

int tv5_decompose_uv(const double* E, double* U, double* V) {
  if (!E || !U || !V) return TV5_ERR_INVALID;
  givens_to_uv(givens_decompose(E), U, V);
  return TV5_OK;
}

int tv5_decompose_batch(tv5_ctx* ctx, void* stream, const double* E, int B, double* angles, double* U,
                        double* V) {
  TV5_LOCK(ctx);
  if (!ctx || !E || B < 0 || (!angles && !U && !V)) return TV5_ERR_INVALID;
  if (B == 0) return TV5_OK;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  decompose_batch<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(E, B, angles, U, V);
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

int tv5_optimise_batch(tv5_ctx* ctx, void* stream, int B, const double* x1, const double* x2,
                       const int64_t* pt_offsets, const uint8_t* mask, double* E_io, double delta,
                       double alpha, int max_reps, int32_t* iters_out) {
  TV5_LOCK(ctx);
  if (!ctx || B < 1 || !pt_offsets || !E_io || max_reps < 0) return TV5_ERR_INVALID;
  max_reps = std::min(max_reps, 1000000);   // keeps the monotonic barrier counter (reps x CTAs) inside 32 bits
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->polish_max_ctas) {
    int per_sm = 0;
    TV5_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, irls_polish, kPolishThreads, 0));
    ctx->polish_max_ctas = std::max(1, per_sm) * ctx->sm_count;
  }
  Workspace& w = ctx->ws;
  int rc;
  if ((rc = submission_enter(ctx, st))) return rc;
  // jobs go out in waves of at most polish_max_ctas CTAs (a cooperative launch must be co-resident)
  for (int b0 = 0; b0 < B;) {
    const int nb = std::min(B - b0, ctx->polish_max_ctas);
    int64_t max_n = 0;
    for (int b = b0; b < b0 + nb; ++b) {
      const int64_t n = pt_offsets[b + 1] - pt_offsets[b];
      if (n < 0 || n > 0x3fffffff || (n > 0 && (!x1 || !x2))) return TV5_ERR_INVALID;
      max_n = std::max(max_n, n);
    }
    // about four points per thread, never more CTAs than fit
    int G = (int)std::min<int64_t>((max_n + 4 * kPolishThreads - 1) / (4 * kPolishThreads),
                                   (int64_t)(ctx->polish_max_ctas / nb));
    G = std::max(G, 1);
    PolishJob* jobs = (PolishJob*)w.polish_jobs;
    if ((size_t)nb > w.polish_jobs_cap || !jobs) {
      if ((rc = grow_same(ctx, jobs, w.polish_jobs_cap, (size_t)nb))) { w.polish_jobs = nullptr; w.polish_jobs_cap = 0; return rc; }
      w.polish_jobs = jobs;
      if ((rc = grow_same(ctx, w.polish_barrier, w.polish_jobs_cap, (size_t)nb))) return rc;
      w.polish_jobs_cap = (size_t)nb;
    }
    if ((rc = grow(ctx, w.polish_partial, w.polish_partial_cap, (size_t)nb * 2 * G * kPolishTerms))) return rc;
    std::vector<PolishJob> hj((size_t)nb);
    for (int b = 0; b < nb; ++b) {
      const int64_t o = pt_offsets[b0 + b];
      hj[b].x1 = x1 + 2 * o;
      hj[b].x2 = x2 + 2 * o;
      hj[b].mask = mask ? mask + o : nullptr;
      hj[b].E = E_io + 9 * (size_t)(b0 + b);
      hj[b].iters_out = iters_out ? iters_out + b0 + b : nullptr;
      hj[b].n = (int32_t)(pt_offsets[b0 + b + 1] - o);
      hj[b].pad = 0;
    }
    TV5_CUDA(ctx, cudaMemcpyAsync(jobs, hj.data(), sizeof(PolishJob) * nb, cudaMemcpyHostToDevice, st));
    TV5_CUDA(ctx, cudaMemsetAsync(w.polish_barrier, 0, sizeof(unsigned int) * nb, st));
    const PolishJob* jp = jobs;
    double* pp = w.polish_partial;
    unsigned int* bp = w.polish_barrier;
    void* args[] = {(void*)&jp, (void*)&G, (void*)&delta, (void*)&alpha, (void*)&max_reps, (void*)&pp, (void*)&bp};
    TV5_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)irls_polish, dim3(nb * G), dim3(kPolishThreads), args, 0, st));
    b0 += nb;
  }
  return submission_leave(ctx, st);
}

int tv5_optimise(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                 const uint8_t* mask, double* E_io, double delta, double alpha, int max_reps,
                 int32_t* iters_out) {
  if (N < 0) return TV5_ERR_INVALID;
  const int64_t off[2] = {0, N};
  return tv5_optimise_batch(ctx, stream, 1, x1, x2, off, mask, E_io, delta, alpha, max_reps, iters_out);
}

int tv5_optimise_host(tv5_ctx* ctx, void* stream, const double* x1, const double* x2, int N,
                      const double* E_init, double delta, double alpha, int max_reps, double* E_out) {
  TV5_LOCK(ctx);
  if (!ctx || N < 0 || !E_init || !E_out || (N > 0 && (!x1 || !x2))) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  Workspace& w = ctx->ws;
  int rc;
  if ((rc = grow(ctx, w.polish_x, w.polish_x_cap, (size_t)std::max(N, 1) * 4))) return rc;
  if (!w.polish_E && ws_malloc(ctx, &w.polish_E, 9 * sizeof(double)) != cudaSuccess) { w.polish_E = nullptr; return TV5_ERR_NOMEM; }
  if (N > 0) {
    TV5_CUDA(ctx, cudaMemcpyAsync(w.polish_x, x1, (size_t)N * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    TV5_CUDA(ctx, cudaMemcpyAsync(w.polish_x + 2 * (size_t)N, x2, (size_t)N * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  TV5_CUDA(ctx, cudaMemcpyAsync(w.polish_E, E_init, 9 * sizeof(double), cudaMemcpyHostToDevice, st));
  rc = tv5_optimise(ctx, stream, w.polish_x, w.polish_x + 2 * (size_t)N, N, nullptr, w.polish_E, delta, alpha,
                    max_reps, nullptr);
  if (rc) return rc;
  TV5_CUDA(ctx, cudaMemcpyAsync(E_out, w.polish_E, 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TV5_CUDA(ctx, cudaStreamSynchronize(st));
  return TV5_OK;
}

// ------------------------------------------------------------------------------------------
// optical flow -> correspondences (flow_points.cuh)
// ------------------------------------------------------------------------------------------
static int flow_offsets(int B, int H, int W, int mode, int margin, const int64_t* pt_offsets,
                        std::vector<int64_t>& off) {
  off.assign((size_t)B + 1, 0);
  if (mode == kFlowCrop) {
    if (margin < 0 || 2 * margin >= H || 2 * margin >= W) return TV5_ERR_INVALID;
    const int64_t n = (int64_t)(H - 2 * margin) * (W - 2 * margin);
    for (int b = 0; b <= B; ++b) off[b] = n * b;
  } else {
    if (!pt_offsets) return TV5_ERR_INVALID;
    for (int b = 0; b <= B; ++b) off[b] = pt_offsets[b] - pt_offsets[0];
    for (int b = 0; b < B; ++b)
      if (off[b + 1] < off[b]) return TV5_ERR_INVALID;
  }
  return off[B] > 0x3fffffff ? TV5_ERR_INVALID : TV5_OK;
}

int tv5_flow_to_points(tv5_ctx* ctx, void* stream, const float* flow, int B, int H, int W,
                       const float* Kinv, int mode, int margin, const void* pts,
                       const int64_t* pt_offsets, double* x1_out, double* x2_out) {
  TV5_LOCK(ctx);
  if (!ctx || !flow || B < 1 || H < 1 || W < 1 || !Kinv || mode < 0 || mode > 2 || !x1_out || !x2_out)
    return TV5_ERR_INVALID;
  if (mode != kFlowCrop && !pts) return TV5_ERR_INVALID;
  if (B > 65535) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> off;
  int rc = flow_offsets(B, H, W, mode, margin, pt_offsets, off);
  if (rc) return rc;
  Workspace& w = ctx->ws;
  FlowJob* jobs = (FlowJob*)w.flow_jobs;
  if ((size_t)B > w.flow_jobs_cap || !jobs) {
    if ((rc = grow_same(ctx, jobs, w.flow_jobs_cap, (size_t)B))) { w.flow_jobs = nullptr; w.flow_jobs_cap = 0; return rc; }
    w.flow_jobs = jobs;
    w.flow_jobs_cap = (size_t)B;
  }
  std::vector<FlowJob> hj((size_t)B);
  int64_t max_n = 0;
  const size_t elem = mode == kFlowGather ? sizeof(int32_t) : sizeof(float);
  for (int b = 0; b < B; ++b) {
    FlowJob& j = hj[b];
    j.flow = flow + (size_t)b * 2 * H * W;
    j.Kinv = Kinv + 9 * (size_t)b;
    j.pts = mode == kFlowCrop ? nullptr : (const char*)pts + (size_t)off[b] * 2 * elem;
    j.out_off = off[b];
    j.n = (int32_t)(off[b + 1] - off[b]);
    j.crop_w = W - 2 * margin;
    j.margin = margin;
    j.pad = 0;
    max_n = std::max<int64_t>(max_n, j.n);
  }
  if (max_n == 0) return TV5_OK;
  if ((rc = submission_enter(ctx, st))) return rc;
  TV5_CUDA(ctx, cudaMemcpyAsync(jobs, hj.data(), sizeof(FlowJob) * B, cudaMemcpyHostToDevice, st));
  flow_points<<<dim3((unsigned)((max_n + 255) / 256), B), 256, 0, st>>>(jobs, H, W, mode, (double2*)x1_out,
                                                                       (double2*)x2_out);
  TV5_CUDA(ctx, cudaGetLastError());
  return submission_leave(ctx, st);
}

int tv5_pose_from_flow(tv5_ctx* ctx, void* stream, const float* flow, int B, int H, int W,
                       const float* Kinv, int mode, int margin, const void* pts,
                       const int64_t* pt_offsets, const int32_t* sets, int iters, double thr,
                       int with_cheirality, float* E32_out, float* P32_out, tv5_result* result,
                       double* E_out, double* P_out) {
  TV5_LOCK(ctx);
  if (!ctx || B < 1 || !result || (!E32_out && !E_out)) return TV5_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int64_t> off;
  int rc = flow_offsets(B, H, W, mode, margin, pt_offsets, off);
  if (rc) return rc;
  for (int b = 0; b < B; ++b)
    if (off[b + 1] == off[b]) return TV5_ERR_INVALID;   // a pair without correspondences
  Workspace& w = ctx->ws;
  if ((rc = grow(ctx, w.flow_x, w.flow_x_cap, (size_t)off[B] * 4))) return rc;
  if ((rc = grow(ctx, w.flow_EP, w.flow_EP_cap, (size_t)B * 21))) return rc;
  double* x1 = w.flow_x;
  double* x2 = w.flow_x + 2 * (size_t)off[B];
  double* E64 = E_out ? E_out : w.flow_EP;
  double* P64 = P_out ? P_out : w.flow_EP + 9 * (size_t)B;
  if ((rc = tv5_flow_to_points(ctx, stream, flow, B, H, W, Kinv, mode, margin, pts, pt_offsets, x1, x2))) return rc;
  if ((rc = tv5_compute_pose_batch(ctx, stream, B, x1, x2, off.data(), sets, iters, 0, 0, thr, with_cheirality,
                                   E64, P64, result, nullptr)))
    return rc;
  if (E32_out || P32_out) {
    pose_to_float<<<(B * 12 + 127) / 128, 128, 0, st>>>(E64, P64, B, E32_out, P32_out);
    TV5_CUDA(ctx, cudaGetLastError());
  }
  return TV5_OK;
}

// ------------------------------------------------------------------------------------------
// plane-sweep cost volume (plane_sweep.cuh)
// ------------------------------------------------------------------------------------------
int tv5_plane_sweep(tv5_ctx* ctx, void* stream, const float* ref_feat, const float* tgt_feat,
                    const float* pose, const float* K, const float* Kinv, int B, int C, int h, int w,
                    int nlabel, float mindepth, int by_depth, float* cost) {
  TV5_LOCK(ctx);
  if (!ctx || !ref_feat || !tgt_feat || !pose || !K || !Kinv || !cost) return TV5_ERR_INVALID;
  if (B < 1 || C < 1 || h < 2 || w < 2 || nlabel < 1 || B > 65535 || nlabel > 65535) return TV5_ERR_INVALID;
  if ((int64_t)h * w > 0x3fffffff) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  SweepParams P;
  P.ref = ref_feat; P.tgt = tgt_feat; P.pose = pose; P.K = K; P.Kinv = Kinv; P.cost = cost;
  P.C = C; P.h = h; P.w = w; P.L = nlabel; P.mindepth = mindepth; P.by_depth = by_depth ? 1 : 0;
  plane_sweep<<<dim3((unsigned)((h * w + 31 + 255) / 256), nlabel, B), 256, 0, (cudaStream_t)stream>>>(P);
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

int tv5_winner_record(tv5_ctx* ctx, void* stream, const double* E, const double* P,
                      const tv5_result* result, int set_offset, void* record_out) {
  TV5_LOCK(ctx);
  if (!ctx || !E || !result || !record_out || set_offset < 0) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  winner_record<<<1, 32, 0, (cudaStream_t)stream>>>(E, P, result, set_offset, (WinnerRecord*)record_out);
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

int tv5_winner_pick(tv5_ctx* ctx, void* stream, const void* records, int G, double* E_out,
                    double* P_out, tv5_result* result_out) {
  TV5_LOCK(ctx);
  if (!ctx || !records || G < 1 || !E_out || !result_out) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  winner_pick<<<1, 32, 0, (cudaStream_t)stream>>>((const WinnerRecord*)records, G, E_out, P_out, result_out);
  TV5_CUDA(ctx, cudaGetLastError());
  return TV5_OK;
}

int tv5_measure_fp32_peak(tv5_ctx* ctx, int mode, double* tflops_out) {
  TV5_LOCK(ctx);
  if (!ctx || !tflops_out || mode < 0 || mode > 1) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  float* sink = nullptr;
  TV5_CUDA(ctx, cudaMalloc(&sink, 4));
  const int iters = 4096, blocks = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, 0);
    if (mode == 0) fp32_peak_kernel<0><<<blocks, 256>>>(sink, iters, 1.0000001f, 1e-9f);
    else fp32_peak_kernel<1><<<blocks, 256>>>(sink, iters, 1.0000001f, 1e-9f);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 128.0 * iters * 256.0 * blocks;  // 128 FMA lanes-ops per thread per iteration
    if (rep > 0 && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  TV5_CUDA(ctx, cudaGetLastError());
  *tflops_out = best;
  return TV5_OK;
}

int tv5_debug_guard(tv5_ctx* ctx, int on, int poison_byte) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  // only on a context that has not allocated anything yet: every buffer is then guarded
  if (ctx->ws.desc || ctx->ws.pp || ctx->ws.E_list || ctx->rng_u || ctx->ws.polish_jobs || ctx->ws.flow_jobs ||
      ctx->ws.h2d_x || ctx->ws.out_E)
    return TV5_ERR_INVALID;
  ctx->guard = on != 0;
  ctx->poison = poison_byte & 0xff;
  return TV5_OK;
}

int tv5_debug_poison(tv5_ctx* ctx, int poison_byte) {
  TV5_LOCK(ctx);
  if (!ctx || !ctx->guard) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  TV5_CUDA(ctx, cudaDeviceSynchronize());
  ctx->poison = poison_byte & 0xff;
  for (auto& g : ctx->guards) {
    if (g.user == (void*)ctx->rng_u) continue;     // the uniform table is written once, by design
    TV5_CUDA(ctx, cudaMemset(g.user, ctx->poison, g.bytes));
  }
  // captured graphs replay the same kernels on the same buffers: nothing to invalidate
  TV5_CUDA(ctx, cudaDeviceSynchronize());
  return TV5_OK;
}

int tv5_debug_check_guards(tv5_ctx* ctx, int64_t* corrupted_bytes_out, int32_t* n_buffers_out) {
  TV5_LOCK(ctx);
  if (!ctx || !corrupted_bytes_out) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  TV5_CUDA(ctx, cudaDeviceSynchronize());
  int64_t bad = 0;
  std::vector<unsigned char> h(kGuardBytes);
  for (auto& g : ctx->guards) {
    const char* zones[2] = {(const char*)g.base, (const char*)g.user + g.payload};
    for (const char* z : zones) {
      TV5_CUDA(ctx, cudaMemcpy(h.data(), z, kGuardBytes, cudaMemcpyDeviceToHost));
      for (unsigned char c : h) bad += c != 0xA5;
    }
    // the slack between the requested size and the 256-byte rounding keeps the zone pattern too
    if (g.payload > g.bytes) {
      std::vector<unsigned char> t(g.payload - g.bytes);
      TV5_CUDA(ctx, cudaMemcpy(t.data(), (const char*)g.user + g.bytes, t.size(), cudaMemcpyDeviceToHost));
      for (unsigned char c : t) bad += c != 0xA5;
    }
  }
  *corrupted_bytes_out = bad;
  if (n_buffers_out) *n_buffers_out = (int32_t)ctx->guards.size();
  return TV5_OK;
}

int tv5_debug_stray_write(tv5_ctx* ctx, int back) {
  TV5_LOCK(ctx);
  if (!ctx || !ctx->guard || ctx->guards.empty()) return TV5_ERR_INVALID;
  TV5_CUDA(ctx, cudaSetDevice(ctx->device));
  const GuardRec& g = ctx->guards.front();
  char* at = back ? (char*)g.user + g.payload : (char*)g.user - 1;   // first byte behind / last byte before the payload
  TV5_CUDA(ctx, cudaMemset(at, 0, 1));
  return TV5_OK;
}

int tv5_set_force_exact(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  ctx->force_exact = on != 0;
  return TV5_OK;
}

int tv5_set_graphs(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  ctx->use_graphs = on != 0;
  return TV5_OK;
}

int tv5_set_early_exit(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  ctx->early_exit = on != 0;
  // development knob: TV5_EARLY_FRAC="0.3,0.52" = interior stage boundaries
  if (const char* e = getenv("TV5_EARLY_FRAC")) {
    int n = 0;
    float f[kEarlyMaxStages + 1] = {0.0f};
    const char* p = e;
    while (*p && n + 1 < kEarlyMaxStages) {
      char* end = nullptr;
      const float v = strtof(p, &end);
      if (end == p) break;
      if (v > f[n] && v < 1.0f) f[++n] = v;
      p = (*end == ',') ? end + 1 : end;
    }
    f[++n] = 1.0f;
    ctx->early_stages = n;
    for (int i = 0; i <= n; ++i) ctx->early_frac[i] = f[i];
  }
  return TV5_OK;
}

int tv5_set_split_solver(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  ctx->split_solver = on != 0;
  return TV5_OK;
}

int tv5_set_overlap(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  ctx->overlap = on != 0;
  return TV5_OK;
}

int tv5_profile_enable(tv5_ctx* ctx, int on) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  profile_collect(ctx);
  ctx->profiling = on != 0;
  return TV5_OK;
}

int tv5_profile_read(tv5_ctx* ctx, double ms_out[TV5_N_STAGES], int64_t launches_out[TV5_N_STAGES],
                     int reset) {
  TV5_LOCK(ctx);
  if (!ctx) return TV5_ERR_INVALID;
  cudaSetDevice(ctx->device);
  profile_collect(ctx);
  for (int i = 0; i < TV5_N_STAGES; ++i) {
    if (ms_out) ms_out[i] = ctx->stage_ms[i];
    if (launches_out) launches_out[i] = ctx->stage_launches[i];
    if (reset) { ctx->stage_ms[i] = 0.0; ctx->stage_launches[i] = 0; }
  }
  return TV5_OK;
}

}  // extern "C"
