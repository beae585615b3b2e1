// polish.cuh — E = U diag(1,1,0) V^T by five Givens rotations and the robust Gauss-Newton
// refinement of E on the five rotation angles (the reference's host-only polish_E.cu:
// Edecomp :147-338, polish_E_robust_parametric :1470-1577, solve_5x5 :340-448, update :450-472).
//
// The reference runs both on one CPU core over CPU tensors.  Here the decomposition is a
// __host__ __device__ function (the arithmetic is IEEE double without contraction, so host and
// device agree bit for bit with the reference), and the refinement is one cooperative kernel:
// G CTAs per problem reduce the 5-vector J^T W e and the 5x5 J^T W J over their points in a fixed
// order, exchange 20 doubles per CTA through global memory across ONE grid barrier per
// iteration, and every CTA applies the same update redundantly — no host round trip between
// iterations, deterministic for a given (N, G).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace tv5 {

// a*b and a+b that the compiler may not contract into an fma (keeps the decomposition bit-equal
// to the reference's host build)
__host__ __device__ __forceinline__ double mul_nc(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
__host__ __device__ __forceinline__ double add_nc(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
__host__ __device__ __forceinline__ double sub_nc(double a, double b) { return add_nc(a, -b); }

struct Rot {  // one plane rotation
  double c, s;
};
__host__ __device__ __forceinline__ Rot rot_annihilating(double keep, double kill) {
  // rotation with c = keep/h, s = -kill/h, h = hypot without scaling (polish_E.cu:160-165)
  double h = sqrt(add_nc(mul_nc(keep, keep), mul_nc(kill, kill)));
  return Rot{keep / h, -kill / h};
}

// Five rotations in the reference's order: z (rows 0,1), y (rows 0,2), x (rows 1,2) from the
// left, u (columns 1,2), v (columns 0,2) from the right.  Only the entries later steps read are
// updated, exactly as the reference does.
struct Givens5 {
  Rot x, y, z, u, v;
};
// `reduced` (optional) receives the working matrix as the reference leaves it in its in/out
// argument — only the entries the later steps read are kept up to date.  The reference's
// `optimise` returns exactly that when it stops before the first update (zero gradient), and
// so does irls_polish.
__host__ __device__ inline Givens5 givens_decompose(const double Ein[9], double* reduced = nullptr) {
  double E[9];
  for (int i = 0; i < 9; ++i) E[i] = Ein[i];
  Givens5 g;
  g.z = rot_annihilating(E[0], E[3]);
  for (int j = 0; j < 3; ++j) {
    double a = E[j], b = E[3 + j];
    E[j] = sub_nc(mul_nc(a, g.z.c), mul_nc(b, g.z.s));
    E[3 + j] = add_nc(mul_nc(a, g.z.s), mul_nc(b, g.z.c));
  }
  g.y = rot_annihilating(E[0], E[6]);
  for (int j = 0; j < 3; ++j) {
    double a = E[j], b = E[6 + j];
    E[j] = sub_nc(mul_nc(a, g.y.c), mul_nc(b, g.y.s));
    E[6 + j] = add_nc(mul_nc(a, g.y.s), mul_nc(b, g.y.c));
  }
  g.x = rot_annihilating(E[4], E[7]);
  for (int j = 1; j < 3; ++j) E[3 + j] = sub_nc(mul_nc(E[3 + j], g.x.c), mul_nc(E[6 + j], g.x.s));
  g.u = rot_annihilating(E[4], E[5]);
  E[2] = add_nc(mul_nc(g.u.s, E[1]), mul_nc(g.u.c, E[2]));
  g.v = rot_annihilating(E[0], E[2]);
  if (reduced)
    for (int i = 0; i < 9; ++i) reduced[i] = E[i];
  return g;
}

__host__ __device__ inline void givens_to_uv(const Givens5& g, double U[9], double V[9]) {
  const double cx = g.x.c, sx = g.x.s, cy = g.y.c, sy = g.y.s, cz = g.z.c, sz = g.z.s;
  const double cu = g.u.c, su = g.u.s, cv = g.v.c, sv = g.v.s;
  U[0] = mul_nc(cy, cz);
  U[1] = add_nc(mul_nc(mul_nc(-cz, sx), sy), mul_nc(cx, sz));
  U[2] = add_nc(mul_nc(mul_nc(cx, cz), sy), mul_nc(sx, sz));
  U[3] = mul_nc(-cy, sz);
  U[4] = add_nc(mul_nc(cx, cz), mul_nc(mul_nc(sx, sy), sz));
  U[5] = sub_nc(mul_nc(cz, sx), mul_nc(mul_nc(cx, sy), sz));
  U[6] = -sy;
  U[7] = mul_nc(-cy, sx);
  U[8] = mul_nc(cx, cy);
  V[0] = cv;
  V[1] = 0.0;
  V[2] = sv;
  V[3] = mul_nc(-su, sv);
  V[4] = cu;
  V[5] = mul_nc(cv, su);
  V[6] = mul_nc(-cu, sv);
  V[7] = -su;
  V[8] = mul_nc(cu, cv);
}

__host__ __device__ inline void givens_to_angles(const Givens5& g, double par[5]) {
  par[0] = atan2(g.x.s, g.x.c);
  par[1] = atan2(g.y.s, g.y.c);
  par[2] = atan2(g.z.s, g.z.c);
  par[3] = atan2(g.u.s, g.u.c);
  par[4] = atan2(g.v.s, g.v.c);
}

// M <- M * R(angle) on columns (c1, c2)   (polish_E.cu:128-145)
__host__ __device__ inline void rotate_columns(double M[9], int c1, int c2, double angle) {
  double c = cos(angle), s = sin(angle);
  for (int i = 0; i < 3; ++i) {
    double a = M[3 * i + c1], b = M[3 * i + c2];
    M[3 * i + c1] = sub_nc(mul_nc(a, c), mul_nc(b, s));
    M[3 * i + c2] = add_nc(mul_nc(a, s), mul_nc(b, c));
  }
}

// 5x5 system by Gaussian elimination with partial pivoting; b <- solution  (polish_E.cu:340-448)
__host__ __device__ inline void solve_sym5(double A[5][5], double b[5]) {
  for (int r = 0; r < 5; ++r) {
    int p = r;
    double big = fabs(A[r][r]);
    for (int i = r + 1; i < 5; ++i)
      if (fabs(A[i][r]) > big) { big = fabs(A[i][r]); p = i; }
    if (p != r) {
      for (int j = r; j < 5; ++j) { double t = A[r][j]; A[r][j] = A[p][j]; A[p][j] = t; }
      double t = b[r]; b[r] = b[p]; b[p] = t;
    }
    for (int i = r + 1; i < 5; ++i) {
      double f = A[i][r] / A[r][r];
      for (int j = r + 1; j < 5; ++j) A[i][j] = sub_nc(A[i][j], mul_nc(f, A[r][j]));
      b[i] = sub_nc(b[i], mul_nc(f, b[r]));
    }
  }
  for (int i = 4; i >= 0; --i) {
    for (int j = i + 1; j < 5; ++j) b[i] = sub_nc(b[i], mul_nc(A[i][j], b[j]));
    b[i] /= A[i][i];
  }
}

// ------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// One thread per matrix.  angles [B,5], U [B,9], V [B,9]; any of them may be null.
__global__ void decompose_batch(const double* __restrict__ E, int B, double* __restrict__ angles,
                                double* __restrict__ U, double* __restrict__ V) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double e[9];
  for (int i = 0; i < 9; ++i) e[i] = E[9 * b + i];
  Givens5 g = givens_decompose(e);
  if (angles) {
    double a[5];
    givens_to_angles(g, a);
    for (int i = 0; i < 5; ++i) angles[5 * b + i] = a[i];
  }
  if (U || V) {
    double u[9], v[9];
    givens_to_uv(g, u, v);
    for (int i = 0; i < 9; ++i) {
      if (U) U[9 * b + i] = u[i];
      if (V) V[9 * b + i] = v[i];
    }
  }
}

struct PolishJob {
  const double* x1;      // [n,2]
  const double* x2;
  const uint8_t* mask;   // [n] or null: points with mask == 0 get weight 0 (extension)
  double* E;             // [9] in: initial estimate, out: refined
  int32_t* iters_out;    // number of parameter updates applied, or null
  int32_t n;
  int32_t pad;
};

constexpr int kPolishThreads = 256;
constexpr int kPolishTerms = 20;  // 5 gradient + 15 upper-triangle normal-matrix entries

// grid = n_jobs * G CTAs (cooperative launch: all co-resident).  partial: [n_jobs][2][G][20]
// doubles, barrier: [n_jobs] counters zeroed before the launch.
__global__ void __launch_bounds__(kPolishThreads) irls_polish(const PolishJob* __restrict__ jobs, int G,
                                                              double delta, double alpha, int max_reps,
                                                              double* __restrict__ partial,
                                                              unsigned int* __restrict__ barrier) {
  const int job = blockIdx.x / G, part = blockIdx.x % G;
  const PolishJob J = jobs[job];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kPolishThreads / 32;
  __shared__ double s_U[9], s_V[9], s_E[9];
  __shared__ double s_warp[kWarps][kPolishTerms];
  __shared__ double s_sum[kPolishTerms];
  __shared__ int s_stop;
  if (tid == 0) {
    double e[9], u[9], v[9];
    for (int i = 0; i < 9; ++i) e[i] = J.E[i];
    double red[9];
    givens_to_uv(givens_decompose(e, red), u, v);
    for (int i = 0; i < 9; ++i) { s_U[i] = u[i]; s_V[i] = v[i]; s_E[i] = red[i]; }
    s_stop = 0;
  }
  __syncthreads();
  int rep = 0;
  for (;; ++rep) {
    double U[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { U[i] = s_U[i]; V[i] = s_V[i]; }
    double acc[kPolishTerms];
#pragma unroll
    for (int i = 0; i < kPolishTerms; ++i) acc[i] = 0.0;
    for (int k = part * kPolishThreads + tid; k < J.n; k += G * kPolishThreads) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(J.x1) + k);
      const double2 b = __ldg(reinterpret_cast<const double2*>(J.x2) + k);
      double p[3], q[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        p[j] = a.x * V[j] + a.y * V[3 + j] + V[6 + j];
        q[j] = b.x * U[j] + b.y * U[3 + j] + U[6 + j];
      }
      const double e = p[0] * q[0] + p[1] * q[1];
      double w = (fabs(e) < delta) ? 1.0 : alpha * delta / fabs(e);
      if (J.mask && !J.mask[k]) w = 0.0;
      double jac[5];
      jac[0] = -p[1] * q[2];
      jac[1] = -p[0] * q[2];
      jac[2] = p[1] * q[0] - p[0] * q[1];
      jac[3] = -p[2] * q[1];
      jac[4] = -p[2] * q[0];
      const double we = -e * w;
      int t = 5;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        acc[i] += jac[i] * we;
        const double wj = w * jac[i];
#pragma unroll
        for (int j = i; j < 5; ++j) acc[t++] += wj * jac[j];
      }
    }
    // CTA reduction in a fixed order: shuffle tree, then warps 0..7 in sequence
#pragma unroll
    for (int i = 0; i < kPolishTerms; ++i) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_warp[warp][i] = v;
    }
    __syncthreads();
    double* mine = partial + ((size_t)(job * 2 + (rep & 1)) * G) * kPolishTerms;
    if (tid < kPolishTerms) {
      double v = 0.0;
      for (int wv = 0; wv < kWarps; ++wv) v += s_warp[wv][tid];
      if (G > 1) __stcg(mine + part * kPolishTerms + tid, v);
      else s_sum[tid] = v;
    }
    if (G > 1) {
      // one barrier per iteration among the G CTAs of this problem (monotonic counter)
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(&barrier[job], 1u);
        const unsigned int target = (unsigned int)(rep + 1) * (unsigned int)G;
        while (*reinterpret_cast<volatile unsigned int*>(&barrier[job]) < target) __nanosleep(32);
        __threadfence();
      }
      __syncthreads();
      if (tid < kPolishTerms) {
        double v = 0.0;
        for (int c = 0; c < G; ++c) v += __ldcg(mine + c * kPolishTerms + tid);
        s_sum[tid] = v;
      }
    }
    __syncthreads();
    if (tid == 0) {
      double g[5], H[5][5];
      int t = 5;
      for (int i = 0; i < 5; ++i) {
        g[i] = s_sum[i];
        for (int j = i; j < 5; ++j) { H[i][j] = s_sum[t]; H[j][i] = s_sum[t]; ++t; }
      }
      double mag = 0.0;
      for (int i = 0; i < 5; ++i) mag += g[i] * g[i];
      if (mag < 1e-20) {
        s_stop = 1;
      } else {
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
            s_E[3 * i + j] = add_nc(mul_nc(s_U[3 * i], s_V[3 * j]), mul_nc(s_U[3 * i + 1], s_V[3 * j + 1]));
        if (rep == max_reps) {
          s_stop = 1;
        } else {
          solve_sym5(H, g);
          double u[9], v[9];
          for (int i = 0; i < 9; ++i) { u[i] = s_U[i]; v[i] = s_V[i]; }
          rotate_columns(u, 0, 1, g[2]);
          rotate_columns(u, 0, 2, g[1]);
          rotate_columns(u, 1, 2, g[0]);
          rotate_columns(v, 1, 2, g[3]);
          rotate_columns(v, 0, 2, g[4]);
          for (int i = 0; i < 9; ++i) { s_U[i] = u[i]; s_V[i] = v[i]; }
        }
      }
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (part == 0 && tid < 9) J.E[tid] = s_E[tid];
  if (part == 0 && tid == 0 && J.iters_out) *J.iters_out = rep;
}
#endif  // __CUDACC__

}  // namespace tv5
