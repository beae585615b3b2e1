// plane_sweep.cuh — plane-sweep cost volume from the estimated pose, one kernel.
//
// Replaces the label loop of PSNet.forward (models/PSNet.py:141-157): for each of the nlabel
// depth planes the reference calls inverse_warp (models/inverse_warp.py:121-153 — pixel2cam
// :31-45, K [R|t] projection and [-1,1] normalisation :48-78, bilinear grid_sample with zero
// padding and align_corners=True) on the 32-channel quarter-resolution target features and
// copies the result and the reference features into a [B, 2C, nlabel, h, w] volume: per plane
// about ten elementwise/bmm launches with full-size temporaries plus two strided copies, 128 times.
// Here one launch computes each (pixel, plane) sample position once, gathers the C channels and
// writes both halves of the volume exactly once.  HBM-bound on the OUTPUT: 2C*4 bytes written
// per (pixel, plane); the feature maps (a few MB) stay in L2.
//
// Arithmetic follows the reference's float32 sequence (torch CUDA semantics: a tensor divided by
// a Python scalar is multiplied by the reciprocal; 3-term products accumulate k = 0,1,2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tv5 {

struct SweepParams {
  const float* ref;    // [B,C,h,w]
  const float* tgt;    // [B,C,h,w]
  const float* pose;   // [B,3,4]  [R|t], camera 1 -> camera 2 (what tv5_compute_pose returns, as float32)
  const float* K;      // [B,3,3]  intrinsics at feature resolution
  const float* Kinv;   // [B,3,3]
  float* cost;         // [B,2C,L,h,w]
  int32_t C, h, w, L;
  float mindepth;
  int32_t by_depth;    // cfg.PREDICT_BY_DEPTH: depth_i = (i+1)*mindepth, else mindepth*L/(i+1)
};

__device__ __forceinline__ float dot3_sgemm(float m0, float m1, float m2, float v0, float v1, float v2) {
  return fmaf(m2, v2, fmaf(m1, v1, __fmul_rn(m0, v0)));
}

// grid (ceil(h*w / 256), L, B), block 256: thread = (pixel, plane)
__global__ void __launch_bounds__(256) plane_sweep(const SweepParams P) {
  const int hw = P.h * P.w;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y, b = blockIdx.z;
  __shared__ float s_proj[12], s_kinv[9];
  if (threadIdx.x < 12) {
    const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
    const float* K = P.K + 9 * b;
    const float* T = P.pose + 12 * b;
    s_proj[threadIdx.x] = dot3_sgemm(K[3 * r], K[3 * r + 1], K[3 * r + 2], T[c], T[4 + c], T[8 + c]);
  } else if (threadIdx.x >= 32 && threadIdx.x < 41) {
    s_kinv[threadIdx.x - 32] = P.Kinv[9 * b + threadIdx.x - 32];
  }
  __syncthreads();
  if (p >= hw) return;
  const int y = p / P.w, x = p - y * P.w;
  // depth of plane i  (PSNet.py:142,150-153)
  float depth;
  if (P.by_depth) {
    depth = __fmul_rn(__fmul_rn(1.0f, (float)(i + 1)), P.mindepth);
  } else {
    const float d2d = __fmul_rn(__fmul_rn(1.0f, P.mindepth), (float)P.L);
    depth = __fmul_rn(d2d, __frcp_rn((float)((double)(i + 1) + 1e-16)));
  }
  // pixel2cam: (Kinv pix) * depth
  const float fx = (float)x, fy = (float)y;
  float cam[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    cam[r] = __fmul_rn(dot3_sgemm(s_kinv[3 * r], s_kinv[3 * r + 1], s_kinv[3 * r + 2], fx, fy, 1.0f), depth);
  // cam2pixel
  float pc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    pc[r] = __fadd_rn(dot3_sgemm(s_proj[4 * r], s_proj[4 * r + 1], s_proj[4 * r + 2], cam[0], cam[1], cam[2]),
                      s_proj[4 * r + 3]);
  const float Z = fmaxf(pc[2], 1e-3f);
  float xn = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fdiv_rn(pc[0], Z)), __frcp_rn((float)(P.w - 1))), 1.0f);
  float yn = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, __fdiv_rn(pc[1], Z)), __frcp_rn((float)(P.h - 1))), 1.0f);
  if (xn > 1.0f || xn < -1.0f) xn = 2.0f;
  if (yn > 1.0f || yn < -1.0f) yn = 2.0f;
  // grid_sample, bilinear, zeros, align_corners=True
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), 0.5f), (float)(P.w - 1));
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), 0.5f), (float)(P.h - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = __fsub_rn(ix, x0f), wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), ix);
  const float wy1 = __fsub_rn(iy, y0f), wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), iy);
  // NaN coordinates (NaN pose) give NaN weights and, like torch, no in-bounds tap
  const bool finite = (ix == ix) && (iy == iy) && fabsf(ix) < 1.0e9f && fabsf(iy) < 1.0e9f;
  const int x0 = finite ? (int)x0f : -10, y0 = finite ? (int)y0f : -10;
  const float wt[4] = {__fmul_rn(wx0, wy0), __fmul_rn(wx1, wy0), __fmul_rn(wx0, wy1), __fmul_rn(wx1, wy1)};
  int off[4];
  bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xs = x0 + (k & 1), ys = y0 + (k >> 1);
    ok[k] = xs >= 0 && xs < P.w && ys >= 0 && ys < P.h;
    off[k] = ok[k] ? ys * P.w + xs : 0;
  }
  const float* __restrict__ tg = P.tgt + (size_t)b * P.C * hw;
  const float* __restrict__ rf = P.ref + (size_t)b * P.C * hw + p;
  const size_t plane = (size_t)P.L * hw;                       // stride between volume channels
  float* __restrict__ out_ref = P.cost + ((size_t)b * 2 * P.C * P.L + i) * hw + p;
  float* __restrict__ out_tgt = out_ref + (size_t)P.C * plane;
#pragma unroll 4
  for (int c = 0; c < P.C; ++c) {
    const float* __restrict__ f = tg + (size_t)c * hw;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) acc = fmaf(__ldg(f + off[k]), wt[k], acc);
    __stcs(out_tgt + (size_t)c * plane, acc);
    __stcs(out_ref + (size_t)c * plane, __ldg(rf + (size_t)c * hw));
  }
}

}  // namespace tv5
