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

// grid (ceil((h*w + 31) / 256), L, B), block 256: thread = (pixel, plane).
// The kernel is bound by L1 wavefronts, not instructions (ncu: l1tex 60 % at 49 % of HBM), so
//  * the pixel index of a warp is shifted per plane so that its 128-byte stores into the volume
//    are 128-byte aligned (one wavefront instead of two; h*w is odd for KITTI's 93 x 307), and
//  * the east taps (x0+1) are taken from the next lane's west taps by shuffle whenever that lane
//    samples the adjacent source pixel (almost always: disparity varies slowly along a row),
//    which halves the gather wavefronts.
__global__ void __launch_bounds__(256) plane_sweep(const SweepParams P) {
  const int hw = P.h * P.w;
  const int i = blockIdx.y, b = blockIdx.z;
  // element index of (b, channel, plane i, pixel 0) is i*hw modulo 32 for every channel (L*hw*c and
  // b*2C*L*hw are multiples of 32 when L is; otherwise the shift is merely not optimal)
  const int shift = (int)(((long long)i * hw) & 31);
  const int p_raw = blockIdx.x * 256 + threadIdx.x - shift;
  const bool live = p_raw >= 0 && p_raw < hw;
  const int p = live ? p_raw : 0;
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
  // taps nw, ne, sw, se = base, base+1, base+w, base+w+1; pointers walk the channel planes so that
  // the loop body carries no 64-bit index arithmetic (it was 55 % of the issued instructions)
  const bool okx0 = x0 >= 0 && x0 < P.w, okx1 = x0 + 1 >= 0 && x0 + 1 < P.w;
  const bool oky0 = y0 >= 0 && y0 < P.h, oky1 = y0 + 1 >= 0 && y0 + 1 < P.h;
  const bool ok0 = okx0 && oky0, ok1 = okx1 && oky0, ok2 = okx0 && oky1, ok3 = okx1 && oky1;
  const bool any = ok0 || ok1 || ok2 || ok3;
  const ptrdiff_t base = any ? (ptrdiff_t)y0 * P.w + x0 : 0;   // only dereferenced under ok*
  // does the next lane's west column coincide with this lane's east column?
  const int lane = threadIdx.x & 31;
  const int nx0 = __shfl_down_sync(0xffffffffu, x0, 1), ny0 = __shfl_down_sync(0xffffffffu, y0, 1);
  const bool nlive = __shfl_down_sync(0xffffffffu, (int)live, 1) != 0;
  const bool nbr = lane < 31 && nlive && nx0 == x0 + 1 && ny0 == y0;
  const float* __restrict__ f0 = P.tgt + (size_t)b * P.C * hw + base;
  const float* __restrict__ f1 = f0 + P.w;
  const float* __restrict__ rf = P.ref + (size_t)b * P.C * hw + p;
  const size_t plane = (size_t)P.L * hw;                       // stride between volume channels
  float* __restrict__ o_r = P.cost + ((size_t)b * 2 * P.C * P.L + i) * hw + p;
  float* __restrict__ o_t = o_r + (size_t)P.C * plane;
#pragma unroll 4
  for (int c = 0; c < P.C; ++c) {
    const float v00 = ok0 ? __ldg(f0) : 0.0f;
    const float v10 = ok2 ? __ldg(f1) : 0.0f;
    const float n00 = __shfl_down_sync(0xffffffffu, v00, 1);
    const float n10 = __shfl_down_sync(0xffffffffu, v10, 1);
    const float v01 = nbr ? n00 : ((ok1) ? __ldg(f0 + 1) : 0.0f);
    const float v11 = nbr ? n10 : ((ok3) ? __ldg(f1 + 1) : 0.0f);
    float acc = 0.0f;
    if (ok0) acc = fmaf(v00, wt[0], acc);
    if (ok1) acc = fmaf(v01, wt[1], acc);
    if (ok2) acc = fmaf(v10, wt[2], acc);
    if (ok3) acc = fmaf(v11, wt[3], acc);
    if (live) {
      __stcs(o_t, acc);
      __stcs(o_r, __ldg(rf));
    }
    f0 += hw; f1 += hw; rf += hw;
    o_t += plane; o_r += plane;
  }
}

}  // namespace tv5
