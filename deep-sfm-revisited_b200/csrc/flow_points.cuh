// flow_points.cuh — optical flow -> K^-1-normalised correspondences, one kernel.
//
// Replaces the chain the reference runs per image before every pose solve
// (models/SFMnet.py): flow2coord (:298-318: pixel grid, grid + flow, homogeneous 1), the point
// selection of pose_by_ransac (:239-254: dense crop `margin:-margin`, integer keypoint gather,
// or bilinear grid_sample at sub-pixel keypoints), bmm(K^-1, .) (:259-260), transpose / [:, :2]
// / contiguous (:262-263) and the .double() of epipolar_utils.py:130 — six N-sized temporaries
// and a dtype round trip — by a single pass that reads the two flow planes once and writes the
// engine's float64 [N,2] arrays.  HBM-bound: 8 B read + 32 B written per correspondence.
//
// Arithmetic is float32 exactly where the reference's is (grid + flow, the sampling weights,
// the 3-term products with K^-1), widened to float64 at the end, so the values are the ones
// the reference hands to `essential_matrix.computeP`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tv5 {

enum FlowMode : int32_t {
  kFlowCrop = 0,      // every pixel of [margin, H-margin) x [margin, W-margin), row-major
  kFlowGather = 1,    // integer pixel list (x, y) int32
  kFlowBilinear = 2,  // sub-pixel list (x, y) float32, bilinear, align_corners=True, zero padding
};

struct FlowJob {
  const float* flow;   // [2,H,W] of this image
  const float* Kinv;   // [3,3] row-major
  const void* pts;     // mode 1: int32 [n,2]; mode 2: float [n,2]; mode 0: unused
  int64_t out_off;     // first output point of this image
  int32_t n;           // points of this image
  int32_t crop_w;      // mode 0: W - 2*margin
  int32_t margin;
  int32_t pad;
};

__device__ __forceinline__ float2 apply_kinv(const float* __restrict__ K, float c0, float c1, float c2) {
  // rows 0 and 1 of K^-1 (c0, c1, c2)^T, accumulated k = 0, 1, 2 like an SGEMM inner loop
  float2 r;
  r.x = fmaf(__ldg(K + 2), c2, fmaf(__ldg(K + 1), c1, __ldg(K + 0) * c0));
  r.y = fmaf(__ldg(K + 5), c2, fmaf(__ldg(K + 4), c1, __ldg(K + 3) * c0));
  return r;
}

// grid (ceil(max_n / 256), B)
__global__ void __launch_bounds__(256) flow_points(const FlowJob* __restrict__ jobs, int H, int W, int mode,
                                                   double2* __restrict__ x1_out, double2* __restrict__ x2_out) {
  const FlowJob J = jobs[blockIdx.y];
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= J.n) return;
  const float* __restrict__ fx = J.flow;
  const float* __restrict__ fy = J.flow + (size_t)H * W;
  float a0, a1, a2, b0, b1, b2;  // homogeneous pixel coordinates in image 1 / image 2
  if (mode == kFlowBilinear) {
    const float2 p = __ldg(reinterpret_cast<const float2*>(J.pts) + k);
    // the reference normalises to [-1,1] with separate torch ops on the GPU (SFMnet.py:247) and
    // grid_sample maps back; torch's CUDA division by a Python scalar multiplies by the
    // reciprocal, and nothing is fused across the three ops
    const float gx = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, p.x), __frcp_rn((float)max(W - 1, 1))), 1.0f);
    const float gy = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, p.y), __frcp_rn((float)max(H - 1, 1))), 1.0f);
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
    const float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
    const float w[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};  // nw, ne, sw, se
    const int xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
    a0 = a1 = a2 = b0 = b1 = b2 = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (xs[c] < 0 || xs[c] >= W || ys[c] < 0 || ys[c] >= H) continue;
      const size_t o = (size_t)ys[c] * W + xs[c];
      const float px = (float)xs[c], py = (float)ys[c];
      a0 += px * w[c];
      a1 += py * w[c];
      a2 += w[c];
      b0 += (px + __ldg(fx + o)) * w[c];
      b1 += (py + __ldg(fy + o)) * w[c];
    }
    b2 = a2;
  } else {
    int x, y;
    if (mode == kFlowCrop) {
      y = k / J.crop_w;
      x = k - y * J.crop_w + J.margin;
      y += J.margin;
    } else {
      const int2 p = __ldg(reinterpret_cast<const int2*>(J.pts) + k);
      // negative indices wrap like the reference's advanced indexing; anything else is clamped
      x = p.x < 0 ? p.x + W : p.x;
      y = p.y < 0 ? p.y + H : p.y;
      x = min(max(x, 0), W - 1);
      y = min(max(y, 0), H - 1);
    }
    const size_t o = (size_t)y * W + x;
    a0 = (float)x;
    a1 = (float)y;
    b0 = a0 + __ldg(fx + o);
    b1 = a1 + __ldg(fy + o);
    a2 = b2 = 1.0f;
  }
  const float2 n1 = apply_kinv(J.Kinv, a0, a1, a2);
  const float2 n2 = apply_kinv(J.Kinv, b0, b1, b2);
  x1_out[J.out_off + k] = make_double2((double)n1.x, (double)n1.y);
  x2_out[J.out_off + k] = make_double2((double)n2.x, (double)n2.y);
}

// float64 E [B,9] / P [B,12] -> the float32 B x 3 x 3 / B x 3 x 4 tensors SFMnet stores
// (SFMnet.py:186-187,272)
__global__ void pose_to_float(const double* __restrict__ E, const double* __restrict__ P, int B,
                              float* __restrict__ E32, float* __restrict__ P32) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 9 && E32) E32[i] = (float)E[i];
  if (i < B * 12 && P32) P32[i] = (float)P[i];
}

}  // namespace tv5
