// SURVEY §8f rank 3 (first slice): the backward passes of the path's non-conv operators, so that the
// "differentiable" in the reference's DiffDLT-style head (train_cdr.py:105-127 back-propagates through
// process_heatmap and dlt) exists on this library too.  Forward kernels: heatmap.cu / geometry.cu.
//   * soft-argmax  (models/cdrnet.py:120-149,250): kp = scale * sum_i p_i (x_i, y_i), p = softmax(h)
//       dL/dh_i = scale * p_i * (gx (x_i - cx) + gy (y_i - cy))           HBM-bound: 16 KB in, 16 KB out per map
//   * DLT          (models/cdrnet.py:151-179): X = v[:3] / v[3], v = right singular vector of the smallest
//       singular value of A(u_l, v_l, u_r, v_r).  With M = A^T A, dv = -(M - s^2 I)^+ dM v (the gauge torch.svd's
//       backward uses; X does not depend on the sign of v), hence for w = -(M - s^2 I)^+ g_v:
//       dL/dA = (A v) w^T + (A w) v^T, and dA/du_view = e_row (x) P_view[2].     fp64 per joint, Jacobi SVD as forward
//   * FTL          (models/cdrnet.py:45-56) is linear: its backward is cdr_ftl with the transposed matrices.
#include "common.cuh"
#include "jacobi.cuh"

namespace cdr {

constexpr int kSaThreads = 128;

// one CTA per heat-map; the map lives in registers between the passes (hw <= 4096 -> 32 floats per thread)
__global__ void __launch_bounds__(kSaThreads)
softargmax_backward_kernel(const float* __restrict__ heat, const float* __restrict__ grad_kp, int H, int W, float scale,
                           float* __restrict__ grad_heat) {
  __shared__ float s_f[kSaThreads / 32];
  __shared__ double s_d[3][kSaThreads / 32];
  const int hw = H * W, nvec = hw >> 2;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(heat + (size_t)blockIdx.x * hw);
  float4* __restrict__ dst = reinterpret_cast<float4*>(grad_heat + (size_t)blockIdx.x * hw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 v[8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * kSaThreads;
    if (idx < nvec) {
      v[i] = __ldg(src + idx);
      m = fmaxf(fmaxf(m, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_f[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(s_f[0], s_f[1]), fmaxf(s_f[2], s_f[3]));
  double S = 0.0, SX = 0.0, SY = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * kSaThreads;
    if (idx < nvec) {
      const int e0 = idx << 2, row = e0 / W, col = e0 - row * W;      // W % 4 == 0: the four share a row
      v[i].x = __expf(v[i].x - m); v[i].y = __expf(v[i].y - m); v[i].z = __expf(v[i].z - m); v[i].w = __expf(v[i].w - m);
      const float se = (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float sx = fmaf(3.f, v[i].w, fmaf(2.f, v[i].z, v[i].y));
      S += (double)se;
      SX += (double)fmaf((float)col, se, sx);
      SY += (double)((float)row * se);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, o);
    SX += __shfl_xor_sync(0xffffffffu, SX, o);
    SY += __shfl_xor_sync(0xffffffffu, SY, o);
  }
  if (lane == 0) { s_d[0][warp] = S; s_d[1][warp] = SX; s_d[2][warp] = SY; }
  __syncthreads();
  S = (s_d[0][0] + s_d[0][1]) + (s_d[0][2] + s_d[0][3]);
  SX = (s_d[1][0] + s_d[1][1]) + (s_d[1][2] + s_d[1][3]);
  SY = (s_d[2][0] + s_d[2][1]) + (s_d[2][2] + s_d[2][3]);
  const float inv = (float)(1.0 / S), cx = (float)(SX / S), cy = (float)(SY / S);
  const float gx = grad_kp[(size_t)blockIdx.x * 2] * scale * inv, gy = grad_kp[(size_t)blockIdx.x * 2 + 1] * scale * inv;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = threadIdx.x + i * kSaThreads;
    if (idx < nvec) {
      const int e0 = idx << 2, row = e0 / W, col = e0 - row * W;
      const float ty = gy * ((float)row - cy), tx = (float)col - cx;
      float4 g;
      g.x = v[i].x * fmaf(gx, tx, ty);
      g.y = v[i].y * fmaf(gx, tx + 1.f, ty);
      g.z = v[i].z * fmaf(gx, tx + 2.f, ty);
      g.w = v[i].w * fmaf(gx, tx + 3.f, ty);
      dst[idx] = g;
    }
  }
}

__global__ void dlt_backward_kernel(const float* __restrict__ P_l, const float* __restrict__ P_r,
                                    const float* __restrict__ kp_l, const float* __restrict__ kp_r,
                                    const float* __restrict__ grad_xyz, long long total, int joints,
                                    float* __restrict__ grad_kp_l, float* __restrict__ grad_kp_r) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / joints;
  const double gX[3] = {(double)grad_xyz[i * 3], (double)grad_xyz[i * 3 + 1], (double)grad_xyz[i * 3 + 2]};
  double gl[4];
  dlt_backward4(P_l + b * 12, P_r + b * 12, (double)kp_l[i * 2], (double)kp_l[i * 2 + 1], (double)kp_r[i * 2],
                (double)kp_r[i * 2 + 1], gX, gl);     // jacobi.cuh (also exercised on the host: tests/test_host_math.py)
  grad_kp_l[i * 2] = (float)gl[0];
  grad_kp_l[i * 2 + 1] = (float)gl[1];
  grad_kp_r[i * 2] = (float)gl[2];
  grad_kp_r[i * 2 + 1] = (float)gl[3];
}

}  // namespace cdr

using namespace cdr;

extern "C" int cdr_softargmax_backward(const float* heat, const float* grad_kp, long long n_maps, int H, int W,
                                       float scale, float* grad_heat, void* stream) {
  CDR_CHECK_ARG(heat && grad_kp && grad_heat && n_maps >= 0, "cdr_softargmax_backward: bad args");
  CDR_CHECK_ARG(H > 0 && W > 0 && W % 4 == 0 && (long long)H * W <= 4096,
                "cdr_softargmax_backward: heat-map %dx%d (W %% 4 == 0, at most 4096 logits)", H, W);
  CDR_CHECK_ARG((((uintptr_t)heat | (uintptr_t)grad_heat) & 15) == 0, "cdr_softargmax_backward: 16-byte alignment");
  CDR_CHECK_ARG(n_maps <= 0x7fffffffLL, "cdr_softargmax_backward: too many maps for one launch");
  if (n_maps == 0) return CDR_OK;
  softargmax_backward_kernel<<<(unsigned)n_maps, kSaThreads, 0, (cudaStream_t)stream>>>(heat, grad_kp, H, W, scale,
                                                                                        grad_heat);
  CDR_LAUNCH_OK("softargmax_backward_kernel");
  return CDR_OK;
}

extern "C" int cdr_dlt_backward(const float* P_l, const float* P_r, const float* kp_l, const float* kp_r,
                                const float* grad_xyz, int batch, int joints, float* grad_kp_l, float* grad_kp_r,
                                void* stream) {
  CDR_CHECK_ARG(P_l && P_r && kp_l && kp_r && grad_xyz && grad_kp_l && grad_kp_r && batch >= 0 && joints > 0,
                "cdr_dlt_backward: bad args");
  const long long total = (long long)batch * joints;
  if (total == 0) return CDR_OK;
  dlt_backward_kernel<<<(unsigned)ceil_div<long long>(total, 64), 64, 0, (cudaStream_t)stream>>>(
      P_l, P_r, kp_l, kp_r, grad_xyz, total, joints, grad_kp_l, grad_kp_r);
  CDR_LAUNCH_OK("dlt_backward_kernel");
  return CDR_OK;
}
