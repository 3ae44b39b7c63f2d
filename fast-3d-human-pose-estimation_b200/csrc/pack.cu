// Weight packing: fold eval-mode BatchNorm into the preceding conv and re-lay-out for the
// GEMM kernels.  y = (conv(x) + b - mean) * gamma / sqrt(var + eps) + beta
//   =>  W' = W * s,  b' = (b - mean) * s + beta,  s = gamma / sqrt(var + 1e-5)
// (models/cdrnet.py:17-43, models/decoder.py:23-37).  s is formed in fp64.
#include "kernels.h"

namespace cdr {

constexpr double kBnEps = 1e-5;

__device__ __forceinline__ double bn_scale(const CdrConvBn& s, int co) {
  return s.bn_weight ? (double)s.bn_weight[co] / sqrt((double)s.bn_var[co] + kBnEps) : 1.0;
}
__device__ __forceinline__ float folded_bias(const CdrConvBn& s, int co) {
  const double b = s.bias ? (double)s.bias[co] : 0.0;
  if (!s.bn_weight) return (float)b;
  return (float)((b - (double)s.bn_mean[co]) * bn_scale(s, co) + (double)s.bn_bias[co]);
}

// ---- fp32 path: B operand as [K][N] (N contiguous)
__global__ void pack_conv1x1_f32_kernel(CdrConvBn s, int cout, int cin, int k_pad, int n_pad,
                                        float* __restrict__ w_out, float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) bias_out[idx] = idx < cout ? folded_bias(s, (int)idx) : 0.f;
  if (idx >= (long long)k_pad * n_pad) return;
  const int k = (int)(idx / n_pad), n = (int)(idx % n_pad);
  float v = 0.f;
  if (k < cin && n < cout) v = (float)((double)s.weight[(size_t)n * cin + k] * bn_scale(s, n));
  w_out[idx] = v;
}

__global__ void pack_deconv_f32_kernel(CdrConvBn s, int cin, int cout, int n_pad,
                                       float* __restrict__ w_out, float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) bias_out[idx] = idx < cout ? folded_bias(s, (int)idx) : 0.f;
  const long long total = 16LL * cin * n_pad;
  if (idx >= total) return;
  const int n = (int)(idx % n_pad);
  long long r = idx / n_pad;
  const int ci = (int)(r % cin);
  r /= cin;
  const int tap = (int)(r & 3), phase = (int)(r >> 2);
  const int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
  const int ky = 1 - py + 2 * ty, kx = 1 - px + 2 * tx;
  float v = 0.f;
  if (n < cout)
    v = (float)((double)s.weight[(((size_t)ci * cout + n) * 4 + ky) * 4 + kx] * bn_scale(s, n));
  w_out[idx] = v;
}

int launch_pack_conv1x1_f32(const CdrConvBn& src, int cout, int cin, int k_pad, int n_pad,
                            float* w_out, float* bias_out, cudaStream_t st) {
  CDR_CHECK_ARG(src.weight && w_out && bias_out, "pack_conv1x1: null pointer");
  const long long total = (long long)k_pad * n_pad;
  pack_conv1x1_f32_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
      src, cout, cin, k_pad, n_pad, w_out, bias_out);
  CDR_LAUNCH_OK("pack_conv1x1_f32_kernel");
  return CDR_OK;
}

int launch_pack_deconv_f32(const CdrConvBn& src, int cin, int cout, int n_pad, float* w_out,
                           float* bias_out, cudaStream_t st) {
  CDR_CHECK_ARG(src.weight && w_out && bias_out, "pack_deconv: null pointer");
  const long long total = 16LL * cin * n_pad;
  pack_deconv_f32_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
      src, cin, cout, n_pad, w_out, bias_out);
  CDR_LAUNCH_OK("pack_deconv_f32_kernel");
  return CDR_OK;
}

}  // namespace cdr
