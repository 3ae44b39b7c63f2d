// SURVEY §8f rank 3, second slice: what a training step of the head needs besides the differentiable operators of
// backward.cu — BatchNorm2d in TRAINING mode (batch statistics, running-stat update, backward) and the reference's three
// losses (models/loss.py:5-98), forward and backward, on the reference's tensor layouts (NCHW fp32 activations,
// (B,J,D) joints).  The convolutions of a training step stay library GEMMs (cuDNN through torch), like the encoder.
//
//   BatchNorm2d(train) — models/cdrnet.py:19,25,28,36,41, models/decoder.py:34 (momentum 0.1, eps 1e-5), optionally
//   fused with the ReLU that follows every BN of the head:
//       mean_c = E[x], var_c = E[(x - mean)^2] over (N, H, W);  y = [relu]((x - mean) * invstd * gamma + beta)
//       running_mean = (1 - m) running_mean + m mean;  running_var = (1 - m) running_var + m var * M / (M - 1)
//       backward (dy' = dy * [y > 0]):  dbeta = sum dy',  dgamma = sum dy' xhat,
//                                       dx = gamma invstd (dy' - dbeta / M - xhat dgamma / M)
//     HBM-bound: forward reads x twice and writes y, backward reads x and dy twice and writes dx.  Sums in fp64.
//
//   Losses on (R = B*J rows, D) tensors with an optional per-row weight w (target_weight[:, j], models/loss.py:23-26):
//       kind 0  JointsMSELoss        :5-32    sum_j 0.5 * mean_{b,d} (w p - w t)^2 / J
//       kind 1  JointsMSESmoothLoss  :35-66   sum_j mean_{b,d} g((w p - w t)^2) / J,  g(v) = v <= thr ? v : v^0.1 thr^0.9
//       kind 2  MPJPELoss            :69-98   sum_j mean_b sqrt(sum_d (w p - w t)^2 + 1e-15) / J
//     Every joint has the same number of terms, so each is one global sum times a constant; fixed-order fp64 tree.
#include "common.cuh"

namespace cdr {

constexpr int kBnThreads = 512;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                   // sh may still be read by the previous call
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];    // every thread: the same fixed order
  return t;
}

// one CTA per channel
__global__ void __launch_bounds__(kBnThreads)
bn_train_stats_kernel(const float* __restrict__ x, int N, int C, int HW, double eps, double momentum,
                      float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                      float* __restrict__ save_invstd) {
  __shared__ double sh[kBnThreads / 32];
  const int c = blockIdx.x;
  double s = 0.0, q = 0.0;
  if ((HW & 3) == 0 && ((uintptr_t)x & 15) == 0) {
    const int hw4 = HW >> 2;
    const long long tot = (long long)N * hw4;
    // four independent 16-byte loads in flight per thread
    for (long long i0 = threadIdx.x; i0 < tot; i0 += 4 * kBnThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * kBnThreads;
        if (i < tot) {
          const long long n = i / hw4;
          v[u] = __ldg(reinterpret_cast<const float4*>(x + ((size_t)n * C + c) * HW) + (int)(i - n * hw4));
        } else {
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += (double)v[u].x + (double)v[u].y + (double)v[u].z + (double)v[u].w;
        q += (double)v[u].x * v[u].x + (double)v[u].y * v[u].y + (double)v[u].z * v[u].z + (double)v[u].w * v[u].w;
      }
    }
  } else {
    for (long long i = threadIdx.x; i < (long long)N * HW; i += kBnThreads) {
      const long long n = i / HW;
      const float v = __ldg(x + ((size_t)n * C + c) * HW + (i - n * HW));
      s += v;
      q += (double)v * v;
    }
  }
  s = block_sum(s, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) {
    const double M = (double)N * HW;
    const double mean = s / M;
    double var = q / M - mean * mean;
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + eps));
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    if (running_var) running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var * (M > 1.0 ? M / (M - 1.0) : 1.0));
  }
}

// elementwise: y = [relu]((x - mean) invstd gamma + beta); idx over float4s of the (N, C, HW) tensor
__global__ void __launch_bounds__(256)
bn_train_apply_kernel(const float* __restrict__ x, long long total4, int C, int hw4, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                      int relu, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)((i / hw4) % C);
  const float a = __ldg(invstd + c) * (gamma ? __ldg(gamma + c) : 1.f);
  const float b = (beta ? __ldg(beta + c) : 0.f) - __ldg(mean + c) * a;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
  float4 o = make_float4(fmaf(v.x, a, b), fmaf(v.y, a, b), fmaf(v.z, a, b), fmaf(v.w, a, b));
  if (relu) {
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  }
  reinterpret_cast<float4*>(y)[i] = o;
}

// one CTA per channel: dbeta = sum dy', dgamma = sum dy' xhat (dy' masked by the ReLU, recomputed from x)
__global__ void __launch_bounds__(kBnThreads)
bn_train_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, int N, int C, int HW,
                          const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                          const float* __restrict__ invstd, int relu, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double sh[kBnThreads / 32];
  const int c = blockIdx.x;
  const float mu = __ldg(mean + c), is = __ldg(invstd + c);
  const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
  double s = 0.0, q = 0.0;
  if ((HW & 3) == 0 && (((uintptr_t)x | (uintptr_t)dy) & 15) == 0) {
    const int hw4 = HW >> 2;
    const long long tot = (long long)N * hw4;
    for (long long i0 = threadIdx.x; i0 < tot; i0 += 2 * kBnThreads) {
      float4 xv[2], dv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long i = i0 + (long long)u * kBnThreads;
        if (i < tot) {
          const long long n = i / hw4;
          const size_t o = ((size_t)n * C + c) * HW;
          xv[u] = __ldg(reinterpret_cast<const float4*>(x + o) + (int)(i - n * hw4));
          dv[u] = __ldg(reinterpret_cast<const float4*>(dy + o) + (int)(i - n * hw4));
        } else {
          xv[u] = make_float4(mu, mu, mu, mu);
          dv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = (xs[e] - mu) * is;
          const float d = (relu && !(fmaf(xh, g, b) > 0.f)) ? 0.f : ds[e];
          s += d;
          q += (double)d * xh;
        }
      }
    }
  } else {
    for (long long i = threadIdx.x; i < (long long)N * HW; i += kBnThreads) {
      const long long n = i / HW;
      const size_t o = ((size_t)n * C + c) * HW + (i - n * HW);
      const float xh = (__ldg(x + o) - mu) * is;
      float d = __ldg(dy + o);
      if (relu && !(fmaf(xh, g, b) > 0.f)) d = 0.f;
      s += d;
      q += (double)d * xh;
    }
  }
  s = block_sum(s, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) {
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
bn_train_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long total, int C, int HW,
                          double inv_m, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                          const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx) {
  // `total` counts groups of VEC elements (VEC = 4 when H*W % 4 == 0 and the tensors are 16-byte aligned, else 1)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)((i / (HW / VEC)) % C);
  const float mu = __ldg(mean + c), is = __ldg(invstd + c);
  const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
  const float k1 = (float)((double)__ldg(dbeta + c) * inv_m), k2 = (float)((double)__ldg(dgamma + c) * inv_m);
  float xs[VEC], ds[VEC], o[VEC];
  if constexpr (VEC == 4) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i), dv = __ldg(reinterpret_cast<const float4*>(dy) + i);
    xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w;
    ds[0] = dv.x; ds[1] = dv.y; ds[2] = dv.z; ds[3] = dv.w;
  } else {
    xs[0] = __ldg(x + i);
    ds[0] = __ldg(dy + i);
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const float xh = (xs[e] - mu) * is;
    const float d = (relu && !(fmaf(xh, g, b) > 0.f)) ? 0.f : ds[e];
    o[e] = g * is * (d - k1 - xh * k2);
  }
  if constexpr (VEC == 4) reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
  else dx[i] = o[0];
}

int launch_bn_train_forward(const float* x, int n, int c, int hw, const float* gamma, const float* beta, double eps,
                            double momentum, float* running_mean, float* running_var, int relu, float* y, float* save_mean,
                            float* save_invstd, cudaStream_t st) {
  if ((hw & 3) != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)y & 15) != 0) {      // before anything is launched: no partial
    set_error("cdr_bn_train_forward: H*W must be a multiple of 4 and the tensors 16-byte aligned");   // running-stat update
    return CDR_ERR_UNSUPPORTED;
  }
  bn_train_stats_kernel<<<c, kBnThreads, 0, st>>>(x, n, c, hw, eps, momentum, running_mean, running_var, save_mean, save_invstd);
  CDR_LAUNCH_OK("bn_train_stats_kernel");
  const long long total4 = (long long)n * c * (hw >> 2);
  bn_train_apply_kernel<<<(unsigned)ceil_div<long long>(total4, 256), 256, 0, st>>>(x, total4, c, hw >> 2, gamma, beta,
                                                                                    save_mean, save_invstd, relu, y);
  CDR_LAUNCH_OK("bn_train_apply_kernel");
  return CDR_OK;
}

int launch_bn_train_backward(const float* x, const float* dy, int n, int c, int hw, const float* gamma, const float* beta,
                             const float* save_mean, const float* save_invstd, int relu, float* dx, float* dgamma,
                             float* dbeta, cudaStream_t st) {
  bn_train_bwd_stats_kernel<<<c, kBnThreads, 0, st>>>(x, dy, n, c, hw, gamma, beta, save_mean, save_invstd, relu, dgamma, dbeta);
  CDR_LAUNCH_OK("bn_train_bwd_stats_kernel");
  const long long total = (long long)n * c * hw;
  if ((hw & 3) == 0 && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0)
    bn_train_bwd_apply_kernel<4><<<(unsigned)ceil_div<long long>(total / 4, 256), 256, 0, st>>>(
        x, dy, total / 4, c, hw, 1.0 / ((double)n * hw), gamma, beta, save_mean, save_invstd, relu, dgamma, dbeta, dx);
  else
    bn_train_bwd_apply_kernel<1><<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
        x, dy, total, c, hw, 1.0 / ((double)n * hw), gamma, beta, save_mean, save_invstd, relu, dgamma, dbeta, dx);
  CDR_LAUNCH_OK("bn_train_bwd_apply_kernel");
  return CDR_OK;
}

// ------------------------------------------------------------------------------------------ losses
constexpr int kLossBlocks = 256, kLossThreads = 256;

__device__ __forceinline__ float loss_diff(const float* __restrict__ p, const float* __restrict__ t, float w, bool use_w,
                                           long long i) {
  const float a = __ldg(p + i), b = __ldg(t + i);
  return use_w ? __fsub_rn(__fmul_rn(a, w), __fmul_rn(b, w)) : __fsub_rn(a, b);     // the reference's operation order
}
__device__ __forceinline__ double smooth_term(float v, float thr) {
  return v > thr ? (double)(powf(v, 0.1f) * powf(thr, 0.9f)) : (double)v;
}

// partial[b] = this block's share of sum_rows term(row) (kind 2) / sum_elements term (kinds 0, 1): one warp per row
__global__ void __launch_bounds__(kLossThreads)
joint_loss_partial_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                          const float* __restrict__ weight, long long rows, int D, float thr, double* __restrict__ partial) {
  __shared__ double sh[kLossThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  for (long long r = (long long)blockIdx.x * (kLossThreads / 32) + warp; r < rows; r += (long long)gridDim.x * (kLossThreads / 32)) {
    const float w = weight ? __ldg(weight + r) : 1.f;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
      const float df = loss_diff(pred, target, w, weight != nullptr, r * D + d);
      const float v = __fmul_rn(df, df);
      s += kind == 1 ? smooth_term(v, thr) : (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    acc += kind == 2 ? (double)sqrtf((float)s + 1e-15f) : s;
  }
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kLossThreads / 32; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void joint_loss_final_kernel(const double* __restrict__ partial, int n, double factor, float* __restrict__ loss) {
  double t = 0.0;
  for (int i = 0; i < n; ++i) t += partial[i];
  *loss = (float)(t * factor);
}
// grad_pred[r, d] = g * factor * d term / d pred
__global__ void __launch_bounds__(kLossThreads)
joint_loss_backward_kernel(int kind, const float* __restrict__ pred, const float* __restrict__ target,
                           const float* __restrict__ weight, long long rows, int D, float thr, double factor,
                           const float* __restrict__ grad_loss, float* __restrict__ grad_pred) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double g = (double)__ldg(grad_loss) * factor;
  for (long long r = (long long)blockIdx.x * (kLossThreads / 32) + warp; r < rows; r += (long long)gridDim.x * (kLossThreads / 32)) {
    const float w = weight ? __ldg(weight + r) : 1.f;
    double s = 0.0;
    if (kind == 2) {
      for (int d = lane; d < D; d += 32) {
        const float df = loss_diff(pred, target, w, weight != nullptr, r * D + d);
        s += (double)__fmul_rn(df, df);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    const double inv_norm = kind == 2 ? 1.0 / (double)sqrtf((float)s + 1e-15f) : 0.0;
    for (int d = lane; d < D; d += 32) {
      const long long i = r * D + d;
      const float df = loss_diff(pred, target, w, weight != nullptr, i);
      double t;
      if (kind == 2) {
        t = (double)df * inv_norm;                                    // d sqrt(sum df^2) / d df
      } else if (kind == 1) {
        const float v = __fmul_rn(df, df);
        // d g(v) / d v * 2 df;  g'(v) = 0.1 v^-0.9 thr^0.9 above the threshold
        t = v > thr ? 0.2 * (double)df * (double)(powf(v, -0.9f) * powf(thr, 0.9f)) : 2.0 * (double)df;
      } else {
        t = 2.0 * (double)df;
      }
      grad_pred[i] = (float)(g * t * (double)w);
    }
  }
}

static double loss_factor(int kind, long long rows, int D) {
  // per joint: mean over B*D elements (B rows per joint) — with J joints and rows = B*J: 1 / (rows * D) overall
  if (kind == 0) return 0.5 / ((double)rows * D);
  if (kind == 1) return 1.0 / ((double)rows * D);
  return 1.0 / (double)rows;
}

int launch_joint_loss_forward(int kind, const float* pred, const float* target, const float* weight, long long rows, int D,
                              double threshold, float* loss, double* scratch, cudaStream_t st) {
  const int blocks = (int)(rows < kLossBlocks * (kLossThreads / 32) ? ceil_div<long long>(rows, kLossThreads / 32) : kLossBlocks);
  joint_loss_partial_kernel<<<blocks, kLossThreads, 0, st>>>(kind, pred, target, weight, rows, D, (float)threshold, scratch);
  CDR_LAUNCH_OK("joint_loss_partial_kernel");
  joint_loss_final_kernel<<<1, 1, 0, st>>>(scratch, blocks, loss_factor(kind, rows, D), loss);
  CDR_LAUNCH_OK("joint_loss_final_kernel");
  return CDR_OK;
}
int launch_joint_loss_backward(int kind, const float* pred, const float* target, const float* weight, long long rows, int D,
                               double threshold, const float* grad_loss, float* grad_pred, cudaStream_t st) {
  const int blocks = (int)(rows < 4 * kLossBlocks * (kLossThreads / 32) ? ceil_div<long long>(rows, kLossThreads / 32) : 4 * kLossBlocks);
  joint_loss_backward_kernel<<<blocks, kLossThreads, 0, st>>>(kind, pred, target, weight, rows, D, (float)threshold,
                                                              loss_factor(kind, rows, D), grad_loss, grad_pred);
  CDR_LAUNCH_OK("joint_loss_backward_kernel");
  return CDR_OK;
}
size_t joint_loss_scratch_bytes() { return (size_t)kLossBlocks * sizeof(double); }

}  // namespace cdr
