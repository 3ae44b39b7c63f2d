// Small-matrix geometry kernels: pinv, DLT, baseline triangulation, MPJPE sums.
// All are HBM/latency-bound (tens of bytes per item); fp64 internals.
#include "common.cuh"
#include "jacobi.cuh"

namespace cdr {

// ---------------------------------------------------------------- pinv
__global__ void pinv_kernel(const float* __restrict__ P, int n, double rtol,
                            float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pinv_3x4<float, float>(P + (size_t)i * 12, rtol, out + (size_t)i * 12);
}

// l1max (optional, pre-zeroed): [0] <- max row L1 norm of the pseudo-inverses, [1] <- of the matrices themselves
// (inflated by 2^-20: upper bounds) — the output-scale bounds of the fp16-plane FTL (layout.cu: ftl_f16p_vec_kernel)
__global__ void pinv2_kernel(const float* __restrict__ P_a, const float* __restrict__ P_b, int n, double rtol,
                             float* __restrict__ out, float* __restrict__ l1max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float la = 0.f, lb = 0.f;
  if (i < 2 * n) {
    const float* P = i < n ? P_a + (size_t)i * 12 : P_b + (size_t)(i - n) * 12;
    float* o = out + (size_t)i * 12;
    pinv_3x4<float, float>(P, rtol, o);
    if (l1max) {
      for (int r = 0; r < 4; ++r) la = fmaxf(la, fabsf(o[r * 3]) + fabsf(o[r * 3 + 1]) + fabsf(o[r * 3 + 2]));
      for (int r = 0; r < 3; ++r)
        lb = fmaxf(lb, fabsf(P[r * 4]) + fabsf(P[r * 4 + 1]) + fabsf(P[r * 4 + 2]) + fabsf(P[r * 4 + 3]));
    }
  }
  if (l1max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      la = fmaxf(la, __shfl_xor_sync(0xffffffffu, la, o));
      lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o));
    }
    if ((threadIdx.x & 31) == 0) {
      const float infl = 1.f + 9.5367431640625e-07f;
      atomicMax(reinterpret_cast<unsigned int*>(l1max), __float_as_uint(la * infl));       // non-negative floats
      atomicMax(reinterpret_cast<unsigned int*>(l1max + 1), __float_as_uint(lb * infl));
    }
  }
}
int launch_pinv2(const float* P_a, const float* P_b, int n, double rtol, float* out, float* l1max, cudaStream_t st) {
  CDR_CHECK_ARG(P_a && P_b && out && n > 0, "pinv2: bad args");
  pinv2_kernel<<<ceil_div(2 * n, 32), 32, 0, st>>>(P_a, P_b, n, rtol, out, l1max);   // one warp per block: spread over SMs
  CDR_LAUNCH_OK("pinv_kernel");
  return CDR_OK;
}

// ---------------------------------------------------------------- DLT (models/cdrnet.py:151-179)
__global__ void dlt_kernel(const float* __restrict__ P_l, const float* __restrict__ P_r,
                           const float* __restrict__ kp_l, const float* __restrict__ kp_r,
                           long long total, int joints, float* __restrict__ xyz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / joints;
  double A[4][4];
  dlt_rows(P_l + b * 12, (double)kp_l[i * 2], (double)kp_l[i * 2 + 1], A, 0);
  dlt_rows(P_r + b * 12, (double)kp_r[i * 2], (double)kp_r[i * 2 + 1], A, 2);
  double x, y, z;
  dlt_solve4(A, x, y, z);
  xyz[i * 3 + 0] = (float)x;
  xyz[i * 3 + 1] = (float)y;
  xyz[i * 3 + 2] = (float)z;
}

// ---------------------------------------------------------------- baseline triangulation
// tools/common.py:61-68: rows [v*P[2]-P[1] ; P[0]-u*P[2]] per view.  Row order and sign do
// not change the null space; eig(M^T M) arg-min == smallest right singular vector of M.
__global__ void triangulate_u8_kernel(const double* __restrict__ P1, const double* __restrict__ P2,
                                      int p_stride, const uint8_t* __restrict__ pts1,
                                      const uint8_t* __restrict__ pts2, long long total,
                                      int joints, double* __restrict__ xyz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / joints;
  const double* p1 = P1 + b * p_stride;
  const double* p2 = P2 + b * p_stride;
  const double u1 = pts1[i * 2], v1 = pts1[i * 2 + 1];
  const double u2 = pts2[i * 2], v2 = pts2[i * 2 + 1];
  double A[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    A[0][c] = v1 * p1[8 + c] - p1[4 + c];
    A[1][c] = p1[c] - u1 * p1[8 + c];
    A[2][c] = v2 * p2[8 + c] - p2[4 + c];
    A[3][c] = p2[c] - u2 * p2[8 + c];
  }
  double x, y, z;
  dlt_solve4(A, x, y, z);
  xyz[i * 3 + 0] = x;
  xyz[i * 3 + 1] = y;
  xyz[i * 3 + 2] = z;
}

// ---------------------------------------------------------------- MPJPE (models/metrics.py:82-95)
template <typename TP>
__global__ void mpjpe_pose_kernel(const TP* __restrict__ p2l, const TP* __restrict__ p2r,
                                  const TP* __restrict__ p3, const double* __restrict__ g3,
                                  const double* __restrict__ g2l, const double* __restrict__ g2r,
                                  const double* __restrict__ w, int w_batched, int w_f32,
                                  long long n, int joints, double* __restrict__ pose_err) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int j = 0; j < joints; ++j) {
    const long long i = b * joints + j;
    const double wt = w ? (w_batched ? w[i] : w[j]) : 1.0;
    auto wp = [&](TP v) -> double {
      // pred * weight: fp32 product when both operands are fp32 in the reference's numpy
      // call (train_cdr.py:196-199), fp64 otherwise
      if (w_f32 && sizeof(TP) == 4) return (double)__fmul_rn((float)v, (float)wt);
      return (double)v * wt;
    };
    double dx = wp(p2l[i * 2]) - g2l[i * 2] * wt, dy = wp(p2l[i * 2 + 1]) - g2l[i * 2 + 1] * wt;
    s0 += sqrt(dx * dx + dy * dy);
    dx = wp(p2r[i * 2]) - g2r[i * 2] * wt;
    dy = wp(p2r[i * 2 + 1]) - g2r[i * 2 + 1] * wt;
    s1 += sqrt(dx * dx + dy * dy);
    dx = wp(p3[i * 3]) - g3[i * 3] * wt;
    dy = wp(p3[i * 3 + 1]) - g3[i * 3 + 1] * wt;
    const double dz = wp(p3[i * 3 + 2]) - g3[i * 3 + 2] * wt;
    s2 += sqrt(dx * dx + dy * dy + dz * dz);
  }
  pose_err[b * 3 + 0] = s0;
  pose_err[b * 3 + 1] = s1;
  pose_err[b * 3 + 2] = s2;
}

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;

// fixed-order tree: block `b` owns rows [b*chunk, (b+1)*chunk)
__global__ void __launch_bounds__(kRedThreads)
reduce3_kernel(const double* __restrict__ in, long long n, long long chunk,
               double* __restrict__ out) {
  __shared__ double sm[3][kRedThreads];
  const long long lo = (long long)blockIdx.x * chunk;
  const long long hi = (lo + chunk < n) ? lo + chunk : n;
  double a0 = 0, a1 = 0, a2 = 0;
  for (long long i = lo + threadIdx.x; i < hi; i += kRedThreads) {
    a0 += in[i * 3];
    a1 += in[i * 3 + 1];
    a2 += in[i * 3 + 2];
  }
  sm[0][threadIdx.x] = a0;
  sm[1][threadIdx.x] = a1;
  sm[2][threadIdx.x] = a2;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sm[0][threadIdx.x] += sm[0][threadIdx.x + s];
      sm[1][threadIdx.x] += sm[1][threadIdx.x + s];
      sm[2][threadIdx.x] += sm[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[blockIdx.x * 3 + 0] = sm[0][0];
    out[blockIdx.x * 3 + 1] = sm[1][0];
    out[blockIdx.x * 3 + 2] = sm[2][0];
  }
}

__global__ void finish_sums_kernel(const double* __restrict__ part, double count,
                                   double* __restrict__ sums) {
  sums[0] = part[0];
  sums[1] = part[1];
  sums[2] = part[2];
  sums[3] = count;
}

static int reduce_pose_err(const double* pose_err, long long n, int joints, double* sums,
                           double* partials, cudaStream_t st) {
  // level 1: <=1024 partials; level 2: one block
  long long blocks = ceil_div<long long>(n, 4096);
  if (blocks > kRedMaxBlocks) blocks = kRedMaxBlocks;
  if (blocks < 1) blocks = 1;
  const long long chunk = ceil_div<long long>(n, blocks);
  reduce3_kernel<<<(unsigned)blocks, kRedThreads, 0, st>>>(pose_err, n, chunk, partials);
  CDR_LAUNCH_OK("reduce3_kernel");
  double* final_part = partials + (size_t)kRedMaxBlocks * 3;
  reduce3_kernel<<<1, kRedThreads, 0, st>>>(partials, blocks, blocks, final_part);
  CDR_LAUNCH_OK("reduce3_kernel");
  finish_sums_kernel<<<1, 1, 0, st>>>(final_part, (double)n * (double)joints, sums);
  CDR_LAUNCH_OK("finish_sums_kernel");
  return CDR_OK;
}

// P = T * K * [R | t] (fp64, numpy's left-to-right order), first three rows as fp32 — what the model consumes:
// tools/common.py:28-32 (get_projection_matrix: K @ hstack(R, T)), dataset/mads_3d.py:223-226 / tools/load.py:60-67
// (T = eye(4) with the 2x3 crop/resize affine in its top-left: `T @ P`, resp. `trans @ K`), inference.py:53-56
// (`numpy2torch(P[:3])`: float32).  One thread per camera.
__global__ void projection_kernel(const double* __restrict__ K, int k_stride, const double* __restrict__ R,
                                  const double* __restrict__ t, const double* __restrict__ trans, int n,
                                  float* __restrict__ P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* k = K + (size_t)i * k_stride;
  const double* r = R + (size_t)i * 9;
  const double* tt = t + (size_t)i * 3;
  double Rt[3][4], KP[3][4];
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) Rt[a][b] = r[a * 3 + b];
    Rt[a][3] = tt[a];
  }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 4; ++b) {
      double s = __dmul_rn(k[a * 3], Rt[0][b]);       // no FMA contraction: numpy rounds every product
      s = __dadd_rn(s, __dmul_rn(k[a * 3 + 1], Rt[1][b]));
      s = __dadd_rn(s, __dmul_rn(k[a * 3 + 2], Rt[2][b]));
      KP[a][b] = s;
    }
  float* out = P + (size_t)i * 12;
  if (trans) {
    // T @ KP with T = [[a00 a01 a02 0] [a10 a11 a12 0] [0 0 1 0] [0 0 0 1]] and KP's 4th row = (0,0,0,1)
    const double* m = trans + (size_t)i * 6;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 4; ++b) {
        double s = __dmul_rn(m[a * 3], KP[0][b]);
        s = __dadd_rn(s, __dmul_rn(m[a * 3 + 1], KP[1][b]));
        s = __dadd_rn(s, __dmul_rn(m[a * 3 + 2], KP[2][b]));
        out[a * 4 + b] = (float)s;
      }
    for (int b = 0; b < 4; ++b) out[8 + b] = (float)KP[2][b];
  } else {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 4; ++b) out[a * 4 + b] = (float)KP[a][b];
  }
}

}  // namespace cdr

using namespace cdr;

extern "C" int cdr_projection_matrices(const double* K, int k_batched, const double* R, const double* t,
                                       const double* trans, int n, float* P, void* stream) {
  CDR_CHECK_ARG(K && R && t && P && n >= 0, "cdr_projection_matrices: null pointer or negative n");
  if (n == 0) return CDR_OK;
  projection_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(K, k_batched ? 9 : 0, R, t, trans, n, P);
  CDR_LAUNCH_OK("projection_kernel");
  return CDR_OK;
}

extern "C" int cdr_pinv(const float* P, int n, double rtol, float* pinv, void* stream) {
  CDR_CHECK_ARG(P && pinv && n >= 0, "cdr_pinv: null pointer or negative n");
  if (n == 0) return CDR_OK;
  pinv_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(P, n, rtol, pinv);
  CDR_LAUNCH_OK("pinv_kernel");
  return CDR_OK;
}

extern "C" int cdr_dlt(const float* P_l, const float* P_r, const float* kp_l, const float* kp_r,
                       int batch, int joints, float* xyz, void* stream) {
  CDR_CHECK_ARG(P_l && P_r && kp_l && kp_r && xyz, "cdr_dlt: null pointer");
  CDR_CHECK_ARG(batch >= 0 && joints > 0, "cdr_dlt: bad batch/joints");
  const long long total = (long long)batch * joints;
  if (total == 0) return CDR_OK;
  dlt_kernel<<<(unsigned)ceil_div<long long>(total, 64), 64, 0, (cudaStream_t)stream>>>(
      P_l, P_r, kp_l, kp_r, total, joints, xyz);
  CDR_LAUNCH_OK("dlt_kernel");
  return CDR_OK;
}

extern "C" int cdr_triangulate_u8(const double* P1, const double* P2, int p_rows, int p_batched,
                                  const uint8_t* pts1, const uint8_t* pts2, long long n_poses,
                                  int joints, double* xyz, void* stream) {
  CDR_CHECK_ARG(P1 && P2 && pts1 && pts2 && xyz, "cdr_triangulate_u8: null pointer");
  CDR_CHECK_ARG(p_rows == 3 || p_rows == 4, "cdr_triangulate_u8: p_rows must be 3 or 4");
  CDR_CHECK_ARG(n_poses >= 0 && joints > 0, "cdr_triangulate_u8: bad sizes");
  const long long total = n_poses * joints;
  if (total == 0) return CDR_OK;
  triangulate_u8_kernel<<<(unsigned)ceil_div<long long>(total, 64), 64, 0, (cudaStream_t)stream>>>(
      P1, P2, p_batched ? p_rows * 4 : 0, pts1, pts2, total, joints, xyz);
  CDR_LAUNCH_OK("triangulate_u8_kernel");
  return CDR_OK;
}

extern "C" size_t cdr_mpjpe_scratch_bytes(long long n) {
  if (n < 0) n = 0;
  return ((size_t)n * 3 + (size_t)(kRedMaxBlocks + 1) * 3) * sizeof(double);
}

extern "C" int cdr_mpjpe_partial(const void* pred2d_l, const void* pred2d_r, const void* pred3d,
                                 int pred_is_f64, const double* gt3d, const double* gt2d_l,
                                 const double* gt2d_r, const double* weight, int weight_batched,
                                 int weight_f32_product, long long n, int joints, double* sums,
                                 void* scratch, void* stream) {
  CDR_CHECK_ARG(pred2d_l && pred2d_r && pred3d && gt3d && gt2d_l && gt2d_r && sums && scratch,
                "cdr_mpjpe_partial: null pointer");
  CDR_CHECK_ARG(n > 0 && joints > 0, "cdr_mpjpe_partial: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  double* pose_err = (double*)scratch;
  double* partials = pose_err + (size_t)n * 3;
  const unsigned grid = (unsigned)ceil_div<long long>(n, 128);
  if (pred_is_f64)
    mpjpe_pose_kernel<double><<<grid, 128, 0, st>>>(
        (const double*)pred2d_l, (const double*)pred2d_r, (const double*)pred3d, gt3d, gt2d_l,
        gt2d_r, weight, weight_batched, 0, n, joints, pose_err);
  else
    mpjpe_pose_kernel<float><<<grid, 128, 0, st>>>(
        (const float*)pred2d_l, (const float*)pred2d_r, (const float*)pred3d, gt3d, gt2d_l,
        gt2d_r, weight, weight_batched, weight_f32_product, n, joints, pose_err);
  CDR_LAUNCH_OK("mpjpe_pose_kernel");
  return reduce_pose_err(pose_err, n, joints, sums, partials, st);
}

extern "C" int cdr_mpjpe_reduce(const double* pose_err, long long n, int joints, double* sums,
                                void* scratch, void* stream) {
  CDR_CHECK_ARG(pose_err && sums && scratch && n > 0 && joints > 0, "cdr_mpjpe_reduce: bad args");
  return reduce_pose_err(pose_err, n, joints, sums, (double*)scratch, (cudaStream_t)stream);
}
