// fp32 implicit-GEMM ("tap-GEMM") on CUDA cores: the parity-precision convolutions.
//
// Every convolution of the head is C[m, n] = sum_taps sum_k A[shift_tap(m), k] * W[tap][k][n]
// over pixel-major activations (row m = one pixel, K = channels contiguous):
//   - 1x1 convs (models/cdrnet.py:17-43, models/decoder.py:15-21): one tap, no shift;
//   - ConvTranspose2d k4 s2 p1 (models/decoder.py:23-37): output phase (py,px) = (oy&1, ox&1)
//     is a 2x2-tap stride-1 conv of the input: tap (ty,tx) reads input (y+py-ty, x+px-tx)
//     with kernel element (ky,kx) = (1-py+2ty, 1-px+2tx)   [oy = 2*iy - 1 + ky].
//     blockIdx.z selects the phase; out-of-range taps contribute zero.
// BN(eval) is folded into W/bias when the weights are packed (pack.cu); the epilogue adds
// the bias, applies ReLU and writes pixel-major rows, the deconv phase scatter, or planar
// NCHW heat-maps.  128x128x16 tiles, 8x8 register micro-tiles, double-buffered smem.
#include "kernels.h"

namespace cdr {

constexpr int kBK = 16;
constexpr int kGemmThreads = 256;

template <int BM, int BN, int TN>
__global__ void __launch_bounds__(kGemmThreads, 2)
tap_gemm_ffma_kernel(const TapGemmParams p) {
  constexpr int TM = 8;
  constexpr int TX = BN / TN;            // threads along N
  static_assert((BM / TM) * TX == kGemmThreads, "tile/thread mismatch");
  constexpr int LA = BM * (kBK / 4) / kGemmThreads;            // float4 A loads per thread
  constexpr int LB = (kBK * BN / 4 + kGemmThreads - 1) / kGemmThreads;
  constexpr int AP = BM + 4, BP = BN + 4;

  __shared__ __align__(16) float As[2][kBK][AP];
  __shared__ __align__(16) float Bs[2][kBK][BP];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int g = blockIdx.z;
  const int HW = p.H * p.W;
  const int M = p.n_img * HW;
  const int ntaps = p.deconv ? 4 : 1;
  const int py = p.deconv ? (g >> 1) : 0, px = p.deconv ? (g & 1) : 0;

  const float* __restrict__ A =
      reinterpret_cast<const float*>(p.A) + (p.deconv ? 0 : (size_t)g * p.a_group_stride);
  const float* __restrict__ Wp = reinterpret_cast<const float*>(p.Wp) + (size_t)g * p.w_group_stride;

  // per-thread A rows
  int a_img[LA], a_y[LA], a_x[LA];
  bool a_ok[LA];
#pragma unroll
  for (int i = 0; i < LA; ++i) {
    const int r = (tid + i * kGemmThreads) >> 2;
    const int m = m0 + r;
    a_ok[i] = m < M;
    const int mm = a_ok[i] ? m : 0;
    a_img[i] = mm / HW;
    const int pix = mm - a_img[i] * HW;
    a_y[i] = pix / p.W;
    a_x[i] = pix - a_y[i] * p.W;
  }
  const int chunks_per_tap = p.cin / kBK;
  const int KT = ntaps * chunks_per_tap;

  float4 ra[LA], rb[LB];
  auto load_global = [&](int kt) {
    const int tap = kt / chunks_per_tap;
    const int ci0 = (kt - tap * chunks_per_tap) * kBK;
    const int dy = py - (tap >> 1), dx = px - (tap & 1);
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int kq = (tid + i * kGemmThreads) & 3;
      const int yy = a_y[i] + dy, xx = a_x[i] + dx;
      const bool ok = a_ok[i] && (unsigned)yy < (unsigned)p.H && (unsigned)xx < (unsigned)p.W;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const size_t off = ((size_t)(a_img[i] * p.H + yy) * p.W + xx) * p.a_pitch + ci0 + kq * 4;
        ra[i] = __ldg(reinterpret_cast<const float4*>(A + off));
      }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + i * kGemmThreads;
      if (e < kBK * BN / 4) {
        const int k = e / (BN / 4), c4 = e % (BN / 4);
        rb[i] = __ldg(reinterpret_cast<const float4*>(
            Wp + (size_t)(kt * kBK + k) * p.n_pad + n0 + c4 * 4));
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      const int e = tid + i * kGemmThreads;
      const int r = e >> 2, kq = e & 3;
      As[buf][kq * 4 + 0][r] = ra[i].x;
      As[buf][kq * 4 + 1][r] = ra[i].y;
      As[buf][kq * 4 + 2][r] = ra[i].z;
      As[buf][kq * 4 + 3][r] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      const int e = tid + i * kGemmThreads;
      if (e < kBK * BN / 4) {
        const int k = e / (BN / 4), c4 = e % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][k][c4 * 4]) = rb[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < KT) load_global(kt + 1);
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      float a[TM], b[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) =
          *reinterpret_cast<const float4*>(&As[cur][k][BM / 2 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      if constexpr (TN == 8)
        *reinterpret_cast<float4*>(&b[4]) =
            *reinterpret_cast<const float4*>(&Bs[cur][k][BN / 2 + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_smem(cur ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  const float* __restrict__ bias = p.bias ? p.bias + (size_t)g * p.bias_group_stride : nullptr;
  float* __restrict__ C =
      reinterpret_cast<float*>(p.C) + (p.deconv ? 0 : (size_t)g * p.c_group_stride);
#pragma unroll
  for (int jh = 0; jh < TN / 4; ++jh) {
    const int col = n0 + jh * (BN / 2) + tx * 4;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(bias + col));
      bv[0] = q.x; bv[1] = q.y; bv[2] = q.z; bv[3] = q.w;
    }
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
      const int row0 = m0 + ih * (BM / 2) + ty * 4;
      float v[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x = acc[ih * 4 + i][jh * 4 + j] + bv[j];
          if (p.relu) x = fmaxf(x, 0.f);
          v[i][j] = (col + j < p.n) ? x : 0.f;
        }
      if (p.out_mode == kOutPlanar) {
        // NCHW: for each channel, 4 consecutive pixels (row0 % 4 == 0, HW % 4 == 0)
        if (row0 < M) {
          const int img = row0 / HW, pix = row0 - img * HW;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + j < p.n)
              *reinterpret_cast<float4*>(C + ((size_t)img * p.n + col + j) * HW + pix) =
                  make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
        }
      } else if (col < p.c_fill) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = row0 + i;
          if (m >= M) continue;
          size_t orow = (size_t)m;
          if (p.out_mode == kOutDeconv) {
            const int img = m / HW, pix = m - img * HW;
            const int y = pix / p.W, x = pix - y * p.W;
            orow = ((size_t)img * 2 * p.H + 2 * y + py) * (2 * p.W) + 2 * x + px;
          }
          *reinterpret_cast<float4*>(C + orow * p.c_pitch + col) =
              make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        }
      }
    }
  }
}

int launch_tap_gemm_ffma(const TapGemmParams& p, int groups, cudaStream_t st) {
  const int M = p.n_img * p.H * p.W;
  CDR_CHECK_ARG(p.cin % kBK == 0 && p.a_pitch % 4 == 0 && p.c_pitch % 4 == 0 && p.c_fill % 4 == 0,
                "tap_gemm_ffma: cin %% 16, pitches %% 4 and c_fill %% 4 must be 0 (cin=%d a_pitch=%d "
                "c_pitch=%d c_fill=%d)", p.cin, p.a_pitch, p.c_pitch, p.c_fill);
  CDR_CHECK_ARG(M > 0 && (p.out_mode != kOutPlanar || (p.H * p.W) % 4 == 0), "tap_gemm_ffma: bad M/HW");
  if (p.n <= 32) {
    CDR_CHECK_ARG(p.n_pad % 32 == 0, "tap_gemm_ffma: n_pad %% 32 != 0");
    dim3 grid(ceil_div(M, 256), p.n_pad / 32, groups);
    tap_gemm_ffma_kernel<256, 32, 4><<<grid, kGemmThreads, 0, st>>>(p);
  } else {
    CDR_CHECK_ARG(p.n_pad % 128 == 0, "tap_gemm_ffma: n_pad %% 128 != 0");
    dim3 grid(ceil_div(M, 128), p.n_pad / 128, groups);
    tap_gemm_ffma_kernel<128, 128, 8><<<grid, kGemmThreads, 0, st>>>(p);
  }
  CDR_LAUNCH_OK("tap_gemm_ffma_kernel");
  return CDR_OK;
}

}  // namespace cdr
