// ResNet stem on B200: conv 7x7 stride 2 pad 3 (3 -> 64) + eval BN + ReLU, then max-pool 3x3 stride 2
// pad 1 (reference models/encoder.py:93-97,122-125), fp32 NCHW images in, bf16 NHWC rows out — the
// layout the tcgen05 bottleneck kernels read.
//
// Cin = 3 rules out the TMA/UMMA operand path (a pixel is 6 bytes), and K = 147 is tiny, so the conv
// is an implicit GEMM on warp-level mma.sync (m16n8k16 bf16, fp32 accumulate) fed from a shared-memory
// image patch: A[m][k] = patch[base(m) + off(k)], with the K axis laid out as 7 rows of 22 (21 taps*
// channels + 1 zero pad) so that a fragment's (k, k+1) pair is one aligned 32-bit LDS.  HBM-bound by
// design (100 MB in, 268 MB out per 128 images); the pool is a second, purely streaming kernel.
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.h"

namespace cdr {

constexpr int kStemCo = 64;
constexpr int kStemKy = 7, kStemRow = 22;          // K' = ky * 22 + (kx * 3 + ci), slot 21 of a row is a zero weight
constexpr int kStemK = 160;                        // 7 * 22 = 154, padded to a multiple of 16
constexpr int kStemWPitch = 168;                   // weight row pitch (elements): conflict-free B-fragment loads
constexpr int kTileH = 8, kTileW = 32;             // conv-output tile of one CTA pass: one row per warp
constexpr int kPatchH = 2 * kTileH + 5;            // 21 input rows
constexpr int kPatchW = 2 * kTileW + 5;            // 69 input columns
constexpr int kPatchPitch = 208;                   // 69 * 3 = 207 elements per patch row, padded to even

// conv1 (64,3,7,7) + bn1 -> bf16 [64][kStemWPitch] in the K' order above, fp32 folded bias [64]
__global__ void pack_stem_kernel(CdrConvBn s, __nv_bfloat16* __restrict__ w, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kStemCo * kStemWPitch) return;
  const int n = idx / kStemWPitch, k = idx - n * kStemWPitch;
  const double sc = (double)s.bn_weight[n] / sqrt((double)s.bn_var[n] + 1e-5);
  if (k == 0) bias[n] = (float)((double)s.bn_bias[n] - (double)s.bn_mean[n] * sc);
  float v = 0.f;
  const int ky = k / kStemRow, r = k - ky * kStemRow;
  if (ky < kStemKy && r < 21) {
    const int kx = r / 3, ci = r - kx * 3;
    v = (float)((double)s.weight[((n * 3 + ci) * 7 + ky) * 7 + kx] * sc);
  }
  w[idx] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Input pixel fetch.  kU8 = false: fp32 NCHW tensors as the reference feeds its model (already normalised).
// kU8 = true: raw uint8 HWC frames with torchvision's ToTensor + Normalize (inference.py:40-44) applied on
// the fly in the same fp32 operation order — x / 255, - mean, / std — so the bf16 patch is bit-identical
// to converting the reference's preprocessed tensor (4x fewer bytes over PCIe, no
// host-side torchvision pass).
struct StemNorm {
  float mean[3], std[3];
};

// grid (conv rows / 8, images); each CTA walks the W/64 column tiles of its row band.
template <bool kU8>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const void* __restrict__ xin, int H, int W, const StemNorm nrm, const __nv_bfloat16* __restrict__ wpk,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) __nv_bfloat16 s_w[kStemCo * kStemWPitch];     // 21.5 KB
  __shared__ __align__(16) __nv_bfloat16 s_p[kPatchH * kPatchPitch];     // 8.7 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int img = blockIdx.y, oy0 = blockIdx.x * kTileH;
  const int Ho = H >> 1, Wo = W >> 1;
  for (int i = threadIdx.x; i < kStemCo * kStemWPitch / 8; i += 256)
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(wpk) + i);
  float bv[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bv[nt][0] = __ldg(bias + nt * 8 + 2 * t);
    bv[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
  }
  // per-lane patch offsets of the two k-pairs of every k-step: off(k') = ky * pitch + r
  int koff[kStemK / 16][2];
#pragma unroll
  for (int s = 0; s < kStemK / 16; ++s) {
#pragma unroll
    for (int hk = 0; hk < 2; ++hk) {
      const int k = 16 * s + 8 * hk + 2 * t;
      const int ky = k / kStemRow, r = k - ky * kStemRow;
      koff[s][hk] = ky < kStemKy ? ky * kPatchPitch + r : 0;       // padded k: weight is zero, any address does
    }
  }
  const float* xi = reinterpret_cast<const float*>(xin) + (size_t)img * 3 * H * W;
  const uint8_t* xu = reinterpret_cast<const uint8_t*>(xin) + (size_t)img * 3 * H * W;
  for (int ox0 = 0; ox0 < Wo; ox0 += kTileW) {
    __syncthreads();                                   // previous pass done with the patch (and s_w visible)
    const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
    if constexpr (!kU8) {
      for (int i = threadIdx.x; i < kPatchH * 3 * kPatchW; i += 256) {
        const int col = i % kPatchW, rc = i / kPatchW;
        const int ci = rc % 3, row = rc / 3;
        const int iy = iy0 + row, ix = ix0 + col;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xi + ((size_t)ci * H + iy) * W + ix);
        s_p[row * kPatchPitch + col * 3 + ci] = __float2bfloat16_rn(v);
      }
    } else {
      for (int i = threadIdx.x; i < kPatchH * 3 * kPatchW; i += 256) {
        const int e = i % (3 * kPatchW), row = i / (3 * kPatchW);     // e = col * 3 + ci: contiguous in HWC
        const int col = e / 3, ci = e - col * 3;
        const int iy = iy0 + row, ix = ix0 + col;
        float v = 0.f;                                                 // zero padding applies AFTER normalisation
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const float u = (float)__ldg(xu + ((size_t)iy * W + ix) * 3 + ci);
          v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), nrm.mean[ci]), nrm.std[ci]);
        }
        s_p[row * kPatchPitch + e] = __float2bfloat16_rn(v);
      }
    }
    if (threadIdx.x < kPatchH) s_p[threadIdx.x * kPatchPitch + 207] = __float2bfloat16_rn(0.f);
    __syncthreads();
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    // this warp: conv row `warp` of the tile, columns mt*16 + {g, g+8}
    const int base0 = (2 * warp) * kPatchPitch + 6 * g;          // pixel (warp, g): patch origin, in elements
#pragma unroll
    for (int s = 0; s < kStemK / 16; ++s) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int b = base0 + 6 * 16 * mt;
        a[mt][0] = *reinterpret_cast<const uint32_t*>(s_p + b + koff[s][0]);
        a[mt][1] = *reinterpret_cast<const uint32_t*>(s_p + b + 48 + koff[s][0]);       // column g + 8
        a[mt][2] = *reinterpret_cast<const uint32_t*>(s_p + b + koff[s][1]);
        a[mt][3] = *reinterpret_cast<const uint32_t*>(s_p + b + 48 + koff[s][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const __nv_bfloat16* wr = s_w + (nt * 8 + g) * kStemWPitch + 16 * s + 2 * t;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma_bf16_16816(acc[0][nt], a[0], b0, b1);
        mma_bf16_16816(acc[1][nt], a[1], b0, b1);
      }
    }
    // + bias, ReLU, bf16, NHWC: pixel (oy0 + warp, ox0 + mt*16 + g [+8]), channels nt*8 + 2t, +1
    const int oy = oy0 + warp;
    if (oy < Ho) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const int ox = ox0 + mt * 16 + g + 8 * hr;
          if (ox < Wo) {
            __nv_bfloat16* o = out + (((size_t)img * Ho + oy) * Wo + ox) * kStemCo + 2 * t;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
              const float v0 = fmaxf(acc[mt][nt][2 * hr] + bv[nt][0], 0.f);
              const float v1 = fmaxf(acc[mt][nt][2 * hr + 1] + bv[nt][1], 0.f);
              *reinterpret_cast<__nv_bfloat162*>(o + nt * 8) = __floats2bfloat162_rn(v0, v1);
            }
          }
        }
      }
    }
  }
}

// max-pool 3x3 stride 2 pad 1 on NHWC bf16 (values are post-ReLU, so 0 is the identity of the max)
__global__ void __launch_bounds__(256)
maxpool3s2_nhwc_kernel(const __nv_bfloat16* __restrict__ in, int Hi, int Wi, int C8, long long total,
                       __nv_bfloat16* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  long long r = idx / C8;
  const int Wo = Wi >> 1, Ho = Hi >> 1;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho);
  const long long img = r / Ho;
  __nv_bfloat162 m[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) m[e] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= Hi) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= Wi) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((img * Hi + iy) * Wi + ix) * (long long)(C8 * 8)) + c8);
      const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], v[e]);
    }
  }
  reinterpret_cast<uint4*>(out + ((img * Ho + oy) * Wo + ox) * (long long)(C8 * 8))[c8] = *reinterpret_cast<const uint4*>(m);
}

int launch_pack_stem(const CdrConvBn& s, void* w, float* bias, cudaStream_t st) {
  CDR_CHECK_ARG(s.weight && s.bn_weight && s.bn_bias && s.bn_mean && s.bn_var && w && bias, "pack_stem: bad args");
  pack_stem_kernel<<<ceil_div(kStemCo * kStemWPitch, 256), 256, 0, st>>>(s, (__nv_bfloat16*)w, bias);
  CDR_LAUNCH_OK("pack_stem_kernel");
  return CDR_OK;
}
size_t stem_weight_bytes() { return (size_t)kStemCo * kStemWPitch * sizeof(__nv_bfloat16); }

// x (n,3,H,W) fp32 -> conv_out (n,H/2,W/2,64) bf16 scratch -> pooled (n,H/4,W/4,64) bf16
int launch_stem(const void* x, int is_u8, const float* mean, const float* std, int n, int H, int W, const void* w,
                const float* bias, void* conv_out, void* pooled, cudaStream_t st) {
  CDR_CHECK_ARG(x && w && bias && conv_out && pooled && n > 0, "stem: bad args");
  CDR_CHECK_ARG(!is_u8 || (mean && std), "stem: uint8 frames need mean / std");
  StemNorm nrm{};
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = is_u8 ? mean[c] : 0.f;
    nrm.std[c] = is_u8 ? std[c] : 1.f;
  }
  CDR_CHECK_ARG(H % 16 == 0 && W % 64 == 0, "stem: image %dx%d must have H %% 16 == 0 and W %% 64 == 0", H, W);
  CDR_CHECK_ARG(((uintptr_t)conv_out & 15) == 0 && ((uintptr_t)pooled & 15) == 0 && ((uintptr_t)w & 15) == 0, "stem: alignment");
  const int Ho = H / 2, Wo = W / 2;
  if (is_u8)
    stem_conv_kernel<true><<<dim3(Ho / kTileH, n), 256, 0, st>>>(x, H, W, nrm, (const __nv_bfloat16*)w, bias,
                                                               (__nv_bfloat16*)conv_out);
  else
    stem_conv_kernel<false><<<dim3(Ho / kTileH, n), 256, 0, st>>>(x, H, W, nrm, (const __nv_bfloat16*)w, bias,
                                                                (__nv_bfloat16*)conv_out);
  CDR_LAUNCH_OK("stem_conv_kernel");
  const long long total = (long long)n * (Ho / 2) * (Wo / 2) * (kStemCo / 8);
  maxpool3s2_nhwc_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
      (const __nv_bfloat16*)conv_out, Ho, Wo, kStemCo / 8, total, (__nv_bfloat16*)pooled);
  CDR_LAUNCH_OK("maxpool3s2_nhwc_kernel");
  return CDR_OK;
}


// ------------------------------------------------------------------------------------------
// The same stem at fp32 accuracy for the f16x2 ("fp32") encoder, first version (kept as the CDR_STEM_FFMA=1 cross-check of the
// tensor-core form further down): fp32 FFMA (the conv is 2.3 % of the encoder's
// FLOPs), fp32 NHWC scratch, then the max-pool writes the scaled fp16 hi/lo planes the tcgen05 layers read.  The
// pooled tensor's scale comes from the EXACT maximum (max-pool commutes with max): the conv kernel atomicMax-es its
// post-ReLU outputs into slot[0], the pool kernel turns that into slot[1] = 2^(13 - ilogb(amax)).
constexpr int kStemK32 = 147;                      // (ky, kx, ci) = 7 * 21
constexpr int kPatchPitch32 = 208;                 // floats per patch row: 69 * 3 = 207 (+1)
constexpr int kStem32SmemBytes = (kStemK32 * kStemCo + kPatchH * kPatchPitch32) * (int)sizeof(float);

// conv1 (64,3,7,7) + bn1 -> fp32 [ky*21 + kx*3 + ci][64] with channels PERMUTED inside a row: position
// 16*j + 4*cg + e holds channel 16*j + 4*cg + e (identity) — kept plain; the kernel's lanes read float4s at 4*cg + 16*j
__global__ void pack_stem_f32_kernel(CdrConvBn s, float* __restrict__ w, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kStemK32 * kStemCo) return;
  const int k = idx / kStemCo, n = idx - k * kStemCo;
  const double sc = (double)s.bn_weight[n] / sqrt((double)s.bn_var[n] + 1e-5);
  if (k == 0) bias[n] = (float)((double)s.bn_bias[n] - (double)s.bn_mean[n] * sc);
  const int ky = k / 21, r = k - ky * 21, kx = r / 3, ci = r - kx * 3;
  w[idx] = (float)((double)s.weight[((n * 3 + ci) * 7 + ky) * 7 + kx] * sc);
}

// grid (conv rows / 8, images), 256 threads: warp = one conv row of the 8 x 32 tile; lane = (pxg = lane >> 2: columns
// pxg + 8*i, i < 4; cg = lane & 3: channels 16*j + 4*cg .. +3, j < 4) -> 4 x 16 outputs per thread.  Patch reads of a
// warp hit 8 distinct banks (stride 6 floats), weight reads are one 64-byte segment per float4 load.
template <bool kU8>
__global__ void __launch_bounds__(256, 2)      // two CTAs per SM: one loads its patch while the other runs its FMAs
stem_conv_f32_kernel(const void* __restrict__ xin, int H, int W, const StemNorm nrm, const float* __restrict__ wpk,
                     const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ amax_out) {
  extern __shared__ __align__(16) float smem_f[];
  float* s_w = smem_f;                               // [147][64]
  float* s_p = smem_f + kStemK32 * kStemCo;          // [21][208]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = lane & 3, pxg = lane >> 2;
  const int img = blockIdx.y, oy0 = blockIdx.x * kTileH;
  const int Ho = H >> 1, Wo = W >> 1;
  for (int i = threadIdx.x; i < kStemK32 * kStemCo / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wpk) + i);
  float4 bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias + 16 * j + 4 * cg));
  const float* xi = reinterpret_cast<const float*>(xin) + (size_t)img * 3 * H * W;
  const uint8_t* xu = reinterpret_cast<const uint8_t*>(xin) + (size_t)img * 3 * H * W;
  float mx = 0.f;
  for (int ox0 = 0; ox0 < Wo; ox0 += kTileW) {
    __syncthreads();
    const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
    if constexpr (!kU8) {
      for (int i = threadIdx.x; i < kPatchH * 3 * kPatchW; i += 256) {
        const int col = i % kPatchW, rc = i / kPatchW;
        const int ci = rc % 3, row = rc / 3;
        const int iy = iy0 + row, ix = ix0 + col;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xi + ((size_t)ci * H + iy) * W + ix);
        s_p[row * kPatchPitch32 + col * 3 + ci] = v;
      }
    } else {
      for (int i = threadIdx.x; i < kPatchH * 3 * kPatchW; i += 256) {
        const int e = i % (3 * kPatchW), row = i / (3 * kPatchW);
        const int col = e / 3, ci = e - col * 3;
        const int iy = iy0 + row, ix = ix0 + col;
        float v = 0.f;                                                 // zero padding applies AFTER normalisation
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const float u = (float)__ldg(xu + ((size_t)iy * W + ix) * 3 + ci);
          v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), nrm.mean[ci]), nrm.std[ci]);
        }
        s_p[row * kPatchPitch32 + e] = v;
      }
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[i][c] = 0.f;
    // pixel (warp, pxg + 8 i): patch origin (2*warp, 2*(pxg + 8 i)) -> element offset 2*warp*pitch + 6*(pxg + 8 i)
    const float* pr = s_p + (2 * warp) * kPatchPitch32 + 6 * pxg;
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      const float* prow = pr + ky * kPatchPitch32;
      const float* wrow = s_w + (ky * 21) * kStemCo + 4 * cg;
#pragma unroll 7
      for (int r = 0; r < 21; ++r) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = prow[48 * i + r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w4 = *reinterpret_cast<const float4*>(wrow + r * kStemCo + 16 * j);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][4 * j] = fmaf(a[i], w4.x, acc[i][4 * j]);
            acc[i][4 * j + 1] = fmaf(a[i], w4.y, acc[i][4 * j + 1]);
            acc[i][4 * j + 2] = fmaf(a[i], w4.z, acc[i][4 * j + 2]);
            acc[i][4 * j + 3] = fmaf(a[i], w4.w, acc[i][4 * j + 3]);
          }
        }
      }
    }
    const int oy = oy0 + warp;
    if (oy < Ho) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ox = ox0 + pxg + 8 * i;
        if (ox < Wo) {
          float* o = out + (((size_t)img * Ho + oy) * Wo + ox) * kStemCo + 4 * cg;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v;
            v.x = fmaxf(acc[i][4 * j] + bv[j].x, 0.f);
            v.y = fmaxf(acc[i][4 * j + 1] + bv[j].y, 0.f);
            v.z = fmaxf(acc[i][4 * j + 2] + bv[j].z, 0.f);
            v.w = fmaxf(acc[i][4 * j + 3] + bv[j].w, 0.f);
            mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            *reinterpret_cast<float4*>(o + 16 * j) = v;
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(amax_out), __float_as_uint(mx));   // mx >= 0
}

// max-pool 3x3 stride 2 pad 1 on NHWC fp32 (post-ReLU) -> scaled fp16 hi/lo planes; slot = {amax (complete), scale}
__global__ void __launch_bounds__(256)
maxpool3s2_f16p_kernel(const float* __restrict__ in, int Hi, int Wi, int C8, long long total, __half* __restrict__ hi,
                       __half* __restrict__ lo, float* __restrict__ slot) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float a = __ldg(slot);
  const float s = (a > 0.f && a < 3.0e38f) ? ldexpf(1.f, 13 - ilogbf(a)) : 1.f;
  if (idx == 0) slot[1] = s;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  long long r = idx / C8;
  const int Wo = Wi >> 1, Ho = Hi >> 1;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho);
  const long long img = r / Ho;
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= Hi) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= Wi) continue;
      const float4* q = reinterpret_cast<const float4*>(in + ((img * Hi + iy) * Wi + ix) * (long long)(C8 * 8)) + 2 * c8;
      const float4 q0 = __ldg(q), q1 = __ldg(q + 1);
      m[0] = fmaxf(m[0], q0.x); m[1] = fmaxf(m[1], q0.y); m[2] = fmaxf(m[2], q0.z); m[3] = fmaxf(m[3], q0.w);
      m[4] = fmaxf(m[4], q1.x); m[5] = fmaxf(m[5], q1.y); m[6] = fmaxf(m[6], q1.z); m[7] = fmaxf(m[7], q1.w);
    }
  }
  __align__(16) __half h8[8];
  __align__(16) __half l8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float X = m[e] * s;
    h8[e] = __float2half_rn(X);
    l8[e] = __float2half_rn((X - __half2float(h8[e])) * 2048.f);
  }
  const long long o = ((img * Ho + oy) * Wo + ox) * (long long)(C8 * 8) + 8 * c8;
  *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(h8);
  *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(l8);
}


// ------------------------------------------------------------------------------------------
// The fp32-accurate stem on the tensor cores (round 2, replaces the FFMA kernel above in the f16x2 encoder: 1.02 ->
// see DESIGN.md §7): the bf16 kernel's implicit GEMM on warp-level mma.sync.m16n8k16, with BOTH operands as fp16 hi/lo
// pairs and three MMAs per product — x w = xh wh + 2^-11 (xh wl + xl wh) (+ 2^-22 xl wl, dropped) — exactly the f16x2
// scheme of gemm_tc.cu.  fp16 products are exact in the fp32 accumulators; main and correction terms accumulate
// separately (K = 147: no long same-signed chain).  Pixels need no scale (|x| < 65504; the lo plane keeps values down
// to 2^-36 absolute), weights carry one power-of-two scale per output channel (2^-t undone in the epilogue).
constexpr int kT16W = 16;                          // conv-output tile of one CTA pass: 8 rows x 16 columns, one row per warp
constexpr int kP16W = 2 * kT16W + 5;               // 37 input columns
constexpr int kP16Pitch = 112;                     // 37 * 3 = 111 elements per patch row, padded to even
constexpr int kStem16SmemBytes = (2 * kStemCo * kStemWPitch + 2 * kPatchH * kP16Pitch) * (int)sizeof(__half);

// conv1 (64,3,7,7) + bn1 -> fp16 hi / lo [64][kStemWPitch] in the K' order of the bf16 kernel, wsi[64] = 2^-t, bias[64]
__global__ void pack_stem_f16x2_kernel(CdrConvBn s, __half* __restrict__ w_hi, __half* __restrict__ w_lo,
                                       float* __restrict__ wsi, float* __restrict__ bias) {
  const int n = blockIdx.x;                        // one block per output channel
  __shared__ float red[8];
  const double sc = (double)s.bn_weight[n] / sqrt((double)s.bn_var[n] + 1e-5);
  auto val = [&](int k) -> float {
    const int ky = k / kStemRow, r = k - ky * kStemRow;
    if (ky >= kStemKy || r >= 21) return 0.f;
    const int kx = r / 3, ci = r - kx * 3;
    return (float)((double)s.weight[((n * 3 + ci) * 7 + ky) * 7 + kx] * sc);
  };
  float mx = 0.f;
  for (int k = threadIdx.x; k < kStemWPitch; k += blockDim.x) mx = fmaxf(mx, fabsf(val(k)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  const int t = mx > 0.f ? 13 - ilogbf(mx) : 0;
  if (threadIdx.x == 0) {
    wsi[n] = ldexpf(1.f, -t);
    bias[n] = (float)((double)s.bn_bias[n] - (double)s.bn_mean[n] * sc);
  }
  const float up = ldexpf(1.f, t);
  for (int k = threadIdx.x; k < kStemWPitch; k += blockDim.x) {
    const float X = val(k) * up;
    const __half h = __float2half_rn(X);
    w_hi[n * kStemWPitch + k] = h;
    w_lo[n * kStemWPitch + k] = __float2half_rn((X - __half2float(h)) * 2048.f);
  }
}

__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (conv rows / 8, images); each CTA walks the W/32 column tiles of its row band; fp32 NHWC output + amax
template <bool kU8>
__global__ void __launch_bounds__(256, 2)
stem_conv_f16x2_kernel(const void* __restrict__ xin, int H, int W, const StemNorm nrm, const __half* __restrict__ wpk_hi,
                       const __half* __restrict__ wpk_lo, const float* __restrict__ wsi, const float* __restrict__ bias,
                       float* __restrict__ out, float* __restrict__ amax_out) {
  extern __shared__ __align__(16) unsigned char smem_h[];
  __half* s_wh = reinterpret_cast<__half*>(smem_h);                 // [64][168]
  __half* s_wl = s_wh + kStemCo * kStemWPitch;
  __half* s_ph = s_wl + kStemCo * kStemWPitch;                      // [21][112]
  __half* s_pl = s_ph + kPatchH * kP16Pitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int img = blockIdx.y, oy0 = blockIdx.x * kTileH;
  const int Ho = H >> 1, Wo = W >> 1;
  for (int i = threadIdx.x; i < kStemCo * kStemWPitch / 8; i += 256) {
    reinterpret_cast<uint4*>(s_wh)[i] = __ldg(reinterpret_cast<const uint4*>(wpk_hi) + i);
    reinterpret_cast<uint4*>(s_wl)[i] = __ldg(reinterpret_cast<const uint4*>(wpk_lo) + i);
  }
  float bv[8][2], sv[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bv[nt][0] = __ldg(bias + nt * 8 + 2 * t);
    bv[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
    sv[nt][0] = __ldg(wsi + nt * 8 + 2 * t);
    sv[nt][1] = __ldg(wsi + nt * 8 + 2 * t + 1);
  }
  int koff[kStemK / 16][2];
#pragma unroll
  for (int s = 0; s < kStemK / 16; ++s) {
#pragma unroll
    for (int hk = 0; hk < 2; ++hk) {
      const int k = 16 * s + 8 * hk + 2 * t;
      const int ky = k / kStemRow, r = k - ky * kStemRow;
      koff[s][hk] = ky < kStemKy ? ky * kP16Pitch + r : 0;         // padded k: weight is zero, any address does
    }
  }
  const float* xi = reinterpret_cast<const float*>(xin) + (size_t)img * 3 * H * W;
  const uint8_t* xu = reinterpret_cast<const uint8_t*>(xin) + (size_t)img * 3 * H * W;
  float mx = 0.f;
  for (int ox0 = 0; ox0 < Wo; ox0 += kT16W) {
    __syncthreads();                                   // previous pass done with the patch (and the weights visible)
    const int iy0 = 2 * oy0 - 3, ix0 = 2 * ox0 - 3;
    for (int i = threadIdx.x; i < kPatchH * 3 * kP16W; i += 256) {
      int row, e;
      float v = 0.f;
      if constexpr (!kU8) {
        const int col = i % kP16W, rc = i / kP16W;
        const int ci = rc % 3;
        row = rc / 3;
        e = col * 3 + ci;
        const int iy = iy0 + row, ix = ix0 + col;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xi + ((size_t)ci * H + iy) * W + ix);
      } else {
        e = i % (3 * kP16W);
        row = i / (3 * kP16W);
        const int col = e / 3, ci = e - col * 3;
        const int iy = iy0 + row, ix = ix0 + col;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {                // zero padding applies AFTER normalisation
          const float u = (float)__ldg(xu + ((size_t)iy * W + ix) * 3 + ci);
          v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), nrm.mean[ci]), nrm.std[ci]);
        }
      }
      const __half h = __float2half_rn(v);
      s_ph[row * kP16Pitch + e] = h;
      s_pl[row * kP16Pitch + e] = __float2half_rn((v - __half2float(h)) * 2048.f);
    }
    if (threadIdx.x < kPatchH) {
      s_ph[threadIdx.x * kP16Pitch + 111] = __float2half_rn(0.f);
      s_pl[threadIdx.x * kP16Pitch + 111] = __float2half_rn(0.f);
    }
    __syncthreads();
    float am[8][4], ac[8][4];                          // main (hi hi) and correction (hi lo + lo hi) accumulators
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) am[nt][e] = ac[nt][e] = 0.f;
    // this warp: conv row `warp` of the tile, columns {g, g + 8}
    const int base0 = (2 * warp) * kP16Pitch + 6 * g;
#pragma unroll
    for (int s = 0; s < kStemK / 16; ++s) {
      uint32_t ah[4], al[4];
      ah[0] = *reinterpret_cast<const uint32_t*>(s_ph + base0 + koff[s][0]);
      ah[1] = *reinterpret_cast<const uint32_t*>(s_ph + base0 + 48 + koff[s][0]);       // column g + 8
      ah[2] = *reinterpret_cast<const uint32_t*>(s_ph + base0 + koff[s][1]);
      ah[3] = *reinterpret_cast<const uint32_t*>(s_ph + base0 + 48 + koff[s][1]);
      al[0] = *reinterpret_cast<const uint32_t*>(s_pl + base0 + koff[s][0]);
      al[1] = *reinterpret_cast<const uint32_t*>(s_pl + base0 + 48 + koff[s][0]);
      al[2] = *reinterpret_cast<const uint32_t*>(s_pl + base0 + koff[s][1]);
      al[3] = *reinterpret_cast<const uint32_t*>(s_pl + base0 + 48 + koff[s][1]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int wo = (nt * 8 + g) * kStemWPitch + 16 * s + 2 * t;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(s_wh + wo), bh1 = *reinterpret_cast<const uint32_t*>(s_wh + wo + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(s_wl + wo), bl1 = *reinterpret_cast<const uint32_t*>(s_wl + wo + 8);
        mma_f16_16816(am[nt], ah, bh0, bh1);
        mma_f16_16816(ac[nt], ah, bl0, bl1);
        mma_f16_16816(ac[nt], al, bh0, bh1);
      }
    }
    // (main + 2^-11 corr) * 2^-t + bias, ReLU, fp32 NHWC: pixel (oy0 + warp, ox0 + g [+8]), channels nt*8 + 2t, +1
    const int oy = oy0 + warp;
    if (oy < Ho) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int ox = ox0 + g + 8 * hr;
        if (ox < Wo) {
          float* o = out + (((size_t)img * Ho + oy) * Wo + ox) * kStemCo + 2 * t;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const float v0 = fmaxf(fmaf(fmaf(ac[nt][2 * hr], 1.f / 2048.f, am[nt][2 * hr]), sv[nt][0], bv[nt][0]), 0.f);
            const float v1 = fmaxf(fmaf(fmaf(ac[nt][2 * hr + 1], 1.f / 2048.f, am[nt][2 * hr + 1]), sv[nt][1], bv[nt][1]), 0.f);
            mx = fmaxf(mx, fmaxf(v0, v1));
            *reinterpret_cast<float2*>(o + nt * 8) = make_float2(v0, v1);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(amax_out), __float_as_uint(mx));   // mx >= 0
}

// packed buffer of the fp32-accurate stem: [fp32 weights [147][64] (FFMA cross-check) | fp16 hi [64][168] | fp16 lo | wsi [64]]
constexpr size_t kStem32WBytes = (size_t)kStemK32 * kStemCo * sizeof(float);                     // 37632
constexpr size_t kStem16PlaneBytes = (size_t)kStemCo * kStemWPitch * sizeof(__half);             // 21504
static bool stem_use_ffma() {          // CDR_STEM_FFMA=1: the CUDA-core kernel (A/B timing, cross-check)
  const char* e = getenv("CDR_STEM_FFMA");
  return e && e[0] == '1';
}
int launch_pack_stem_f32(const CdrConvBn& s, float* w, float* bias, cudaStream_t st) {
  CDR_CHECK_ARG(s.weight && s.bn_weight && s.bn_bias && s.bn_mean && s.bn_var && w && bias, "pack_stem: bad args");
  pack_stem_f32_kernel<<<ceil_div(kStemK32 * kStemCo, 256), 256, 0, st>>>(s, w, bias);
  CDR_LAUNCH_OK("pack_stem_f32_kernel");
  uint8_t* b = reinterpret_cast<uint8_t*>(w) + kStem32WBytes;
  pack_stem_f16x2_kernel<<<kStemCo, 64, 0, st>>>(s, reinterpret_cast<__half*>(b), reinterpret_cast<__half*>(b + kStem16PlaneBytes),
                                                 reinterpret_cast<float*>(b + 2 * kStem16PlaneBytes), bias);
  CDR_LAUNCH_OK("pack_stem_f16x2_kernel");
  return CDR_OK;
}
size_t stem_weight_bytes_f32() { return kStem32WBytes + 2 * kStem16PlaneBytes + kStemCo * sizeof(float); }

int launch_stem_f32(const void* x, int is_u8, const float* mean, const float* std, int n, int H, int W, const float* w,
                    const float* bias, float* conv_out, void* pooled_hi, void* pooled_lo, float* slot, cudaStream_t st) {
  CDR_CHECK_ARG(x && w && bias && conv_out && pooled_hi && pooled_lo && slot && n > 0, "stem: bad args");
  CDR_CHECK_ARG(!is_u8 || (mean && std), "stem: uint8 frames need mean / std");
  StemNorm nrm{};
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = is_u8 ? mean[c] : 0.f;
    nrm.std[c] = is_u8 ? std[c] : 1.f;
  }
  CDR_CHECK_ARG(H % 16 == 0 && W % 64 == 0, "stem: image %dx%d must have H %% 16 == 0 and W %% 64 == 0", H, W);
  CDR_CHECK_ARG(((uintptr_t)conv_out & 15) == 0 && ((uintptr_t)pooled_hi & 15) == 0 && ((uintptr_t)pooled_lo & 15) == 0 &&
                    ((uintptr_t)w & 15) == 0, "stem: alignment");
  static DeviceOnce attr_set[2], attr16_set[2];
  const int Ho = H / 2, Wo = W / 2;
  if (!stem_use_ffma()) {
    // tensor-core form: fp16 hi/lo operands, 3 mma.sync per product
    const uint8_t* b = reinterpret_cast<const uint8_t*>(w) + kStem32WBytes;
    const __half* wh = reinterpret_cast<const __half*>(b);
    const __half* wl = reinterpret_cast<const __half*>(b + kStem16PlaneBytes);
    const float* wsi = reinterpret_cast<const float*>(b + 2 * kStem16PlaneBytes);
    if (is_u8) {
      if (attr16_set[1].need()) {
        CDR_CUDA(cudaFuncSetAttribute(stem_conv_f16x2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStem16SmemBytes));
        attr16_set[1].done();
      }
      stem_conv_f16x2_kernel<true><<<dim3(Ho / kTileH, n), 256, kStem16SmemBytes, st>>>(x, H, W, nrm, wh, wl, wsi, bias, conv_out, slot);
    } else {
      if (attr16_set[0].need()) {
        CDR_CUDA(cudaFuncSetAttribute(stem_conv_f16x2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStem16SmemBytes));
        attr16_set[0].done();
      }
      stem_conv_f16x2_kernel<false><<<dim3(Ho / kTileH, n), 256, kStem16SmemBytes, st>>>(x, H, W, nrm, wh, wl, wsi, bias, conv_out, slot);
    }
    CDR_LAUNCH_OK("stem_conv_f16x2_kernel");
  } else if (is_u8) {
    if (attr_set[1].need()) {
      CDR_CUDA(cudaFuncSetAttribute(stem_conv_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStem32SmemBytes));
      attr_set[1].done();
    }
    stem_conv_f32_kernel<true><<<dim3(Ho / kTileH, n), 256, kStem32SmemBytes, st>>>(x, H, W, nrm, w, bias, conv_out, slot);
  } else {
    if (attr_set[0].need()) {
      CDR_CUDA(cudaFuncSetAttribute(stem_conv_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStem32SmemBytes));
      attr_set[0].done();
    }
    stem_conv_f32_kernel<false><<<dim3(Ho / kTileH, n), 256, kStem32SmemBytes, st>>>(x, H, W, nrm, w, bias, conv_out, slot);
  }
  if (stem_use_ffma()) CDR_LAUNCH_OK("stem_conv_f32_kernel");
  const long long total = (long long)n * (Ho / 2) * (Wo / 2) * (kStemCo / 8);
  maxpool3s2_f16p_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
      conv_out, Ho, Wo, kStemCo / 8, total, (__half*)pooled_hi, (__half*)pooled_lo, slot);
  CDR_LAUNCH_OK("maxpool3s2_f16p_kernel");
  return CDR_OK;
}

}  // namespace cdr
