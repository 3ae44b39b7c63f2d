// bf16 implicit-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same "tap-GEMM" arithmetic as gemm_ffma.cu — C[m,n] = sum_tap sum_k A[shift_tap(m),k] W[n][tap,k]
// over pixel-major bf16 activations — mapped to the Blackwell execution model:
//
//   * persistent CTAs (one per SM), 192 threads = 6 warps with fixed roles:
//       warp 0  TMA producer   one elected lane issues cp.async.bulk.tensor loads:
//                              A tile = 128 pixels x 64 channels.  For 1x1 convs a 2-D box of the
//                              (pixels, channels) matrix; for the transposed convs a 4-D box
//                              (64ch, W, rows, images) of the NHWC tensor whose start coordinate is
//                              shifted by the tap (dy,dx) — the TMA unit zero-fills out-of-range
//                              pixels, so the im2col never exists and borders cost nothing.
//                              B tile = BN output channels x 64 of the K-major packed weights.
//       warp 1  MMA issuer     one lane issues tcgen05.mma (M=128, N=BN, K=16, kind::f16, bf16 in,
//                              fp32 accumulate in TMEM); tcgen05.commit releases smem stages and
//                              publishes finished accumulators.  Also owns TMEM alloc/dealloc.
//       warps 2-5 epilogue     tcgen05.ld the 128 x BN fp32 accumulator (one TMEM lane = one pixel
//                              per thread), + folded-BN bias, ReLU, convert, store: pixel-major bf16
//                              rows, the transposed-conv phase scatter, or planar fp32 heat-maps.
//   * 128B-swizzled smem tiles shared by TMA and the UMMA descriptors, 4-6 stage mbarrier ring;
//   * two TMEM accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1.
//
// Tile order: consecutive CTAs take the 4 phases / N-tiles of the SAME 128-pixel block, so the
// shifted A tiles they share hit in L2; the packed weights (<= 16 MB) stay L2-resident.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "ptx.cuh"
#include "tc_api.h"

namespace cdr {

// ------------------------------------------------------------------------------------------
// tensor maps
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

// bf16 tensor, dims[0] innermost (contiguous); strides in elements for dims 1..rank-1
static int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box) {
  auto enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return CDR_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_elems[i - 1] * 2;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return CDR_ERR_CUDA;
  }
  return CDR_OK;
}

// ------------------------------------------------------------------------------------------
constexpr int kTcBM = 128;
constexpr int kTcBK = 64;                        // 64 bf16 = one 128-byte swizzle row
constexpr int kTcThreads = 192;
constexpr int kABytes = kTcBM * kTcBK * 2;       // 16 KB
constexpr int kSmemBudget = 200 * 1024;

template <int BN>
struct TcCfg {
  static constexpr int kBBytes = BN * kTcBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(kStages >= 2, "pipeline too shallow");
};

struct TcGemmParams {
  int M;                  // output rows (pixels) per group
  int n_img, H, W;        // pixel grid of the A operand
  int cin;                // K per tap
  int ntaps;              // 1 or 4
  int a4d;                // 1: A through the 4-D NHWC map (transposed conv), 0: 2-D (rows, channels)
  int box_rows, box_imgs; // 4-D box: W x box_rows x box_imgs pixels = 128
  int groups;             // phases (deconv) or independent problems stacked along rows
  int a_group_rows;       // 2-D A: row offset per group
  int b_group_rows;       // weight-row offset per group
  int n_tiles;            // N tiles of BN per group
  int n;                  // valid output channels
  const float* bias;      // (groups?, n_pad)
  int bias_group_stride;
  void* C;
  long long c_group_stride;
  int c_pitch, c_fill;
  int relu;
  int out_mode;           // kOutRows / kOutDeconv (bf16) or kOutPlanar (fp32)
  int num_tiles;
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
tap_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                   const __grid_constant__ CUtensorMap tmap_b, const TcGemmParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // S x 16 KB, each 1024-aligned
  uint8_t* smem_b = smem + (size_t)S * kABytes;            // S x BN*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * Cfg::kStageBytes);
  uint64_t* full = bars;                // [S]
  uint64_t* empty = bars + S;           // [S]
  uint64_t* tmem_full = bars + 2 * S;   // [2]
  uint64_t* tmem_empty = bars + 2 * S + 2;  // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);     // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int kb_per_tap = (p.cin + kTcBK - 1) / kTcBK;
  const int num_kb = p.ntaps * kb_per_tap;
  const int HW = p.H * p.W;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int g = (tile / p.n_tiles) % p.groups;
        const int m0 = (tile / (p.n_tiles * p.groups)) * kTcBM;
        const int py = g >> 1, px = g & 1;
        const int img0 = m0 / HW, y0 = (m0 - img0 * HW) / p.W;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t par = (it / S) & 1;
          ptx::mbar_wait(&empty[s], par ^ 1u);
          ptx::mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
          const int tap = kb / kb_per_tap;
          const int k0 = (kb - tap * kb_per_tap) * kTcBK;
          if (p.a4d) {
            const int dy = py - (tap >> 1), dx = px - (tap & 1);
            ptx::tma_load_4d(smem_a + (size_t)s * kABytes, &tmap_a, &full[s], k0, dx, y0 + dy, img0);
          } else {
            ptx::tma_load_2d(smem_a + (size_t)s * kABytes, &tmap_a, &full[s], k0, g * p.a_group_rows + m0);
          }
          ptx::tma_load_2d(smem_b + (size_t)s * Cfg::kBBytes, &tmap_b, &full[s], tap * p.cin + k0,
                           g * p.b_group_rows + n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=BN, M=128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(kTcBM >> 4) << 24);
      // smem matrix descriptor (K-major, SWIZZLE_128B): LBO=1, SBO=1024 B, version=1, layout=2
      constexpr uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                   ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const int acc = tl & 1;
        const uint32_t acc_par = (tl >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[acc], acc_par ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t par = (it / S) & 1;
          ptx::mbar_wait(&full[s], par);
          ptx::tc_fence_after();
          const uint64_t da = desc_hi | (uint64_t)((ptx::smem_u32(smem_a + (size_t)s * kABytes) >> 4) & 0x3FFF);
          const uint64_t db = desc_hi | (uint64_t)((ptx::smem_u32(smem_b + (size_t)s * Cfg::kBBytes) >> 4) & 0x3FFF);
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k)    // +32 bytes (= 2 x 16 B) per K=16 step inside the swizzle row
            ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          ptx::umma_commit(&empty[s]);             // smem stage reusable once these MMAs retire
        }
        ptx::umma_commit(&tmem_full[acc]);         // accumulator complete
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const int n_tile = tile % p.n_tiles;
      const int g = (tile / p.n_tiles) % p.groups;
      const int m0 = (tile / (p.n_tiles * p.groups)) * kTcBM;
      const int acc = tl & 1;
      const uint32_t acc_par = (tl >> 1) & 1;
      ptx::mbar_wait(&tmem_full[acc], acc_par);
      ptx::tc_fence_after();

      const int m = m0 + q * 32 + lane;            // this thread's pixel
      const bool row_ok = m < p.M;
      const int n0 = n_tile * BN;
      const float* __restrict__ bias = p.bias ? p.bias + (size_t)g * p.bias_group_stride : nullptr;
      size_t orow = (size_t)(row_ok ? m : 0);
      int img = 0, pix = 0;
      if (p.out_mode == kOutDeconv) {
        img = (int)(orow / HW);
        pix = (int)(orow - (size_t)img * HW);
        const int y = pix / p.W, x = pix - y * p.W;
        orow = ((size_t)img * 2 * p.H + 2 * y + (g >> 1)) * (size_t)(2 * p.W) + 2 * x + (g & 1);
      } else if (p.out_mode == kOutPlanar) {
        img = (int)(orow / HW);
        pix = (int)(orow - (size_t)img * HW);
      }
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)c, r);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]);
          if (bias) x += __ldg(bias + n0 + c + j);
          if (p.relu) x = fmaxf(x, 0.f);
          v[j] = (n0 + c + j < p.n) ? x : 0.f;
        }
        if (!row_ok) continue;
        if (p.out_mode == kOutPlanar) {
          float* __restrict__ C = reinterpret_cast<float*>(p.C);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c + j < p.n) C[((size_t)img * p.n + n0 + c + j) * HW + pix] = v[j];
        } else {
          __nv_bfloat16* __restrict__ C = reinterpret_cast<__nv_bfloat16*>(p.C) +
                                          (p.out_mode == kOutDeconv ? 0 : (size_t)g * p.c_group_stride) +
                                          orow * p.c_pitch + n0 + c;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (n0 + c + j8 * 8 < p.c_fill) {
              uint4 o;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j8 * 8 + 0], v[j8 * 8 + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j8 * 8 + 2], v[j8 * 8 + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j8 * 8 + 4], v[j8 * 8 + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j8 * 8 + 6], v[j8 * 8 + 7]);
              o.x = *reinterpret_cast<uint32_t*>(&h0);
              o.y = *reinterpret_cast<uint32_t*>(&h1);
              o.z = *reinterpret_cast<uint32_t*>(&h2);
              o.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(C + j8 * 8) = o;
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above) -> hand it back
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
    }
  }

  // ------------------------------------------------------------------ teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
struct TcLaunch {
  const __nv_bfloat16* A;   // activations
  int a_pitch;
  int n_img, H, W, cin, deconv, groups;
  long long a_rows_total;   // 2-D map: total rows addressable (groups stacked)
  const CUtensorMap* tmap_b;
  int b_group_rows, n, n_pad;
  const float* bias;
  int bias_group_stride;
  void* C;
  long long c_group_stride;
  int c_pitch, c_fill, relu, out_mode;
};

template <int BN>
static int launch_tc(const TcLaunch& l, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CDR_CUDA(cudaFuncSetAttribute(tap_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)Cfg::kSmemBytes));
    attr_set = true;
  }
  TcGemmParams p{};
  const int HW = l.H * l.W;
  p.M = l.n_img * HW;
  p.n_img = l.n_img; p.H = l.H; p.W = l.W; p.cin = l.cin;
  p.ntaps = l.deconv ? 4 : 1;
  p.a4d = l.deconv;
  p.groups = l.groups;
  p.a_group_rows = l.deconv ? 0 : p.M;
  p.b_group_rows = l.b_group_rows;
  p.n_tiles = l.n_pad / BN;
  p.n = l.n;
  p.bias = l.bias; p.bias_group_stride = l.bias_group_stride;
  p.C = l.C; p.c_group_stride = l.c_group_stride; p.c_pitch = l.c_pitch; p.c_fill = l.c_fill;
  p.relu = l.relu; p.out_mode = l.out_mode;
  const int m_tiles = ceil_div(p.M, kTcBM);
  p.num_tiles = m_tiles * p.groups * p.n_tiles;
  CDR_CHECK_ARG(l.n_pad % BN == 0 && l.a_pitch % 8 == 0 && ((uintptr_t)l.A & 15) == 0,
                "tap_gemm_tc: n_pad %% BN, a_pitch %% 8 or A alignment violated");
  CDR_CHECK_ARG(l.out_mode == kOutPlanar || (l.c_pitch % 8 == 0 && l.c_fill % 8 == 0),
                "tap_gemm_tc: bf16 output needs c_pitch %% 8 == 0 and c_fill %% 8 == 0");

  CUtensorMap tmap_a;
  if (l.deconv) {
    CDR_CHECK_ARG(l.W <= kTcBM && kTcBM % l.W == 0, "tap_gemm_tc: W=%d must divide 128", l.W);
    int rows = kTcBM / l.W;
    if (rows > l.H) rows = l.H;
    const int imgs = kTcBM / (l.W * rows);
    CDR_CHECK_ARG(l.H % rows == 0 && l.cin % kTcBK == 0, "tap_gemm_tc: unsupported deconv geometry");
    p.box_rows = rows; p.box_imgs = imgs;
    const uint64_t dims[4] = {(uint64_t)l.cin, (uint64_t)l.W, (uint64_t)l.H, (uint64_t)l.n_img};
    const uint64_t strides[3] = {(uint64_t)l.a_pitch, (uint64_t)l.a_pitch * l.W, (uint64_t)l.a_pitch * HW};
    const uint32_t box[4] = {(uint32_t)kTcBK, (uint32_t)l.W, (uint32_t)rows, (uint32_t)imgs};
    if (int rc = make_tmap_bf16(&tmap_a, l.A, 4, dims, strides, box)) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)l.cin, (uint64_t)l.a_rows_total};
    const uint64_t strides[1] = {(uint64_t)l.a_pitch};
    const uint32_t box[2] = {(uint32_t)kTcBK, (uint32_t)kTcBM};
    if (int rc = make_tmap_bf16(&tmap_a, l.A, 2, dims, strides, box)) return rc;
  }
  int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  tap_gemm_tc_kernel<BN><<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(tmap_a, *l.tmap_b, p);
  CDR_LAUNCH_OK("tap_gemm_tc_kernel");
  return CDR_OK;
}

// ------------------------------------------------------------------------------------------
// bf16 weight packing (K-major B operands) — BN folded as in pack.cu
constexpr double kBnEpsTc = 1e-5;
__device__ __forceinline__ double tc_bn_scale(const CdrConvBn& s, int co) {
  return s.bn_weight ? (double)s.bn_weight[co] / sqrt((double)s.bn_var[co] + kBnEpsTc) : 1.0;
}
__device__ __forceinline__ float tc_folded_bias(const CdrConvBn& s, int co) {
  const double b = s.bias ? (double)s.bias[co] : 0.0;
  if (!s.bn_weight) return (float)b;
  return (float)((b - (double)s.bn_mean[co]) * tc_bn_scale(s, co) + (double)s.bn_bias[co]);
}
// (Cout,Cin) -> [n_pad][k_pitch] bf16
__global__ void pack_conv1x1_bf16_kernel(CdrConvBn s, int cout, int cin, int k_pitch, int n_pad,
                                         __nv_bfloat16* __restrict__ w_out, float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) bias_out[idx] = idx < cout ? tc_folded_bias(s, (int)idx) : 0.f;
  if (idx >= (long long)n_pad * k_pitch) return;
  const int n = (int)(idx / k_pitch), k = (int)(idx % k_pitch);
  float v = 0.f;
  if (n < cout && k < cin) v = (float)((double)s.weight[(size_t)n * cin + k] * tc_bn_scale(s, n));
  w_out[idx] = __float2bfloat16_rn(v);
}
// (Cin,Cout,4,4) -> [phase][n_pad][tap*Cin + ci] bf16
__global__ void pack_deconv_bf16_kernel(CdrConvBn s, int cin, int cout, int n_pad,
                                        __nv_bfloat16* __restrict__ w_out, float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) bias_out[idx] = idx < cout ? tc_folded_bias(s, (int)idx) : 0.f;
  const long long K = 4LL * cin, total = 4LL * n_pad * K;
  if (idx >= total) return;
  const int kk = (int)(idx % K);
  long long r = idx / K;
  const int n = (int)(r % n_pad), phase = (int)(r / n_pad);
  const int tap = kk / cin, ci = kk - tap * cin;
  const int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
  const int ky = 1 - py + 2 * ty, kx = 1 - px + 2 * tx;
  float v = 0.f;
  if (n < cout) v = (float)((double)s.weight[(((size_t)ci * cout + n) * 4 + ky) * 4 + kx] * tc_bn_scale(s, n));
  w_out[idx] = __float2bfloat16_rn(v);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                   long long rows, int in_pitch, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = (int)(idx - r * cols);
  out[idx] = __bfloat162float(in[r * in_pitch + c]);
}

// ------------------------------------------------------------------------------------------
struct TcMaps {  // weight tensor maps live with the weights
  CUtensorMap cf1, cf2a, cf2b, out, dc[3], fin;
};
static constexpr int kTcCf1NPad = 384, kTcCf2NPad = 512;   // BN=128 tiles
static const int kTcDcCin[3] = {kFeatC, kDecC, kDecC};

template <typename T>
static T* bump(uint8_t*& p, size_t count) {
  T* r = (T*)p;
  p += round_up<size_t>(count * sizeof(T), 1024);
  return r;
}

static size_t plan_tc_weights(TcWeights& w, uint8_t* base) {
  uint8_t* p = base;
  if (w.has_fusion) {
    w.w_cf1 = bump<__nv_bfloat16>(p, (size_t)kTcCf1NPad * kFeatC);
    w.w_cf2a = bump<__nv_bfloat16>(p, (size_t)kTcCf2NPad * 2 * kHid2);
    w.w_cf2b = bump<__nv_bfloat16>(p, (size_t)kTcCf2NPad * kHid2);
    w.w_out = bump<__nv_bfloat16>(p, (size_t)2 * kFeatC * kHid1Pad);
    w.b_cf1 = bump<float>(p, kTcCf1NPad);
    w.b_cf2a = bump<float>(p, kTcCf2NPad);
    w.b_cf2b = bump<float>(p, kTcCf2NPad);
    w.b_out = bump<float>(p, (size_t)2 * kFeatC);
  }
  for (int i = 0; i < 3; ++i) {
    w.w_dc[i] = bump<__nv_bfloat16>(p, (size_t)4 * kDecC * 4 * kTcDcCin[i]);
    w.b_dc[i] = bump<float>(p, kDecC);
  }
  w.w_fin = bump<__nv_bfloat16>(p, (size_t)w.fin_npad * kDecC);
  w.b_fin = bump<float>(p, w.fin_npad);
  return (size_t)(p - base);
}

static int weight_map(CUtensorMap* m, const __nv_bfloat16* w, int rows, int k, int k_pitch, int bn) {
  const uint64_t dims[2] = {(uint64_t)k, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)k_pitch};
  const uint32_t box[2] = {(uint32_t)kTcBK, (uint32_t)bn};
  return make_tmap_bf16(m, w, 2, dims, strides, box);
}

int tc_weights_create(const CdrWeightPtrs& src, TcWeights& w, cudaStream_t st) {
  w.joints = src.num_joints;
  w.has_fusion = src.has_fusion;
  w.fin_npad = round_up(src.num_joints, 32);
  const size_t bytes = plan_tc_weights(w, nullptr);
  CDR_CUDA(cudaMalloc(&w.pool, bytes));
  plan_tc_weights(w, (uint8_t*)w.pool);
  auto pack1 = [&](const CdrConvBn& s, int cout, int cin, int k_pitch, int n_pad, __nv_bfloat16* wo,
                   float* bo) -> int {
    const long long total = (long long)n_pad * k_pitch;
    pack_conv1x1_bf16_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(s, cout, cin, k_pitch, n_pad, wo, bo);
    CDR_LAUNCH_OK("pack_conv1x1_bf16_kernel");
    return CDR_OK;
  };
  int rc;
  TcMaps* maps = new TcMaps();
  w.maps = maps;
  if (w.has_fusion) {
    if ((rc = pack1(src.cf_conv1, kHid1, kFeatC, kFeatC, kTcCf1NPad, w.w_cf1, w.b_cf1))) return rc;
    if ((rc = pack1(src.cf_conv2a, kHid2, 2 * kHid2, 2 * kHid2, kTcCf2NPad, w.w_cf2a, w.b_cf2a))) return rc;
    if ((rc = pack1(src.cf_conv2b, kHid2, kHid2, kHid2, kTcCf2NPad, w.w_cf2b, w.b_cf2b))) return rc;
    for (int v = 0; v < 2; ++v)
      if ((rc = pack1(src.cf_out[v], kFeatC, kHid1, kHid1Pad, kFeatC, w.w_out + (size_t)v * kFeatC * kHid1Pad,
                      w.b_out + (size_t)v * kFeatC)))
        return rc;
    if ((rc = weight_map(&maps->cf1, w.w_cf1, kTcCf1NPad, kFeatC, kFeatC, 128))) return rc;
    if ((rc = weight_map(&maps->cf2a, w.w_cf2a, kTcCf2NPad, 2 * kHid2, 2 * kHid2, 128))) return rc;
    if ((rc = weight_map(&maps->cf2b, w.w_cf2b, kTcCf2NPad, kHid2, kHid2, 128))) return rc;
    if ((rc = weight_map(&maps->out, w.w_out, 2 * kFeatC, kHid1, kHid1Pad, 256))) return rc;
  }
  for (int i = 0; i < 3; ++i) {
    const long long total = 4LL * kDecC * 4 * kTcDcCin[i];
    pack_deconv_bf16_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, st>>>(
        src.deconv[i], kTcDcCin[i], kDecC, kDecC, w.w_dc[i], w.b_dc[i]);
    CDR_LAUNCH_OK("pack_deconv_bf16_kernel");
    if ((rc = weight_map(&maps->dc[i], w.w_dc[i], 4 * kDecC, 4 * kTcDcCin[i], 4 * kTcDcCin[i], 256))) return rc;
  }
  if ((rc = pack1(src.final_layer, w.joints, kDecC, kDecC, w.fin_npad, w.w_fin, w.b_fin))) return rc;
  if ((rc = weight_map(&maps->fin, w.w_fin, w.fin_npad, kDecC, kDecC, 32))) return rc;
  return CDR_OK;
}

void tc_weights_destroy(TcWeights& w) {
  if (w.pool) cudaFree(w.pool);
  w.pool = nullptr;
  delete (TcMaps*)w.maps;
  w.maps = nullptr;
}

// ------------------------------------------------------------------------------------------
struct TcHeadWs {
  float* pinv;
  __nv_bfloat16 *x0, *y1, *z, *f1, *f2, *g, *x1, *d1, *d2, *d3;
  float* hm;
  size_t bytes;
};
static TcHeadWs plan_tc_head(void* base, int B, int J) {
  uint8_t* p = (uint8_t*)base;
  const size_t N = 2 * (size_t)B;
  TcHeadWs w;
  w.pinv = bump<float>(p, N * 12);
  w.x0 = bump<__nv_bfloat16>(p, N * kFeatHW * kFeatC);
  w.y1 = bump<__nv_bfloat16>(p, N * kFeatHW * kHid1Pad);
  w.z = bump<__nv_bfloat16>(p, (size_t)B * kFeatHW * 2 * kHid2);
  w.f1 = bump<__nv_bfloat16>(p, (size_t)B * kFeatHW * kHid2);
  w.f2 = bump<__nv_bfloat16>(p, (size_t)B * kFeatHW * kHid2);
  w.g = bump<__nv_bfloat16>(p, N * kFeatHW * kHid1Pad);
  w.x1 = bump<__nv_bfloat16>(p, N * kFeatHW * kFeatC);
  w.d1 = bump<__nv_bfloat16>(p, N * 256 * kDecC);
  w.d2 = bump<__nv_bfloat16>(p, N * 1024 * kDecC);
  w.d3 = bump<__nv_bfloat16>(p, N * 4096 * kDecC);
  w.hm = bump<float>(p, N * J * 4096);
  w.bytes = (size_t)(p - (uint8_t*)base);
  return w;
}
struct TcDecWs {
  __nv_bfloat16 *x1, *d1, *d2, *d3;
  size_t bytes;
};
static TcDecWs plan_tc_dec(void* base, int N) {
  uint8_t* p = (uint8_t*)base;
  TcDecWs w;
  w.x1 = bump<__nv_bfloat16>(p, (size_t)N * kFeatHW * kFeatC);
  w.d1 = bump<__nv_bfloat16>(p, (size_t)N * 256 * kDecC);
  w.d2 = bump<__nv_bfloat16>(p, (size_t)N * 1024 * kDecC);
  w.d3 = bump<__nv_bfloat16>(p, (size_t)N * 4096 * kDecC);
  w.bytes = (size_t)(p - (uint8_t*)base);
  return w;
}

int tc_head_workspace_bytes(const TcWeights& w, int batch, size_t* bytes) {
  *bytes = plan_tc_head(nullptr, batch, w.joints).bytes;
  return CDR_OK;
}
int tc_decoder_workspace_bytes(const TcWeights&, int n_images, size_t* bytes) {
  *bytes = plan_tc_dec(nullptr, n_images).bytes;
  return CDR_OK;
}

static int tc_decoder(const TcWeights& w, const __nv_bfloat16* x1, int N, __nv_bfloat16* d1,
                      __nv_bfloat16* d2, __nv_bfloat16* d3, float* heat, cudaStream_t st) {
  const TcMaps* maps = (const TcMaps*)w.maps;
  static const char* const kDcName[3] = {"deconv1", "deconv2", "deconv3"};
  const __nv_bfloat16* in = x1;
  __nv_bfloat16* outs[3] = {d1, d2, d3};
  int side = 8;
  for (int i = 0; i < 3; ++i) {
    set_stage(kDcName[i]);
    TcLaunch l{};
    l.A = in; l.a_pitch = kTcDcCin[i]; l.n_img = N; l.H = l.W = side; l.cin = kTcDcCin[i];
    l.deconv = 1; l.groups = 4; l.tmap_b = &maps->dc[i]; l.b_group_rows = kDecC; l.n = kDecC; l.n_pad = kDecC;
    l.bias = w.b_dc[i]; l.bias_group_stride = 0;
    l.C = outs[i]; l.c_pitch = kDecC; l.c_fill = kDecC; l.relu = 1; l.out_mode = kOutDeconv;
    if (int rc = launch_tc<256>(l, st)) return rc;
    in = outs[i];
    side *= 2;
  }
  set_stage("final_1x1");
  TcLaunch l{};
  l.A = d3; l.a_pitch = kDecC; l.n_img = N; l.H = l.W = kHeat; l.cin = kDecC; l.groups = 1;
  l.a_rows_total = (long long)N * 4096;
  l.tmap_b = &maps->fin; l.n = w.joints; l.n_pad = w.fin_npad; l.bias = w.b_fin;
  l.C = heat; l.relu = 0; l.out_mode = kOutPlanar;
  const int rc = launch_tc<32>(l, st);
  set_stage(nullptr);
  return rc;
}

static int tap_to_f32(float* dst, const __nv_bfloat16* src, long long rows, int pitch, int cols, cudaStream_t st) {
  if (!dst) return CDR_OK;
  bf16_to_f32_kernel<<<(unsigned)ceil_div<long long>(rows * cols, 256), 256, 0, st>>>(src, dst, rows, pitch, cols);
  CDR_LAUNCH_OK("bf16_to_f32_kernel");
  return CDR_OK;
}

int tc_head_forward(const TcWeights& w, const float* feat_l, const float* feat_r, const float* P_l,
                    const float* P_r, const float* pinv_l, const float* pinv_r, double pinv_rtol,
                    int batch, float scale, float* kp2d_l, float* kp2d_r, float* xyz,
                    const CdrHeadTaps* taps, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const TcMaps* maps = (const TcMaps*)w.maps;
  const int B = batch, N = 2 * batch, J = w.joints;
  TcHeadWs ws = plan_tc_head(workspace, B, J);
  if (ws.bytes > workspace_bytes) {
    set_error("cdr_head_forward: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return CDR_ERR_WORKSPACE;
  }
  int rc;
  set_stage("pinv");
  const float* pinv[2] = {pinv_l, pinv_r};
  if (!pinv_l) {
    if ((rc = cdr_pinv(P_l, B, pinv_rtol, ws.pinv, st))) return rc;
    if ((rc = cdr_pinv(P_r, B, pinv_rtol, ws.pinv + (size_t)B * 12, st))) return rc;
    pinv[0] = ws.pinv;
    pinv[1] = ws.pinv + (size_t)B * 12;
  }
  set_stage("nchw_to_rows");
  if ((rc = launch_nchw_to_rows_bf16(feat_l, B, kFeatC, kFeatHW, ws.x0, kFeatC, st))) return rc;
  if ((rc = launch_nchw_to_rows_bf16(feat_r, B, kFeatC, kFeatHW, ws.x0 + (size_t)B * kFeatHW * kFeatC, kFeatC, st))) return rc;
  set_stage("cf_conv1");
  {
    TcLaunch l{};
    l.A = ws.x0; l.a_pitch = kFeatC; l.n_img = N; l.H = l.W = 8; l.cin = kFeatC; l.groups = 1;
    l.a_rows_total = (long long)N * kFeatHW;
    l.tmap_b = &maps->cf1; l.n = kHid1; l.n_pad = kTcCf1NPad; l.bias = w.b_cf1;
    l.C = ws.y1; l.c_pitch = kHid1Pad; l.c_fill = kHid1Pad; l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc<128>(l, st))) return rc;
  }
  set_stage("ftl_inv");
  for (int v = 0; v < 2; ++v)
    if ((rc = launch_ftl<__nv_bfloat16>(ws.y1 + (size_t)v * B * kFeatHW * kHid1Pad, kHid1Pad, pinv[v], 4, 3,
                                        kFtlBlk, B, kFeatHW, ws.z + v * kHid2, 2 * kHid2, kHid2, st)))
      return rc;
  set_stage("cf_conv2");
  {
    TcLaunch l{};
    l.A = ws.z; l.a_pitch = 2 * kHid2; l.n_img = B; l.H = l.W = 8; l.cin = 2 * kHid2; l.groups = 1;
    l.a_rows_total = (long long)B * kFeatHW;
    l.tmap_b = &maps->cf2a; l.n = kHid2; l.n_pad = kTcCf2NPad; l.bias = w.b_cf2a;
    l.C = ws.f1; l.c_pitch = kHid2; l.c_fill = kHid2; l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc<128>(l, st))) return rc;
    l.A = ws.f1; l.a_pitch = kHid2; l.cin = kHid2; l.tmap_b = &maps->cf2b; l.bias = w.b_cf2b; l.C = ws.f2;
    if ((rc = launch_tc<128>(l, st))) return rc;
  }
  set_stage("ftl_fwd");
  const float* Pv[2] = {P_l, P_r};
  for (int v = 0; v < 2; ++v)
    if ((rc = launch_ftl<__nv_bfloat16>(ws.f2, kHid2, Pv[v], 3, 4, kFtlBlk, B, kFeatHW,
                                        ws.g + (size_t)v * B * kFeatHW * kHid1Pad, kHid1Pad, kHid1Pad, st)))
      return rc;
  set_stage("cf_out");
  {
    TcLaunch l{};
    l.A = ws.g; l.a_pitch = kHid1Pad; l.n_img = B; l.H = l.W = 8; l.cin = kHid1; l.groups = 2;
    l.a_rows_total = (long long)N * kFeatHW;
    l.tmap_b = &maps->out; l.b_group_rows = kFeatC; l.n = kFeatC; l.n_pad = kFeatC;
    l.bias = w.b_out; l.bias_group_stride = kFeatC;
    l.C = ws.x1; l.c_group_stride = (long long)B * kFeatHW * kFeatC; l.c_pitch = kFeatC; l.c_fill = kFeatC;
    l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc<256>(l, st))) return rc;
  }
  if ((rc = tc_decoder(w, ws.x1, N, ws.d1, ws.d2, ws.d3, ws.hm, st))) return rc;
  set_stage("softargmax_dlt");
  if ((rc = cdr_softargmax_dlt(ws.hm, ws.hm + (size_t)B * J * 4096, 0, P_l, P_r, B, J, kHeat, kHeat, scale,
                               kp2d_l, kp2d_r, xyz, nullptr, nullptr, nullptr, nullptr, nullptr, st)))
    return rc;
  set_stage(nullptr);
  if (taps) {
    if (taps->pinv) {
      CDR_CUDA(cudaMemcpyAsync(taps->pinv, pinv[0], (size_t)B * 48, cudaMemcpyDeviceToDevice, st));
      CDR_CUDA(cudaMemcpyAsync(taps->pinv + (size_t)B * 12, pinv[1], (size_t)B * 48, cudaMemcpyDeviceToDevice, st));
    }
    if ((rc = tap_to_f32(taps->cf_cat, ws.z, (long long)B * kFeatHW, 2 * kHid2, 2 * kHid2, st))) return rc;
    if ((rc = tap_to_f32(taps->cf_f, ws.f2, (long long)B * kFeatHW, kHid2, kHid2, st))) return rc;
    if ((rc = tap_to_f32(taps->f_out, ws.x1, (long long)N * kFeatHW, kFeatC, kFeatC, st))) return rc;
    if (taps->heatmaps)
      CDR_CUDA(cudaMemcpyAsync(taps->heatmaps, ws.hm, (size_t)N * J * 4096 * 4, cudaMemcpyDeviceToDevice, st));
  }
  return CDR_OK;
}

int tc_decoder_forward(const TcWeights& w, const float* feat, int n_images, float* heatmaps,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TcDecWs ws = plan_tc_dec(workspace, n_images);
  if (ws.bytes > workspace_bytes) {
    set_error("cdr_decoder_forward: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return CDR_ERR_WORKSPACE;
  }
  set_stage("nchw_to_rows");
  if (int rc = launch_nchw_to_rows_bf16(feat, n_images, kFeatC, kFeatHW, ws.x1, kFeatC, st)) return rc;
  return tc_decoder(w, ws.x1, n_images, ws.d1, ws.d2, ws.d3, heatmaps, st);
}

}  // namespace cdr
