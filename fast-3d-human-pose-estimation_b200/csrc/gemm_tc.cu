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

// dims[0] innermost (contiguous); strides in elements for dims 1..rank-1; elem_bytes 2 (bf16) or 4 (fp32)
static int make_tmap(CUtensorMap* map, const void* base, int elem_bytes, int rank, const uint64_t* dims,
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
    if (i > 0) gstr[i - 1] = strides_elems[i - 1] * (uint64_t)elem_bytes;
  }
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   (cuuint32_t)rank, const_cast<void*>(base),
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
constexpr int kTcThreads = 192;
constexpr int kABytes = kTcBM * 128;             // 128 rows x one 128-byte swizzle row = 16 KB
constexpr int kSmemBudget = 200 * 1024;

// Operand kinds.
//   kKindBF16  : bf16 operands, kind::f16, one MMA per K=16 step.
//   kKindTF32X3: fp32 accuracy on the tensor cores.  Every fp32 operand is stored as two fp32
//                planes hi = rn_tf32(x), lo = x - hi (common.cuh: split_tf32); x*y ~= hi*hi' + hi*lo' + lo*hi' with kind::tf32 MMAs accumulating in
//                fp32 (dropped lo*lo' term ~2^-22 relative, random sign).
enum TcKind { kKindBF16 = 0, kKindTF32X3 = 1 };
template <int KIND> struct KindTraits;
template <> struct KindTraits<kKindBF16>   { static constexpr int kElem = 2, kBK = 64, kPlanes = 1; };
template <> struct KindTraits<kKindTF32X3> { static constexpr int kElem = 4, kBK = 32, kPlanes = 2; };

template <int BN, int KIND>
struct TcCfg {
  static constexpr int kBK = KindTraits<KIND>::kBK;
  static constexpr int kPlanes = KindTraits<KIND>::kPlanes;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kAccBufs = KIND == kKindTF32X3 ? 4 : 2;   // see the TMEM column map in the kernel
  static constexpr int kTmemCols = (kAccBufs * BN <= 32) ? 32 : (kAccBufs * BN <= 64) ? 64
                                   : (kAccBufs * BN <= 128) ? 128 : (kAccBufs * BN <= 256) ? 256 : 512;
  static_assert(kAccBufs * BN <= 512, "TMEM has 512 columns");
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers*/;
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
  void* C;                // bf16 rows, or the fp32 hi plane (tf32x3), or planar fp32 heat-maps
  void* C_lo;             // tf32x3: the lo plane
  long long c_group_stride;
  int c_pitch, c_fill;
  int relu;
  int out_mode;           // kOutRows / kOutDeconv or kOutPlanar (fp32)
  int num_tiles;
};

// Epilogue for one 32-column slab of a finished row: + bias, ReLU, mask, convert, store.
template <int KIND>
__device__ __forceinline__ void store_slab(const TcGemmParams& p, const float (&acc)[32], int n0c, int g, bool row_ok,
                                           size_t orow, int img, int pix, int HW, const float* __restrict__ bias) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = acc[j];
    if (bias) x += __ldg(bias + n0c + j);
    if (p.relu) x = fmaxf(x, 0.f);
    v[j] = (n0c + j < p.n) ? x : 0.f;
  }
  if (!row_ok) return;
  if (p.out_mode == kOutPlanar) {
    float* __restrict__ C = reinterpret_cast<float*>(p.C);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0c + j < p.n) C[((size_t)img * p.n + n0c + j) * HW + pix] = v[j];
    return;
  }
  const size_t off = (p.out_mode == kOutDeconv ? 0 : (size_t)g * p.c_group_stride) + orow * p.c_pitch + n0c;
  if constexpr (KIND == kKindTF32X3) {
    float* __restrict__ Ch = reinterpret_cast<float*>(p.C) + off;
    float* __restrict__ Cl = reinterpret_cast<float*>(p.C_lo) + off;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      if (n0c + j4 * 4 < p.c_fill) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[j4 * 4 + e], hi[e], lo[e]);
        *reinterpret_cast<float4*>(Ch + j4 * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(Cl + j4 * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  } else {
    __nv_bfloat16* __restrict__ C = reinterpret_cast<__nv_bfloat16*>(p.C) + off;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      if (n0c + j8 * 8 < p.c_fill) {
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

// K-blocks accumulated inside TMEM before the 3xTF32 main term is drained into fp32 registers.
// The tensor core adds into its fp32 accumulator with round-toward-zero; over a long K chain
// of same-signed partial sums that bias grows linearly (measured: 1e-5 relative at K=2048, 40x
// worse than FFMA).  Chains of 4 K-blocks (16 MMAs) keep it below 1e-6; the cross-chunk sum is
// done by the epilogue warps in registers with round-to-nearest.
constexpr int kSplitChunk = 4;

template <int BN, int KIND>
__global__ void __launch_bounds__(kTcThreads, 1)
tap_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
                   const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b_lo,
                   const TcGemmParams p) {
  using Cfg = TcCfg<BN, KIND>;
  constexpr int S = Cfg::kStages;
  constexpr int kTcBK = Cfg::kBK;
  constexpr bool kSplit = KIND == kKindTF32X3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage s: [A (16 KB) | A_lo (tf32x3) | B (BN*128 B) | B_lo (tf32x3)], every piece 1024-aligned
  auto stage_a = [&](int s, int plane) { return smem + (size_t)s * Cfg::kStageBytes + (size_t)plane * kABytes; };
  auto stage_b = [&](int s, int plane) {
    return smem + (size_t)s * Cfg::kStageBytes + (size_t)Cfg::kPlanes * kABytes + (size_t)plane * Cfg::kBBytes;
  };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * Cfg::kStageBytes);
  uint64_t* full = bars;                      // [S]  TMA -> MMA
  uint64_t* empty = bars + S;                 // [S]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * S;         // [2]  finished tile accumulator (bf16) / correction accumulator (tf32x3)
  uint64_t* tmem_empty = bars + 2 * S + 2;    // [2]
  uint64_t* chunk_full = bars + 2 * S + 4;    // [2]  tf32x3: main-term chunk accumulator
  uint64_t* chunk_empty = bars + 2 * S + 6;   // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (kSplit) {
      ptx::prefetch_tmap(&tmap_a_lo);
      ptx::prefetch_tmap(&tmap_b_lo);
    }
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 4);     // one arrive per epilogue warp
      ptx::mbar_init(&chunk_full[a], 1);
      ptx::mbar_init(&chunk_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // TMEM columns: bf16  : [0,BN) [BN,2BN)               two tile accumulators
  //               tf32x3: [0,BN) [BN,2BN)               two main-term chunk accumulators
  //                       [2BN,3BN) [3BN,4BN)           two correction-term tile accumulators

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
            ptx::tma_load_4d(stage_a(s, 0), &tmap_a, &full[s], k0, dx, y0 + dy, img0);
            if (kSplit) ptx::tma_load_4d(stage_a(s, 1), &tmap_a_lo, &full[s], k0, dx, y0 + dy, img0);
          } else {
            ptx::tma_load_2d(stage_a(s, 0), &tmap_a, &full[s], k0, g * p.a_group_rows + m0);
            if (kSplit) ptx::tma_load_2d(stage_a(s, 1), &tmap_a_lo, &full[s], k0, g * p.a_group_rows + m0);
          }
          ptx::tma_load_2d(stage_b(s, 0), &tmap_b, &full[s], tap * p.cin + k0, g * p.b_group_rows + n_tile * BN);
          if (kSplit)
            ptx::tma_load_2d(stage_b(s, 1), &tmap_b_lo, &full[s], tap * p.cin + k0, g * p.b_group_rows + n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A/B format (1 = bf16, 2 = tf32), both K-major, N=BN, M=128
      constexpr uint32_t fmt = kSplit ? 2u : 1u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(kTcBM >> 4) << 24);
      // smem matrix descriptor (K-major, SWIZZLE_128B): LBO=1, SBO=1024 B, version=1, layout=2
      constexpr uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
                                   ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      auto desc = [&](const uint8_t* ptr) { return desc_hi | (uint64_t)((ptx::smem_u32(ptr) >> 4) & 0x3FFF); };
      uint32_t it = 0, tl = 0, ch = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const int acc = tl & 1;
        ptx::mbar_wait(&tmem_empty[acc], ((tl >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        if constexpr (!kSplit) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % S;
            ptx::mbar_wait(&full[s], (it / S) & 1);
            ptx::tc_fence_after();
            const uint64_t da = desc(stage_a(s, 0)), db = desc(stage_b(s, 0));
#pragma unroll
            for (int k = 0; k < 4; ++k)    // +32 bytes (= 2 x 16 B) per K step (16 bf16) inside the swizzle row
              ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            ptx::umma_commit(&empty[s]);           // smem stage reusable once these MMAs retire
          }
        } else {
          const uint32_t d_corr = tmem_base + (uint32_t)((2 + acc) * BN);
          for (int kb0 = 0; kb0 < num_kb; kb0 += kSplitChunk, ++ch) {
            const int buf = ch & 1;
            ptx::mbar_wait(&chunk_empty[buf], ((ch >> 1) & 1) ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_main = tmem_base + (uint32_t)(buf * BN);
            const int kb1 = kb0 + kSplitChunk < num_kb ? kb0 + kSplitChunk : num_kb;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
              const int s = it % S;
              ptx::mbar_wait(&full[s], (it / S) & 1);
              ptx::tc_fence_after();
              const uint64_t da = desc(stage_a(s, 0)), dal = desc(stage_a(s, 1));
              const uint64_t db = desc(stage_b(s, 0)), dbl = desc(stage_b(s, 1));
#pragma unroll
              for (int k = 0; k < 4; ++k) {        // K step = 8 tf32 = 32 bytes
                const uint64_t o = (uint64_t)(2 * k);
                ptx::umma_tf32(d_corr, dal + o, db + o, idesc, (kb | k) != 0);   // lo*hi  } whole-tile chain: these
                ptx::umma_tf32(d_corr, da + o, dbl + o, idesc, 1u);              // hi*lo  } terms are 2^-11 of the main one
                ptx::umma_tf32(d_main, da + o, db + o, idesc, (kb > kb0 || k > 0));  // hi*hi, short chain
              }
              ptx::umma_commit(&empty[s]);
            }
            ptx::umma_commit(&chunk_full[buf]);    // main-term chunk ready to be drained
          }
        }
        ptx::umma_commit(&tmem_full[acc]);         // tile (bf16) / correction accumulator (tf32x3) complete
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    uint32_t tl = 0, ch = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const int n_tile = tile % p.n_tiles;
      const int g = (tile / p.n_tiles) % p.groups;
      const int m0 = (tile / (p.n_tiles * p.groups)) * kTcBM;
      const int acc = tl & 1;
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
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);

      if constexpr (!kSplit) {
        ptx::mbar_wait(&tmem_full[acc], (tl >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(acc * BN + c), r);
          ptx::tmem_ld_wait();
          float a32[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a32[j] = __uint_as_float(r[j]);
          store_slab<KIND>(p, a32, n0 + c, g, row_ok, orow, img, pix, HW, bias);
        }
        // all TMEM reads of this accumulator are complete (wait::ld above) -> hand it back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      } else {
        // running fp32 sum of this thread's output row, round-to-nearest
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.f;
        for (int kb0 = 0; kb0 < num_kb; kb0 += kSplitChunk, ++ch) {
          const int buf = ch & 1;
          ptx::mbar_wait(&chunk_full[buf], (ch >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(buf * BN + c), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c + j] += __uint_as_float(r[j]);
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&chunk_empty[buf]);
        }
        ptx::mbar_wait(&tmem_full[acc], (tl >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((2 + acc) * BN + c), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c + j] += __uint_as_float(r[j]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
#pragma unroll
        for (int c = 0; c < BN; c += 32) {
          float a32[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a32[j] = sum[c + j];
          store_slab<KIND>(p, a32, n0 + c, g, row_ok, orow, img, pix, HW, bias);
        }
      }
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
// host side: packed layers, launches, orchestration
struct TcLayer {            // one packed conv: K-major B operand (1 plane bf16 / 2 planes fp32) + bias
  void* w[2] = {nullptr, nullptr};
  float* bias = nullptr;
  CUtensorMap map[2];
  int rows = 0, k = 0, k_pitch = 0, bn = 0, n_pad = 0;
};
struct TcPack {
  int kind = kKindBF16;
  TcLayer cf1, cf2a, cf2b, out, dc[3], fin;
};
struct Act {                // an activation buffer: 1 plane (bf16) or hi/lo planes (fp32)
  void* p[2] = {nullptr, nullptr};
};
static inline Act act_offset(const Act& a, size_t elems, int elem_bytes) {
  Act r;
  r.p[0] = a.p[0] ? (uint8_t*)a.p[0] + elems * elem_bytes : nullptr;
  r.p[1] = a.p[1] ? (uint8_t*)a.p[1] + elems * elem_bytes : nullptr;
  return r;
}

struct TcLaunch {
  Act A;                    // activations
  int a_pitch;
  int n_img, H, W, cin, deconv, groups;
  long long a_rows_total;   // 2-D map: total rows addressable (groups stacked)
  const TcLayer* layer;
  int b_group_rows, n;
  int bias_group_stride;
  Act C;                    // output planes (or planar fp32 heat-maps in C.p[0])
  long long c_group_stride;
  int c_pitch, c_fill, relu, out_mode;
};

template <int BN, int KIND>
static int launch_tc_t(const TcLaunch& l, cudaStream_t st) {
  using Cfg = TcCfg<BN, KIND>;
  constexpr int kElem = KindTraits<KIND>::kElem;
  constexpr int kBK = Cfg::kBK;
  static bool attr_set = false;
  if (!attr_set) {
    CDR_CUDA(cudaFuncSetAttribute(tap_gemm_tc_kernel<BN, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
  p.n_tiles = l.layer->n_pad / BN;
  p.n = l.n;
  p.bias = l.layer->bias; p.bias_group_stride = l.bias_group_stride;
  p.C = l.C.p[0]; p.C_lo = l.C.p[1];
  p.c_group_stride = l.c_group_stride; p.c_pitch = l.c_pitch; p.c_fill = l.c_fill;
  p.relu = l.relu; p.out_mode = l.out_mode;
  const int m_tiles = ceil_div(p.M, kTcBM);
  p.num_tiles = m_tiles * p.groups * p.n_tiles;
  CDR_CHECK_ARG(l.layer->bn == BN && l.layer->n_pad % BN == 0, "tap_gemm_tc: layer packed for BN=%d, launched with %d",
                l.layer->bn, BN);
  CDR_CHECK_ARG((l.a_pitch * kElem) % 16 == 0 && ((uintptr_t)l.A.p[0] & 15) == 0, "tap_gemm_tc: A pitch/alignment");
  CDR_CHECK_ARG(l.out_mode == kOutPlanar || ((l.c_pitch * kElem) % 16 == 0 && (l.c_fill * kElem) % 16 == 0),
                "tap_gemm_tc: output pitch / fill must be 16-byte multiples");

  CUtensorMap tmap_a[2];
  for (int pl = 0; pl < KindTraits<KIND>::kPlanes; ++pl) {
    if (l.deconv) {
      CDR_CHECK_ARG(l.W <= kTcBM && kTcBM % l.W == 0, "tap_gemm_tc: W=%d must divide 128", l.W);
      int rows = kTcBM / l.W;
      if (rows > l.H) rows = l.H;
      const int imgs = kTcBM / (l.W * rows);
      CDR_CHECK_ARG(l.H % rows == 0 && l.cin % kBK == 0, "tap_gemm_tc: unsupported deconv geometry");
      p.box_rows = rows; p.box_imgs = imgs;
      const uint64_t dims[4] = {(uint64_t)l.cin, (uint64_t)l.W, (uint64_t)l.H, (uint64_t)l.n_img};
      const uint64_t strides[3] = {(uint64_t)l.a_pitch, (uint64_t)l.a_pitch * l.W, (uint64_t)l.a_pitch * HW};
      const uint32_t box[4] = {(uint32_t)kBK, (uint32_t)l.W, (uint32_t)rows, (uint32_t)imgs};
      if (int rc = make_tmap(&tmap_a[pl], l.A.p[pl], kElem, 4, dims, strides, box)) return rc;
    } else {
      const uint64_t dims[2] = {(uint64_t)l.cin, (uint64_t)l.a_rows_total};
      const uint64_t strides[1] = {(uint64_t)l.a_pitch};
      const uint32_t box[2] = {(uint32_t)kBK, (uint32_t)kTcBM};
      if (int rc = make_tmap(&tmap_a[pl], l.A.p[pl], kElem, 2, dims, strides, box)) return rc;
    }
  }
  if (KindTraits<KIND>::kPlanes == 1) tmap_a[1] = tmap_a[0];
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  tap_gemm_tc_kernel<BN, KIND><<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(
      tmap_a[0], tmap_a[1], l.layer->map[0], l.layer->map[KindTraits<KIND>::kPlanes - 1], p);
  CDR_LAUNCH_OK("tap_gemm_tc_kernel");
  return CDR_OK;
}

static int launch_tc(int kind, const TcLaunch& l, cudaStream_t st) {
  const int bn = l.layer->bn;
  if (kind == kKindBF16) {
    if (bn == 256) return launch_tc_t<256, kKindBF16>(l, st);
    if (bn == 128) return launch_tc_t<128, kKindBF16>(l, st);
    if (bn == 32) return launch_tc_t<32, kKindBF16>(l, st);
  } else {
    if (bn == 128) return launch_tc_t<128, kKindTF32X3>(l, st);
    if (bn == 32) return launch_tc_t<32, kKindTF32X3>(l, st);
  }
  set_error("tap_gemm_tc: no kernel for kind %d BN %d", kind, bn);
  return CDR_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------
// weight packing (K-major B operands) — BN folded in fp64 as in pack.cu
constexpr double kBnEpsTc = 1e-5;
__device__ __forceinline__ double tc_bn_scale(const CdrConvBn& s, int co) {
  return s.bn_weight ? (double)s.bn_weight[co] / sqrt((double)s.bn_var[co] + kBnEpsTc) : 1.0;
}
__device__ __forceinline__ float tc_folded_bias(const CdrConvBn& s, int co) {
  const double b = s.bias ? (double)s.bias[co] : 0.0;
  if (!s.bn_weight) return (float)b;
  return (float)((b - (double)s.bn_mean[co]) * tc_bn_scale(s, co) + (double)s.bn_bias[co]);
}
template <bool kSplit>
__device__ __forceinline__ void store_weight(void* w0, void* w1, long long idx, float v) {
  if constexpr (kSplit) {
    float hi, lo;
    split_tf32(v, hi, lo);
    reinterpret_cast<float*>(w0)[idx] = hi;
    reinterpret_cast<float*>(w1)[idx] = lo;
  } else {
    reinterpret_cast<__nv_bfloat16*>(w0)[idx] = __float2bfloat16_rn(v);
  }
}
// (Cout,Cin) -> [n_pad][k_pitch]
template <bool kSplit>
__global__ void pack_conv1x1_tc_kernel(CdrConvBn s, int cout, int cin, int k_pitch, int n_pad, void* w0,
                                       void* w1, float* __restrict__ bias_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_pad) bias_out[idx] = idx < cout ? tc_folded_bias(s, (int)idx) : 0.f;
  if (idx >= (long long)n_pad * k_pitch) return;
  const int n = (int)(idx / k_pitch), k = (int)(idx % k_pitch);
  float v = 0.f;
  if (n < cout && k < cin) v = (float)((double)s.weight[(size_t)n * cin + k] * tc_bn_scale(s, n));
  store_weight<kSplit>(w0, w1, idx, v);
}
// (Cin,Cout,4,4) -> [phase][n_pad][tap*Cin + ci]
template <bool kSplit>
__global__ void pack_deconv_tc_kernel(CdrConvBn s, int cin, int cout, int n_pad, void* w0, void* w1,
                                      float* __restrict__ bias_out) {
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
  store_weight<kSplit>(w0, w1, idx, v);
}

// plane(s) -> fp32 (parity taps)
__global__ void act_to_f32_kernel(const void* __restrict__ p0, const void* __restrict__ p1, int split,
                                  float* __restrict__ out, long long rows, int in_pitch, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = (int)(idx - r * cols);
  const long long i = r * in_pitch + c;
  out[idx] = split ? reinterpret_cast<const float*>(p0)[i] + reinterpret_cast<const float*>(p1)[i]
                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p0)[i]);
}

// ------------------------------------------------------------------------------------------
static const int kTcDcCin[3] = {kFeatC, kDecC, kDecC};

struct Bump1K {
  uint8_t* base;
  size_t off = 0;
  explicit Bump1K(void* b) : base((uint8_t*)b) {}
  void* take(size_t bytes) {
    void* r = base ? base + off : nullptr;
    off += round_up<size_t>(bytes, 1024);
    return r;
  }
};

static void plan_layer(TcLayer& L, Bump1K& b, int kind, int rows, int k, int k_pitch, int bn, int n_pad, int bias_n) {
  const int elem = kind == kKindBF16 ? 2 : 4;
  L.rows = rows; L.k = k; L.k_pitch = k_pitch; L.bn = bn; L.n_pad = n_pad;
  L.w[0] = b.take((size_t)rows * k_pitch * elem);
  L.w[1] = kind == kKindTF32X3 ? b.take((size_t)rows * k_pitch * elem) : nullptr;
  L.bias = (float*)b.take((size_t)bias_n * sizeof(float));
}

static size_t plan_tc_weights(TcPack& pk, const TcWeights& w, void* base) {
  Bump1K b(base);
  const int kind = pk.kind;
  const int bn_wide = kind == kKindBF16 ? 256 : 128;
  if (w.has_fusion) {
    plan_layer(pk.cf1, b, kind, 384, kFeatC, kFeatC, 128, 384, 384);
    plan_layer(pk.cf2a, b, kind, 512, 2 * kHid2, 2 * kHid2, 128, 512, 512);
    plan_layer(pk.cf2b, b, kind, 512, kHid2, kHid2, 128, 512, 512);
    plan_layer(pk.out, b, kind, 2 * kFeatC, kHid1, kHid1Pad, bn_wide, kFeatC, 2 * kFeatC);
  }
  for (int i = 0; i < 3; ++i)
    plan_layer(pk.dc[i], b, kind, 4 * kDecC, 4 * kTcDcCin[i], 4 * kTcDcCin[i], bn_wide, kDecC, kDecC);
  plan_layer(pk.fin, b, kind, w.fin_npad, kDecC, kDecC, 32, w.fin_npad, w.fin_npad);
  return b.off;
}

static int layer_maps(TcLayer& L, int kind) {
  const int elem = kind == kKindBF16 ? 2 : 4;
  const int bk = kind == kKindBF16 ? 64 : 32;
  for (int pl = 0; pl < (kind == kKindTF32X3 ? 2 : 1); ++pl) {
    const uint64_t dims[2] = {(uint64_t)L.k, (uint64_t)L.rows};
    const uint64_t strides[1] = {(uint64_t)L.k_pitch};
    const uint32_t box[2] = {(uint32_t)bk, (uint32_t)L.bn};
    if (int rc = make_tmap(&L.map[pl], L.w[pl], elem, 2, dims, strides, box)) return rc;
  }
  return CDR_OK;
}

int tc_weights_create(const CdrWeightPtrs& src, int kind, TcWeights& w, cudaStream_t st) {
  w.joints = src.num_joints;
  w.has_fusion = src.has_fusion;
  w.fin_npad = round_up(src.num_joints, 32);
  w.kind = kind;
  TcPack* pk = new TcPack();
  pk->kind = kind;
  w.impl = pk;
  const size_t bytes = plan_tc_weights(*pk, w, nullptr);
  CDR_CUDA(cudaMalloc(&w.pool, bytes));
  plan_tc_weights(*pk, w, w.pool);
  const bool split = kind == kKindTF32X3;
  auto pack1 = [&](const CdrConvBn& s, int cout, int cin, TcLayer& L, size_t row_off, size_t bias_off) -> int {
    const int rows = cout <= L.n_pad ? L.n_pad : cout;   // rows packed by this call
    const long long total = (long long)rows * L.k_pitch;
    const int elem = split ? 4 : 2;
    void* w0 = (uint8_t*)L.w[0] + row_off * L.k_pitch * elem;
    void* w1 = split ? (uint8_t*)L.w[1] + row_off * L.k_pitch * elem : nullptr;
    const unsigned grid = (unsigned)ceil_div<long long>(total, 256);
    if (split)
      pack_conv1x1_tc_kernel<true><<<grid, 256, 0, st>>>(s, cout, cin, L.k_pitch, rows, w0, w1, L.bias + bias_off);
    else
      pack_conv1x1_tc_kernel<false><<<grid, 256, 0, st>>>(s, cout, cin, L.k_pitch, rows, w0, w1, L.bias + bias_off);
    CDR_LAUNCH_OK("pack_conv1x1_tc_kernel");
    return CDR_OK;
  };
  int rc;
  if (w.has_fusion) {
    if ((rc = pack1(src.cf_conv1, kHid1, kFeatC, pk->cf1, 0, 0))) return rc;
    if ((rc = pack1(src.cf_conv2a, kHid2, 2 * kHid2, pk->cf2a, 0, 0))) return rc;
    if ((rc = pack1(src.cf_conv2b, kHid2, kHid2, pk->cf2b, 0, 0))) return rc;
    for (int v = 0; v < 2; ++v)
      if ((rc = pack1(src.cf_out[v], kFeatC, kHid1, pk->out, (size_t)v * kFeatC, (size_t)v * kFeatC))) return rc;
    if ((rc = layer_maps(pk->cf1, kind)) || (rc = layer_maps(pk->cf2a, kind)) || (rc = layer_maps(pk->cf2b, kind)) ||
        (rc = layer_maps(pk->out, kind)))
      return rc;
  }
  for (int i = 0; i < 3; ++i) {
    const long long total = 4LL * kDecC * 4 * kTcDcCin[i];
    const unsigned grid = (unsigned)ceil_div<long long>(total, 256);
    TcLayer& L = pk->dc[i];
    if (split)
      pack_deconv_tc_kernel<true><<<grid, 256, 0, st>>>(src.deconv[i], kTcDcCin[i], kDecC, kDecC, L.w[0], L.w[1], L.bias);
    else
      pack_deconv_tc_kernel<false><<<grid, 256, 0, st>>>(src.deconv[i], kTcDcCin[i], kDecC, kDecC, L.w[0], L.w[1], L.bias);
    CDR_LAUNCH_OK("pack_deconv_tc_kernel");
    if ((rc = layer_maps(L, kind))) return rc;
  }
  if ((rc = pack1(src.final_layer, w.joints, kDecC, pk->fin, 0, 0))) return rc;
  return layer_maps(pk->fin, kind);
}

void tc_weights_destroy(TcWeights& w) {
  if (w.pool) cudaFree(w.pool);
  w.pool = nullptr;
  delete (TcPack*)w.impl;
  w.impl = nullptr;
}

// ------------------------------------------------------------------------------------------
struct TcHeadWs {
  float* pinv;
  Act x0, y1, z, f1, f2, g, x1, d1, d2, d3;
  float* hm;
  size_t bytes;
};
static Act take_act(Bump1K& b, size_t elems, int kind) {
  Act a;
  const int elem = kind == kKindBF16 ? 2 : 4;
  a.p[0] = b.take(elems * elem);
  a.p[1] = kind == kKindTF32X3 ? b.take(elems * elem) : nullptr;
  return a;
}
static TcHeadWs plan_tc_head(void* base, int B, int J, int kind) {
  Bump1K b(base);
  const size_t N = 2 * (size_t)B;
  TcHeadWs w;
  w.pinv = (float*)b.take(N * 12 * sizeof(float));
  w.x0 = take_act(b, N * kFeatHW * kFeatC, kind);
  w.y1 = take_act(b, N * kFeatHW * kHid1Pad, kind);
  w.z = take_act(b, (size_t)B * kFeatHW * 2 * kHid2, kind);
  w.f1 = take_act(b, (size_t)B * kFeatHW * kHid2, kind);
  w.f2 = take_act(b, (size_t)B * kFeatHW * kHid2, kind);
  w.g = take_act(b, N * kFeatHW * kHid1Pad, kind);
  w.x1 = take_act(b, N * kFeatHW * kFeatC, kind);
  w.d1 = take_act(b, N * 256 * kDecC, kind);
  w.d2 = take_act(b, N * 1024 * kDecC, kind);
  w.d3 = take_act(b, N * 4096 * kDecC, kind);
  w.hm = (float*)b.take(N * J * 4096 * sizeof(float));
  w.bytes = b.off;
  return w;
}
struct TcDecWs {
  Act x1, d1, d2, d3;
  size_t bytes;
};
static TcDecWs plan_tc_dec(void* base, int N, int kind) {
  Bump1K b(base);
  TcDecWs w;
  w.x1 = take_act(b, (size_t)N * kFeatHW * kFeatC, kind);
  w.d1 = take_act(b, (size_t)N * 256 * kDecC, kind);
  w.d2 = take_act(b, (size_t)N * 1024 * kDecC, kind);
  w.d3 = take_act(b, (size_t)N * 4096 * kDecC, kind);
  w.bytes = b.off;
  return w;
}

int tc_head_workspace_bytes(const TcWeights& w, int batch, size_t* bytes) {
  *bytes = plan_tc_head(nullptr, batch, w.joints, w.kind).bytes;
  return CDR_OK;
}
int tc_decoder_workspace_bytes(const TcWeights& w, int n_images, size_t* bytes) {
  *bytes = plan_tc_dec(nullptr, n_images, w.kind).bytes;
  return CDR_OK;
}

static int to_rows(const float* feat, int n_img, const Act& out, int kind, cudaStream_t st) {
  if (kind == kKindBF16)
    return launch_nchw_to_rows_bf16(feat, n_img, kFeatC, kFeatHW, (__nv_bfloat16*)out.p[0], kFeatC, st);
  return launch_nchw_to_rows_split(feat, n_img, kFeatC, kFeatHW, (float*)out.p[0], (float*)out.p[1], kFeatC, st);
}
static int ftl_act(const Act& in, int in_pitch, const float* mats, int rows, int cols, int n, const Act& out,
                   int out_pitch, int out_fill, int kind, cudaStream_t st) {
  if (kind == kKindBF16)
    return launch_ftl<__nv_bfloat16>((const __nv_bfloat16*)in.p[0], in_pitch, mats, rows, cols, kFtlBlk, n, kFeatHW,
                                     (__nv_bfloat16*)out.p[0], out_pitch, out_fill, st);
  return launch_ftl_split((const float*)in.p[0], (const float*)in.p[1], in_pitch, mats, rows, cols, kFtlBlk, n,
                          kFeatHW, (float*)out.p[0], (float*)out.p[1], out_pitch, out_fill, st);
}

static int tc_decoder(const TcWeights& w, const Act& x1, int N, const Act& d1, const Act& d2, const Act& d3,
                      float* heat, cudaStream_t st) {
  const TcPack* pk = (const TcPack*)w.impl;
  static const char* const kDcName[3] = {"deconv1", "deconv2", "deconv3"};
  Act in = x1;
  const Act outs[3] = {d1, d2, d3};
  int side = 8;
  for (int i = 0; i < 3; ++i) {
    set_stage(kDcName[i]);
    TcLaunch l{};
    l.A = in; l.a_pitch = kTcDcCin[i]; l.n_img = N; l.H = l.W = side; l.cin = kTcDcCin[i];
    l.deconv = 1; l.groups = 4; l.layer = &pk->dc[i]; l.b_group_rows = kDecC; l.n = kDecC;
    l.bias_group_stride = 0;
    l.C = outs[i]; l.c_pitch = kDecC; l.c_fill = kDecC; l.relu = 1; l.out_mode = kOutDeconv;
    if (int rc = launch_tc(w.kind, l, st)) return rc;
    in = outs[i];
    side *= 2;
  }
  set_stage("final_1x1");
  TcLaunch l{};
  l.A = d3; l.a_pitch = kDecC; l.n_img = N; l.H = l.W = kHeat; l.cin = kDecC; l.groups = 1;
  l.a_rows_total = (long long)N * 4096;
  l.layer = &pk->fin; l.n = w.joints;
  l.C.p[0] = heat; l.relu = 0; l.out_mode = kOutPlanar;
  const int rc = launch_tc(w.kind, l, st);
  set_stage(nullptr);
  return rc;
}

static int tap_to_f32(float* dst, const Act& src, int kind, long long rows, int pitch, int cols, cudaStream_t st) {
  if (!dst) return CDR_OK;
  act_to_f32_kernel<<<(unsigned)ceil_div<long long>(rows * cols, 256), 256, 0, st>>>(
      src.p[0], src.p[1], kind == kKindTF32X3, dst, rows, pitch, cols);
  CDR_LAUNCH_OK("act_to_f32_kernel");
  return CDR_OK;
}

int tc_head_forward(const TcWeights& w, const float* feat_l, const float* feat_r, const float* P_l,
                    const float* P_r, const float* pinv_l, const float* pinv_r, double pinv_rtol,
                    int batch, float scale, float* kp2d_l, float* kp2d_r, float* xyz,
                    const CdrHeadTaps* taps, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const TcPack* pk = (const TcPack*)w.impl;
  const int kind = w.kind;
  const int elem = kind == kKindBF16 ? 2 : 4;
  const int B = batch, N = 2 * batch, J = w.joints;
  TcHeadWs ws = plan_tc_head(workspace, B, J, kind);
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
  if ((rc = to_rows(feat_l, B, ws.x0, kind, st))) return rc;
  if ((rc = to_rows(feat_r, B, act_offset(ws.x0, (size_t)B * kFeatHW * kFeatC, elem), kind, st))) return rc;
  set_stage("cf_conv1");
  {
    TcLaunch l{};
    l.A = ws.x0; l.a_pitch = kFeatC; l.n_img = N; l.H = l.W = 8; l.cin = kFeatC; l.groups = 1;
    l.a_rows_total = (long long)N * kFeatHW;
    l.layer = &pk->cf1; l.n = kHid1;
    l.C = ws.y1; l.c_pitch = kHid1Pad; l.c_fill = kHid1Pad; l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc(kind, l, st))) return rc;
  }
  set_stage("ftl_inv");
  for (int v = 0; v < 2; ++v)
    if ((rc = ftl_act(act_offset(ws.y1, (size_t)v * B * kFeatHW * kHid1Pad, elem), kHid1Pad, pinv[v], 4, 3, B,
                      act_offset(ws.z, (size_t)v * kHid2, elem), 2 * kHid2, kHid2, kind, st)))
      return rc;
  set_stage("cf_conv2");
  {
    TcLaunch l{};
    l.A = ws.z; l.a_pitch = 2 * kHid2; l.n_img = B; l.H = l.W = 8; l.cin = 2 * kHid2; l.groups = 1;
    l.a_rows_total = (long long)B * kFeatHW;
    l.layer = &pk->cf2a; l.n = kHid2;
    l.C = ws.f1; l.c_pitch = kHid2; l.c_fill = kHid2; l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc(kind, l, st))) return rc;
    l.A = ws.f1; l.a_pitch = kHid2; l.cin = kHid2; l.layer = &pk->cf2b; l.C = ws.f2;
    if ((rc = launch_tc(kind, l, st))) return rc;
  }
  set_stage("ftl_fwd");
  const float* Pv[2] = {P_l, P_r};
  for (int v = 0; v < 2; ++v)
    if ((rc = ftl_act(ws.f2, kHid2, Pv[v], 3, 4, B, act_offset(ws.g, (size_t)v * B * kFeatHW * kHid1Pad, elem),
                      kHid1Pad, kHid1Pad, kind, st)))
      return rc;
  set_stage("cf_out");
  {
    TcLaunch l{};
    l.A = ws.g; l.a_pitch = kHid1Pad; l.n_img = B; l.H = l.W = 8; l.cin = kHid1; l.groups = 2;
    l.a_rows_total = (long long)N * kFeatHW;
    l.layer = &pk->out; l.b_group_rows = kFeatC; l.n = kFeatC; l.bias_group_stride = kFeatC;
    l.C = ws.x1; l.c_group_stride = (long long)B * kFeatHW * kFeatC; l.c_pitch = kFeatC; l.c_fill = kFeatC;
    l.relu = 1; l.out_mode = kOutRows;
    if ((rc = launch_tc(kind, l, st))) return rc;
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
    if ((rc = tap_to_f32(taps->cf_cat, ws.z, kind, (long long)B * kFeatHW, 2 * kHid2, 2 * kHid2, st))) return rc;
    if ((rc = tap_to_f32(taps->cf_f, ws.f2, kind, (long long)B * kFeatHW, kHid2, kHid2, st))) return rc;
    if ((rc = tap_to_f32(taps->f_out, ws.x1, kind, (long long)N * kFeatHW, kFeatC, kFeatC, st))) return rc;
    if (taps->heatmaps)
      CDR_CUDA(cudaMemcpyAsync(taps->heatmaps, ws.hm, (size_t)N * J * 4096 * 4, cudaMemcpyDeviceToDevice, st));
  }
  return CDR_OK;
}

int tc_decoder_forward(const TcWeights& w, const float* feat, int n_images, float* heatmaps,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
  TcDecWs ws = plan_tc_dec(workspace, n_images, w.kind);
  if (ws.bytes > workspace_bytes) {
    set_error("cdr_decoder_forward: workspace %zu < required %zu bytes", workspace_bytes, ws.bytes);
    return CDR_ERR_WORKSPACE;
  }
  set_stage("nchw_to_rows");
  if (int rc = to_rows(feat, n_images, ws.x1, w.kind, st)) return rc;
  return tc_decoder(w, ws.x1, n_images, ws.d1, ws.d2, ws.d3, heatmaps, st);
}

}  // namespace cdr
