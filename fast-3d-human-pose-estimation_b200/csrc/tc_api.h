// Interface between api.cu and the bf16 tensor-core path (gemm_tc.cu).
#pragma once
#include "kernels.h"

namespace cdr {

struct TcWeights {
  void* pool = nullptr;
  void* maps = nullptr;   // TcMaps: tensor maps of the packed weights (gemm_tc.cu)
  int joints = 0, has_fusion = 0, fin_npad = 0;
  // bf16 B operands, K-major: [n_pad][k_pad] (1x1) or [phase][n][tap*cin] (deconv)
  __nv_bfloat16 *w_cf1 = nullptr, *w_cf2a = nullptr, *w_cf2b = nullptr, *w_out = nullptr;
  __nv_bfloat16 *w_dc[3] = {nullptr, nullptr, nullptr}, *w_fin = nullptr;
  float *b_cf1 = nullptr, *b_cf2a = nullptr, *b_cf2b = nullptr, *b_out = nullptr;
  float *b_dc[3] = {nullptr, nullptr, nullptr}, *b_fin = nullptr;
};

int tc_weights_create(const CdrWeightPtrs& src, TcWeights& w, cudaStream_t st);
void tc_weights_destroy(TcWeights& w);
int tc_head_workspace_bytes(const TcWeights& w, int batch, size_t* bytes);
int tc_decoder_workspace_bytes(const TcWeights& w, int n_images, size_t* bytes);
int tc_head_forward(const TcWeights& w, const float* feat_l, const float* feat_r, const float* P_l,
                    const float* P_r, const float* pinv_l, const float* pinv_r, double pinv_rtol,
                    int batch, float scale, float* kp2d_l, float* kp2d_r, float* xyz,
                    const CdrHeadTaps* taps, void* workspace, size_t workspace_bytes,
                    cudaStream_t st);
int tc_decoder_forward(const TcWeights& w, const float* feat, int n_images, float* heatmaps,
                       void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace cdr
