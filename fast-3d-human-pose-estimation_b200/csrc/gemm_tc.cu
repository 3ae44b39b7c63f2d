// bf16 tcgen05 path — placeholder until the tensor-core kernels land (next milestone).
#include "tc_api.h"

namespace cdr {

static int unsupported(const char* who) {
  set_error("%s: CDR_PREC_BF16 (tcgen05 path) is not built yet", who);
  return CDR_ERR_UNSUPPORTED;
}
int tc_weights_create(const CdrWeightPtrs&, TcWeights&, cudaStream_t) { return unsupported("cdr_weights_create"); }
void tc_weights_destroy(TcWeights& w) { if (w.pool) cudaFree(w.pool); w.pool = nullptr; }
int tc_head_workspace_bytes(const TcWeights&, int, size_t*) { return unsupported("cdr_head_workspace_bytes"); }
int tc_decoder_workspace_bytes(const TcWeights&, int, size_t*) { return unsupported("cdr_decoder_workspace_bytes"); }
int tc_head_forward(const TcWeights&, const float*, const float*, const float*, const float*,
                    const float*, const float*, double, int, float, float*, float*, float*,
                    const CdrHeadTaps*, void*, size_t, cudaStream_t) { return unsupported("cdr_head_forward"); }
int tc_decoder_forward(const TcWeights&, const float*, int, float*, void*, size_t, cudaStream_t) { return unsupported("cdr_decoder_forward"); }

}  // namespace cdr
