// ktb_mlp.cu — bf16 MLP policy callable (BASELINE config C4). Placeholder until the tcgen05
// kernel lands: the entry points exist so the ABI is complete, and fail loudly.
#include "ktb_common.cuh"

using namespace ktb;

extern "C" {

size_t ktb_mlp_scratch_bytes(size_t M, int d_hidden) {
  (void)M;
  (void)d_hidden;
  return 0;
}

int ktb_mlp_bf16(int dev, const void* obs, size_t M, int d_in, int d_hidden, int d_out, const void* W1,
                 const void* W2, const void* W3, void* logits, void* scratch, uintptr_t stream) {
  (void)dev; (void)obs; (void)M; (void)d_in; (void)d_hidden; (void)d_out; (void)W1; (void)W2; (void)W3;
  (void)logits; (void)scratch; (void)stream;
  set_error("ktb_mlp_bf16: not implemented in this build");
  return KTB_ERR_UNSUPPORTED;
}

}  // extern "C"
