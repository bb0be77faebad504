// tcgen05 / TMEM implementation of the TransformerBlock passes (C = 136).
// Placeholder until the tensor-core kernels land: reports "not built" so the
// C ABI routes OTP_PREC_BF16 requests to the fp32 CUDA-core kernels.
#include "block_common.cuh"

namespace otp {
bool block_tc_built() { return false; }
size_t block_tc_packed_bytes(int) { return 0; }
int block_tc_pack(const otp_block_params *, int, void *, cudaStream_t) { return OTP_OK; }
size_t block_tc_workspace_bytes(int, int, int, int) { return 0; }
int block_forward_tc(const void *, const void *, const float *, float *, int, int, int, int, void *, void *,
                     cudaStream_t) {
  set_error("tcgen05 block path not built");
  return OTP_ERR_UNSUPPORTED;
}
}  // namespace otp
