// C ABI of the TransformerBlock passes (pack / workspace query / forward).
#include "block_common.cuh"

namespace otp {
int block_pack_fp32(const otp_block_params *p, int c, float *f, cudaStream_t st);
int block_forward_simt(const void *packed, const float *x, float *y, int b, int c, int t, int stride,
                       void *ws, cudaStream_t st);
// tcgen05 path (block_tc.cu)
bool block_tc_built();
void block_tc_trace(int on);
int block_tc_trace_read(unsigned long long *out, int n);
size_t block_tc_packed_bytes(int c);
int block_tc_pack(const otp_block_params *p, int c, void *packed_tc, cudaStream_t st);
size_t block_tc_workspace_bytes(int b, int c, int t, int stride);
int block_forward_tc(const void *packed_fp32, const void *packed_tc, const float *x, float *y, int b,
                     int c, int t, int stride, int f16, void *ws_tc, cudaStream_t st);

static int check_shape(int c, int n_head) {
  if (!((c == 136 && n_head == 2) || (c == 17 && n_head == 1))) {
    set_error("TransformerBlock width C=%d / n_head=%d not built (OTPose uses C=136,nh=2 and C=17,nh=1)",
              c, n_head);
    return OTP_ERR_UNSUPPORTED;
  }
  return OTP_OK;
}
static size_t fp32_pack_bytes(int c) { return align_up(block_pack_layout(c).total * 4, 1024); }
}  // namespace otp

using namespace otp;

extern "C" int otp_has_tensor_core_path(void) { return block_tc_built() ? 1 : 0; }

extern "C" int otp_debug_trace(int on) {
  block_tc_trace(on);
  return OTP_OK;
}
extern "C" int otp_debug_trace_read(unsigned long long *out, int n) {
  OTP_REQUIRE(out != nullptr);
  return block_tc_trace_read(out, n);
}

extern "C" size_t otp_block_packed_bytes(int c, int n_head) {
  if (check_shape(c, n_head) != OTP_OK) return 0;
  return fp32_pack_bytes(c) + block_tc_packed_bytes(c);
}

extern "C" int otp_block_pack(const otp_block_params *p, int c, int n_head, void *packed,
                              size_t packed_bytes, otp_stream_t stream) {
  if (int e = check_shape(c, n_head)) return e;
  OTP_REQUIRE(p != nullptr && packed != nullptr);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0);
  if (packed_bytes < otp_block_packed_bytes(c, n_head)) {
    set_error("otp_block_pack: buffer of %zu B, need %zu B", packed_bytes, otp_block_packed_bytes(c, n_head));
    return OTP_ERR_WORKSPACE;
  }
  OTP_REQUIRE(p->ln1_w && p->ln1_b && p->ln2_w && p->ln2_b && p->q_conv_w && p->k_conv_w && p->v_conv_w);
  OTP_REQUIRE(p->q_norm_w && p->q_norm_b && p->k_norm_w && p->k_norm_b && p->v_norm_w && p->v_norm_b);
  OTP_REQUIRE(p->q_w && p->q_b && p->k_w && p->k_b && p->v_w && p->v_b && p->proj_w && p->proj_b);
  OTP_REQUIRE(p->mlp0_w && p->mlp0_b && p->mlp3_w && p->mlp3_b);
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = block_pack_fp32(p, c, static_cast<float *>(packed), st)) return e;
  if (block_tc_packed_bytes(c) > 0)
    return block_tc_pack(p, c, static_cast<char *>(packed) + fp32_pack_bytes(c), st);
  return OTP_OK;
}

static bool use_tc(int c, int precision) {
  return (precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16) && block_tc_packed_bytes(c) > 0;
}

extern "C" size_t otp_block_workspace_bytes(int b, int c, int t, int n_head, int stride, int precision) {
  if (check_shape(c, n_head) != OTP_OK || b <= 0 || t <= 0 || (stride != 1 && stride != 2)) return 0;
  if (use_tc(c, precision)) return block_tc_workspace_bytes(b, c, t, stride);
  return block_workspace(b, c, t, n_head, stride).total;
}

extern "C" int otp_block_forward(const void *packed, const float *x, float *y, int b, int c, int t,
                                 int n_head, int stride, int precision, void *workspace,
                                 size_t workspace_bytes, otp_stream_t stream) {
  if (int e = check_shape(c, n_head)) return e;
  OTP_REQUIRE(b >= 0 && t > 0 && b <= 65535);
  OTP_REQUIRE(stride == 1 || stride == 2);
  OTP_REQUIRE(precision == OTP_PREC_FP32 || precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(packed && x && y && workspace && x != y);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  const size_t need = otp_block_workspace_bytes(b, c, t, n_head, stride, precision);
  if (workspace_bytes < need) {
    set_error("otp_block_forward: workspace of %zu B, need %zu B", workspace_bytes, need);
    return OTP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(c, precision))
    return block_forward_tc(packed, static_cast<const char *>(packed) + fp32_pack_bytes(c), x, y, b, c, t, stride,
                            precision == OTP_PREC_FP16, workspace, st);
  // C = 17 (flow encoder, 0.6 % of the head's FLOPs) has no tensor-core shape:
  // it runs the fp32 CUDA-core kernels in every precision mode.
  return block_forward_simt(packed, x, y, b, c, t, stride, workspace, st);
}
