// pybind11 module `deform_conv_cuda` with the entry points the reference's OWN
// thirdparty/deform_conv/functions/deform_conv.py binds (reference
// thirdparty/deform_conv/src/deform_conv_cuda.cpp:474-480, 551-558, 666-680), implemented on the
// C ABI of libotpose_b200.so.  With this module on the import path the reference's
// ModulatedDeformConvFunction / ModulatedDeformConv run UNCHANGED on the B200 kernels
// (INTEGRATION.md section 3); tests/test_gpu_shim.py drives it with the reference's call sequence.
//
// Contracts kept: caller allocates `output` and the (zeroed) gradient tensors; `ones` / `columns` are
// caller-passed scratch tensors (unused here: the kernels keep no im2col buffer); contiguity and device
// are checked with TORCH_CHECK -> Python RuntimeError.  Differences, on purpose: kernels run on
// PyTorch's CURRENT stream (the reference launches on the legacy default stream,
// deform_conv_cuda_kernel.cu:726), launch errors raise instead of being printf-ed (.cu:732-736), and
// the gradients are written, not accumulated -- identical for the reference's caller, which passes
// zeros_like tensors (functions/deform_conv.py:152-156).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "otpose_b200.h"

namespace {
void check_f32_cuda(const at::Tensor &t, const char *name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.is_contiguous(), name, " tensor has to be contiguous");
  TORCH_CHECK(t.scalar_type() == at::kFloat, name, ": the B200 DCN kernels are built for float32 (got ",
              t.scalar_type(), ")");
}
void check_square(int a, int b, const char *what) {
  TORCH_CHECK(a == b, "modulated_deform_conv: ", what, " must be equal in h and w (got ", a, ", ", b, ")");
}
void raise_if(int status, const char *what) { TORCH_CHECK(status == 0, what, ": ", otp_last_error()); }
}  // namespace

void modulated_deform_conv_cuda_forward(at::Tensor input, at::Tensor weight, at::Tensor bias, at::Tensor ones,
                                        at::Tensor offset, at::Tensor mask, at::Tensor output, at::Tensor columns,
                                        int kernel_h, int kernel_w, const int stride_h, const int stride_w,
                                        const int pad_h, const int pad_w, const int dilation_h,
                                        const int dilation_w, const int group, const int deformable_group,
                                        const bool with_bias) {
  (void)ones;
  (void)columns;
  check_f32_cuda(input, "input");
  check_f32_cuda(weight, "weight");
  check_f32_cuda(offset, "offset");
  check_f32_cuda(mask, "mask");
  check_f32_cuda(output, "output");
  if (with_bias) check_f32_cuda(bias, "bias");
  check_square(stride_h, stride_w, "stride");
  check_square(pad_h, pad_w, "padding");
  check_square(dilation_h, dilation_w, "dilation");
  TORCH_CHECK(weight.size(2) == kernel_h && weight.size(3) == kernel_w, "kernel size mismatch");
  TORCH_CHECK(weight.size(1) * group == input.size(1), "input channels / weight / group mismatch");
  c10::cuda::CUDAGuard guard(input.device());
  raise_if(otp_mdcn_forward(input.data_ptr<float>(), offset.data_ptr<float>(), mask.data_ptr<float>(),
                            weight.data_ptr<float>(), with_bias ? bias.data_ptr<float>() : nullptr,
                            output.data_ptr<float>(), (int)input.size(0), (int)input.size(1), (int)input.size(2),
                            (int)input.size(3), (int)weight.size(0), kernel_h, kernel_w, stride_h, pad_h, dilation_h,
                            group, deformable_group, 1.0f, 0, at::cuda::getCurrentCUDAStream().stream()),
           "modulated_deform_conv_cuda_forward");
}

void modulated_deform_conv_cuda_backward(at::Tensor input, at::Tensor weight, at::Tensor bias, at::Tensor ones,
                                         at::Tensor offset, at::Tensor mask, at::Tensor columns,
                                         at::Tensor grad_input, at::Tensor grad_weight, at::Tensor grad_bias,
                                         at::Tensor grad_offset, at::Tensor grad_mask, at::Tensor grad_output,
                                         int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                                         int dilation_h, int dilation_w, int group, int deformable_group,
                                         const bool with_bias) {
  (void)ones;
  (void)columns;
  (void)bias;
  check_f32_cuda(input, "input");
  check_f32_cuda(weight, "weight");
  check_f32_cuda(offset, "offset");
  check_f32_cuda(mask, "mask");
  check_f32_cuda(grad_input, "grad_input");
  check_f32_cuda(grad_weight, "grad_weight");
  check_f32_cuda(grad_offset, "grad_offset");
  check_f32_cuda(grad_mask, "grad_mask");
  if (with_bias) check_f32_cuda(grad_bias, "grad_bias");
  check_square(stride_h, stride_w, "stride");
  check_square(pad_h, pad_w, "padding");
  check_square(dilation_h, dilation_w, "dilation");
  at::Tensor go = grad_output.contiguous();   // the reference does the same (deform_conv_cuda.cpp:560)
  check_f32_cuda(go, "grad_output");
  c10::cuda::CUDAGuard guard(input.device());
  const int b = (int)input.size(0), c = (int)input.size(1), h = (int)input.size(2), w = (int)input.size(3);
  const int cout = (int)weight.size(0);
  const size_t nws = otp_mdcn_backward_workspace_bytes(b, c, h, w, cout, kernel_h, kernel_w, stride_h, pad_h, dilation_h);
  TORCH_CHECK(nws > 0, "modulated_deform_conv_cuda_backward: ", otp_last_error());
  at::Tensor ws = at::empty({(int64_t)nws}, input.options().dtype(at::kByte));
  raise_if(otp_mdcn_backward(input.data_ptr<float>(), offset.data_ptr<float>(), mask.data_ptr<float>(),
                             weight.data_ptr<float>(), go.data_ptr<float>(), grad_input.data_ptr<float>(),
                             grad_offset.data_ptr<float>(), grad_mask.data_ptr<float>(), grad_weight.data_ptr<float>(),
                             with_bias ? grad_bias.data_ptr<float>() : nullptr, b, c, h, w, cout, kernel_h, kernel_w,
                             stride_h, pad_h, dilation_h, group, deformable_group, ws.data_ptr(), nws,
                             at::cuda::getCurrentCUDAStream().stream()),
           "modulated_deform_conv_cuda_backward");
}

// DCN v1 (deform_conv_cuda.cpp:151-472): exported by the reference module, never reached by OTPose
// (model/OTPose.py builds ModulatedDeformConv only; SURVEY.md section 2 marks v1 out of scope).
static int dcn_v1_not_built() {
  TORCH_CHECK(false, "deform_conv v1 is not part of the OTPose hot path and is not built in otpose_b200");
  return 0;
}
int deform_conv_forward_cuda(at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, int, int, int, int,
                             int, int, int, int, int, int, int) {
  return dcn_v1_not_built();
}
int deform_conv_backward_input_cuda(at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor,
                                    at::Tensor, int, int, int, int, int, int, int, int, int, int, int) {
  return dcn_v1_not_built();
}
int deform_conv_backward_parameters_cuda(at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor, int,
                                         int, int, int, int, int, int, int, int, int, float, int) {
  return dcn_v1_not_built();
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("deform_conv_forward_cuda", &deform_conv_forward_cuda, "deform forward (CUDA) -- not built");
  m.def("deform_conv_backward_input_cuda", &deform_conv_backward_input_cuda, "deform_conv_backward_input (CUDA) -- not built");
  m.def("deform_conv_backward_parameters_cuda", &deform_conv_backward_parameters_cuda,
        "deform_conv_backward_parameters (CUDA) -- not built");
  m.def("modulated_deform_conv_cuda_forward", &modulated_deform_conv_cuda_forward,
        "modulated deform conv forward (CUDA, otpose_b200)");
  m.def("modulated_deform_conv_cuda_backward", &modulated_deform_conv_cuda_backward,
        "modulated deform conv backward (CUDA, otpose_b200)");
}
