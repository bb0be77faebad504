// a9: modulated deformable convolution, forward.
// Replaces thirdparty/deform_conv/src/deform_conv_cuda.cpp:474-549 (per-sample host
// loop: im2col kernel -> cuBLAS addmm -> bias) and deform_conv_cuda_kernel.cu:402-432,
// 505-571 with ONE launch over the whole batch: each thread owns one output pixel,
// streams its offset/mask values with coalesced loads (consecutive threads =
// consecutive pixels of the same offset channel), samples the (L1/L2 resident)
// input plane bilinearly and contracts with the weight tile held in shared memory,
// accumulating all output channels in registers -- the (C*kh*kw, H*W) `columns`
// matrix never exists.
//
// HBM-bound: algorithmic bytes per launch =
//   B * Ho*Wo * 4 * (3*dg*kh*kw [offset+mask] + Cout [out]) + B*C*H*W*4 [x].
#include "common.cuh"

namespace otp {

constexpr int kDcnThreads = 128;

template <int CO>
__global__ void __launch_bounds__(kDcnThreads)
mdcn_fwd_kernel(const float *__restrict__ x, const float *__restrict__ offset,
                const float *__restrict__ mask, const float *__restrict__ weight,
                const float *__restrict__ bias, float *__restrict__ out, int C, int H, int W,
                int Cout, int kh, int kw, int stride, int pad, int dil, int dg, int Ho, int Wo,
                float alpha, int accumulate) {
  extern __shared__ float ws[];  // [C*kh*kw][CO] weight tile, output channel fastest
  const int K2 = kh * kw;
  const int CK = C * K2;
  const int co0 = blockIdx.z * CO;
  for (int e = threadIdx.x; e < CK * CO; e += kDcnThreads) {
    int o = e % CO, ck = e / CO;
    ws[e] = (co0 + o < Cout) ? __ldg(weight + (size_t)(co0 + o) * CK + ck) : 0.f;
  }
  __syncthreads();

  const int P = Ho * Wo;
  const int p = blockIdx.x * kDcnThreads + threadIdx.x;
  if (p >= P) return;
  const int b = blockIdx.y;
  const int h_col = p / Wo, w_col = p % Wo;
  const int h_in = h_col * stride - pad, w_in = w_col * stride - pad;
  const int cpg = C / dg;
  const float *__restrict__ offb = offset + (size_t)b * dg * 2 * K2 * P + p;
  const float *__restrict__ mskb = mask + (size_t)b * dg * K2 * P + p;
  const float *__restrict__ xb = x + (size_t)b * C * H * W;

  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = 0.f;

  for (int c = 0; c < C; ++c) {
    const int g = c / cpg;
    const float *__restrict__ img = xb + (size_t)c * H * W;
    const float *__restrict__ og = offb + (size_t)g * 2 * K2 * P;
    const float *__restrict__ mg = mskb + (size_t)g * K2 * P;
    for (int i = 0; i < kh; ++i) {
      for (int j = 0; j < kw; ++j) {
        const int tap = i * kw + j;
        const float off_h = __ldg(og + (size_t)(2 * tap) * P);
        const float off_w = __ldg(og + (size_t)(2 * tap + 1) * P);
        const float m = __ldg(mg + (size_t)tap * P);
        const float h_im = (float)(h_in + i * dil) + off_h;
        const float w_im = (float)(w_in + j * dil) + off_w;
        float val = 0.f;
        if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
          // dmcn_im2col_bilinear (deform_conv_cuda_kernel.cu:402-432)
          const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
          const int h_high = h_low + 1, w_high = w_low + 1;
          const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
          const float hh = 1.f - lh, hw = 1.f - lw;
          float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
          if (h_low >= 0 && w_low >= 0) v1 = __ldg(img + h_low * W + w_low);
          if (h_low >= 0 && w_high <= W - 1) v2 = __ldg(img + h_low * W + w_high);
          if (h_high <= H - 1 && w_low >= 0) v3 = __ldg(img + h_high * W + w_low);
          if (h_high <= H - 1 && w_high <= W - 1) v4 = __ldg(img + h_high * W + w_high);
          val = (hh * hw) * v1 + (hh * lw) * v2 + (lh * hw) * v3 + (lh * lw) * v4;
        }
        const float col = val * m;
        const float *__restrict__ wr = ws + (c * K2 + tap) * CO;
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = fmaf(wr[o], col, acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    if (co0 + o < Cout) {
      float v = acc[o] + (bias ? __ldg(bias + co0 + o) : 0.f);
      float *dst = out + ((size_t)b * Cout + co0 + o) * P + p;
      *dst = accumulate ? fmaf(alpha, v, *dst) : alpha * v;
    }
  }
}

template <int CO>
static int launch_mdcn(const float *x, const float *offset, const float *mask, const float *weight,
                       const float *bias, float *out, int b, int c, int h, int w, int cout, int kh,
                       int kw, int stride, int pad, int dil, int dg, int ho, int wo, float alpha,
                       int accumulate, cudaStream_t st) {
  size_t smem = (size_t)c * kh * kw * CO * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("otp_mdcn_forward: weight tile of %zu B does not fit shared memory", smem);
    return OTP_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024 && !set_max_smem(mdcn_fwd_kernel<CO>, smem, "mdcn_fwd_kernel")) return OTP_ERR_CUDA;
  dim3 grid(ceil_div(ho * wo, kDcnThreads), b, ceil_div(cout, CO));
  LaunchScope ls(K_MDCN, st);
  mdcn_fwd_kernel<CO><<<grid, kDcnThreads, smem, st>>>(x, offset, mask, weight, bias, out, c, h, w,
                                                        cout, kh, kw, stride, pad, dil, dg, ho, wo,
                                                        alpha, accumulate);
  return check_launch("mdcn_fwd_kernel");
}

}  // namespace otp

extern "C" int otp_mdcn_forward(const float *x, const float *offset, const float *mask,
                                const float *weight, const float *bias, float *out, int b, int c,
                                int h, int w, int cout, int kh, int kw, int stride, int pad,
                                int dilation, int groups, int deformable_groups, float alpha,
                                int accumulate, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && h > 0 && w > 0 && cout > 0 && kh > 0 && kw > 0);
  OTP_REQUIRE(stride > 0 && pad >= 0 && dilation > 0 && deformable_groups > 0);
  OTP_REQUIRE(c % deformable_groups == 0);
  if (groups != 1) {
    otp::set_error("otp_mdcn_forward: groups=%d unsupported (OTPose uses groups=1)", groups);
    return OTP_ERR_UNSUPPORTED;
  }
  const int ho = (h + 2 * pad - (dilation * (kh - 1) + 1)) / stride + 1;
  const int wo = (w + 2 * pad - (dilation * (kw - 1) + 1)) / stride + 1;
  OTP_REQUIRE(ho > 0 && wo > 0);
  OTP_REQUIRE(b <= 65535);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && offset && mask && weight && out);
  cudaStream_t st = (cudaStream_t)stream;
  if (cout == 17)
    return otp::launch_mdcn<17>(x, offset, mask, weight, bias, out, b, c, h, w, cout, kh, kw, stride,
                                pad, dilation, deformable_groups, ho, wo, alpha, accumulate, st);
  return otp::launch_mdcn<16>(x, offset, mask, weight, bias, out, b, c, h, w, cout, kh, kw, stride,
                              pad, dilation, deformable_groups, ho, wo, alpha, accumulate, st);
}
