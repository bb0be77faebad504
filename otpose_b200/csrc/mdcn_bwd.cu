// a12 (native part): modulated deformable convolution, backward.
// Replaces modulated_deform_conv_cuda_backward (thirdparty/deform_conv/src/
// deform_conv_cuda.cpp:551-664: per-sample host loop of GEMM + col2im + col2im_coord
// kernels + GEMM) and deform_conv_cuda_kernel.cu:434-503, 573-705 with two launches
// over the whole batch:
//
//   mdcn_bwd_kernel     one thread per output pixel: d_col = W^T d_out in registers,
//                       grad_offset / grad_mask written directly (the thread owns them),
//                       grad_input scattered with 64-bit FIXED-POINT atomics -- integer
//                       addition is associative, so unlike the reference's float atomicAdd
//                       (deform_conv_cuda_kernel.cu:626) the result is bit-reproducible;
//                       the CTA then contracts its im2col tile (kept in shared memory) with
//                       d_out into a per-CTA partial of grad_weight / grad_bias.
//   mdcn_bwd_finish     fixed-order sum of the per-CTA partials; fixed-point -> fp32 grad_input.
#include "common.cuh"

namespace otp {

constexpr int kBwThreads = 128;
constexpr double kFixScale = 1073741824.0;   // 2^30: |grad| < 8.6e9, resolution 9.3e-10

template <int CO>
__global__ void __launch_bounds__(kBwThreads)
mdcn_bwd_kernel(const float *__restrict__ x, const float *__restrict__ offset, const float *__restrict__ mask,
                const float *__restrict__ weight, const float *__restrict__ gout, long long *__restrict__ gx_fix,
                float *__restrict__ goff, float *__restrict__ gmask, float *__restrict__ gw_part,
                float *__restrict__ gb_part, int C, int H, int W, int kh, int kw, int stride, int pad, int dil,
                int dg, int Ho, int Wo) {
  extern __shared__ float sm[];
  const int K2 = kh * kw, CK = C * K2, LD = kBwThreads + 1;
  float *ws = sm;                    // [CK][CO]   weight, output channel fastest
  float *cols = ws + CK * CO;        // [CK][LD]   im2col tile of this CTA's pixels
  float *douts = cols + CK * LD;     // [CO][LD]
  for (int e = threadIdx.x; e < CK * CO; e += kBwThreads) {
    const int o = e % CO, ck = e / CO;
    ws[e] = __ldg(weight + (size_t)o * CK + ck);
  }
  const int P = Ho * Wo, b = blockIdx.y;
  const int p = blockIdx.x * kBwThreads + threadIdx.x;
  const bool live = p < P;
  float dout[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    dout[o] = live ? __ldg(gout + ((size_t)b * CO + o) * P + p) : 0.f;
    douts[o * LD + threadIdx.x] = dout[o];
  }
  __syncthreads();
  const int h_col = live ? p / Wo : 0, w_col = live ? p % Wo : 0;
  const int h_in = h_col * stride - pad, w_in = w_col * stride - pad;
  const int cpg = C / dg;
  const float *xb = x + (size_t)b * C * H * W;
  for (int g = 0; g < dg; ++g) {
    const float *og = offset + ((size_t)b * dg + g) * 2 * K2 * P + p;
    const float *mg = mask + ((size_t)b * dg + g) * K2 * P + p;
    for (int t = 0; t < K2; ++t) {
      const float off_h = live ? __ldg(og + (size_t)(2 * t) * P) : 0.f;
      const float off_w = live ? __ldg(og + (size_t)(2 * t + 1) * P) : 0.f;
      const float m = live ? __ldg(mg + (size_t)t * P) : 0.f;
      const float h_im = (float)(h_in + (t / kw) * dil) + off_h;
      const float w_im = (float)(w_in + (t % kw) * dil) + off_w;
      const bool in = live && h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
      const float hf = floorf(fminf(fmaxf(h_im, -2.f), (float)H + 1.f));
      const float wf = floorf(fminf(fmaxf(w_im, -2.f), (float)W + 1.f));
      const int h_low = (int)hf, w_low = (int)wf;
      const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
      const bool ok1 = in && h_low >= 0 && w_low >= 0, ok2 = in && h_low >= 0 && w_low + 1 <= W - 1;
      const bool ok3 = in && h_low + 1 <= H - 1 && w_low >= 0, ok4 = in && h_low + 1 <= H - 1 && w_low + 1 <= W - 1;
      const int a1 = h_low * W + w_low, a2 = a1 + 1, a3 = a1 + W, a4 = a3 + 1;
      float g_h = 0.f, g_w = 0.f, g_m = 0.f;
      for (int cc = 0; cc < cpg; ++cc) {
        const int c = g * cpg + cc;
        const float *img = xb + (size_t)c * H * W;
        const float v1 = ok1 ? __ldg(img + a1) : 0.f, v2 = ok2 ? __ldg(img + a2) : 0.f;
        const float v3 = ok3 ? __ldg(img + a3) : 0.f, v4 = ok4 ? __ldg(img + a4) : 0.f;
        const float val = (hh * hw) * v1 + (hh * lw) * v2 + (lh * hw) * v3 + (lh * lw) * v4;
        cols[(c * K2 + t) * LD + threadIdx.x] = val * m;     // forward im2col value (deform_conv_cuda.cpp:626-636)
        // d_col = sum_o W[o][c][t] * d_out[o]                (deform_conv_cuda.cpp:602-605)
        const float *wr = ws + (c * K2 + t) * CO;
        float dcol = 0.f;
#pragma unroll
        for (int o = 0; o < CO; ++o) dcol = fmaf(wr[o], dout[o], dcol);
        // d mask, d offset (dmcn_get_coordinate_weight, deform_conv_cuda_kernel.cu:463-503, 633-705)
        g_m = fmaf(dcol, val, g_m);
        const float dm = dcol * m;
        g_h = fmaf(dm, -hw * v1 - lw * v2 + hw * v3 + lw * v4, g_h);
        g_w = fmaf(dm, -hh * v1 + hh * v2 - lh * v3 + lh * v4, g_w);
        // d input: bilinear scatter (dmcn_get_gradient_weight, :434-461, 573-631), fixed-point atomics
        long long *gx = gx_fix + ((size_t)b * C + c) * H * W;
        if (ok1) atomicAdd(reinterpret_cast<unsigned long long *>(gx + a1), (unsigned long long)__double2ll_rn((double)(dm * hh * hw) * kFixScale));
        if (ok2) atomicAdd(reinterpret_cast<unsigned long long *>(gx + a2), (unsigned long long)__double2ll_rn((double)(dm * hh * lw) * kFixScale));
        if (ok3) atomicAdd(reinterpret_cast<unsigned long long *>(gx + a3), (unsigned long long)__double2ll_rn((double)(dm * lh * hw) * kFixScale));
        if (ok4) atomicAdd(reinterpret_cast<unsigned long long *>(gx + a4), (unsigned long long)__double2ll_rn((double)(dm * lh * lw) * kFixScale));
      }
      if (live) {
        float *go = goff + ((size_t)b * dg + g) * 2 * K2 * P + p;
        go[(size_t)(2 * t) * P] = g_h;
        go[(size_t)(2 * t + 1) * P] = g_w;
        gmask[(((size_t)b * dg + g) * K2 + t) * P + p] = g_m;
      }
    }
  }
  __syncthreads();
  // per-CTA partial of grad_weight[o][ck] = sum_px d_out[o][px] * col[ck][px], grad_bias[o] = sum_px d_out[o][px]
  const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int e = threadIdx.x; e < CO * CK; e += kBwThreads) {
    const int o = e / CK, ck = e % CK;
    const float *dr = douts + o * LD, *cr = cols + ck * LD;
    float acc = 0.f;
#pragma unroll 8
    for (int q = 0; q < kBwThreads; ++q) acc = fmaf(dr[q], cr[q], acc);
    gw_part[cta * CO * CK + e] = acc;
  }
  if (gb_part) {
    for (int o = threadIdx.x; o < CO; o += kBwThreads) {
      float acc = 0.f;
      for (int q = 0; q < kBwThreads; ++q) acc += douts[o * LD + q];
      gb_part[cta * CO + o] = acc;
    }
  }
}

__global__ void mdcn_bwd_finish_kernel(const float *__restrict__ gw_part, const float *__restrict__ gb_part,
                                       const long long *__restrict__ gx_fix, int ncta, int nw, int nb, long long nx,
                                       float *__restrict__ gw, float *__restrict__ gb, float *__restrict__ gx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nw) {
    float acc = 0.f;
    for (int c = 0; c < ncta; ++c) acc += __ldg(gw_part + (size_t)c * nw + e);   // fixed order
    gw[e] = acc;
  } else if (e < nw + nb) {
    const int o = (int)(e - nw);
    float acc = 0.f;
    for (int c = 0; c < ncta; ++c) acc += __ldg(gb_part + (size_t)c * nb + o);
    gb[o] = acc;
  } else if (e < nw + nb + nx) {
    const long long i = e - nw - nb;
    gx[i] = (float)((double)gx_fix[i] * (1.0 / kFixScale));
  }
}

struct BwdShape {
  int ho, wo, ncta_x;
  size_t gx_fix, gw_part, gb_part, total, smem;
};
static BwdShape bwd_shape(int b, int c, int h, int w, int cout, int kh, int kw, int stride, int pad, int dil) {
  BwdShape s{};
  s.ho = (h + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
  s.wo = (w + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
  s.ncta_x = ceil_div(s.ho * s.wo, kBwThreads);
  const size_t ncta = (size_t)s.ncta_x * b, ck = (size_t)c * kh * kw;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += align_up(bytes, 256);
    return r;
  };
  s.gx_fix = take((size_t)b * c * h * w * 8);
  s.gw_part = take(ncta * cout * ck * 4);
  s.gb_part = take(ncta * cout * 4);
  s.total = o;
  s.smem = (ck * cout + ck * (kBwThreads + 1) + (size_t)cout * (kBwThreads + 1)) * 4;
  return s;
}

}  // namespace otp

using namespace otp;

extern "C" size_t otp_mdcn_backward_workspace_bytes(int b, int c, int h, int w, int cout, int kh, int kw, int stride,
                                                    int pad, int dilation) {
  if (b <= 0 || c <= 0 || h <= 0 || w <= 0 || cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || dilation <= 0) return 0;
  return bwd_shape(b, c, h, w, cout, kh, kw, stride, pad, dilation).total;
}

extern "C" int otp_mdcn_backward(const float *x, const float *offset, const float *mask, const float *weight,
                                 const float *grad_out, float *grad_x, float *grad_offset, float *grad_mask,
                                 float *grad_weight, float *grad_bias, int b, int c, int h, int w, int cout, int kh,
                                 int kw, int stride, int pad, int dilation, int groups, int deformable_groups,
                                 void *workspace, size_t workspace_bytes, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && h > 0 && w > 0 && cout > 0 && kh > 0 && kw > 0);
  OTP_REQUIRE(stride > 0 && pad >= 0 && dilation > 0 && deformable_groups > 0 && c % deformable_groups == 0);
  if (groups != 1) {
    set_error("otp_mdcn_backward: groups=%d unsupported (OTPose uses groups=1)", groups);
    return OTP_ERR_UNSUPPORTED;
  }
  if (cout != 17) {
    set_error("otp_mdcn_backward: built for 17 output channels (got %d)", cout);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && offset && mask && weight && grad_out && grad_x && grad_offset && grad_mask && grad_weight && workspace);
  OTP_REQUIRE(b <= 65535);
  const BwdShape s = bwd_shape(b, c, h, w, cout, kh, kw, stride, pad, dilation);
  OTP_REQUIRE(s.ho > 0 && s.wo > 0);
  if (workspace_bytes < s.total) {
    set_error("otp_mdcn_backward: workspace of %zu B, need %zu B", workspace_bytes, s.total);
    return OTP_ERR_WORKSPACE;
  }
  if (s.smem > 200 * 1024) {
    set_error("otp_mdcn_backward: im2col tile of %zu B does not fit shared memory", s.smem);
    return OTP_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char *ws = static_cast<char *>(workspace);
  long long *gx_fix = reinterpret_cast<long long *>(ws + s.gx_fix);
  float *gw_part = reinterpret_cast<float *>(ws + s.gw_part);
  float *gb_part = reinterpret_cast<float *>(ws + s.gb_part);
  cudaMemsetAsync(gx_fix, 0, (size_t)b * c * h * w * 8, st);
  static PerDeviceOnce attr;
  if (attr.first() && !set_max_smem(mdcn_bwd_kernel<17>, 200 * 1024, "mdcn_bwd_kernel")) return OTP_ERR_CUDA;
  {
    LaunchScope ls(K_MDCN_BWD, st);
    mdcn_bwd_kernel<17><<<dim3(s.ncta_x, b), kBwThreads, s.smem, st>>>(
        x, offset, mask, weight, grad_out, gx_fix, grad_offset, grad_mask, gw_part, grad_bias ? gb_part : nullptr, c, h,
        w, kh, kw, stride, pad, dilation, deformable_groups, s.ho, s.wo);
  }
  {
    const int nw = cout * c * kh * kw, nb = grad_bias ? cout : 0;
    const long long nx = (long long)b * c * h * w, total = nw + nb + nx;
    LaunchScope ls(K_MDCN_BWD, st);
    mdcn_bwd_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gw_part, gb_part, gx_fix, s.ncta_x * b, nw, nb,
                                                                            nx, grad_weight, grad_bias, grad_x);
  }
  return check_launch("mdcn_bwd_kernel");
}
