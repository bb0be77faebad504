// a12 (training config), native backward of the small-channel Conv2d of conv2d.cu -- the dilated
// offset / mask convs that feed the modulated DCN (reference model/OTPose.py:168-177; their backward in
// the reference is ATen's cudnn_convolution_backward, reached through autograd from
// ModulatedDeformConvFunction.backward, functions/deform_conv.py:148-167):
//
//   grad_input   = the same conv applied to grad_output with the weights transposed and flipped
//                  (stride 1, 'same' padding): no kernel of its own, the host calls otp_conv2d;
//   grad_weight[o][c][i][j] = sum_{b,h,w} grad_output[b,o,h,w] * x[b,c,h+(i-k/2)d, w+(j-k/2)d]   (this file)
//   grad_bias[o] = sum_{b,h,w} grad_output[b,o,h,w]                                              (this file)
//
// grad_weight is an implicit GEMM dY (Cout x N) . im2col(x) (N x Cin k^2) with N = B*H*W (221 k at 32 clips):
// a CTA owns a 32 x 32 (o, c) tile of ONE tap over a slice of the image rows, stages 32 x W fp32 tiles of
// dY and of the shifted x row in shared memory and accumulates 2 x 2 outputs per thread from 16-byte reads;
// the per-slice partials are summed in fixed order by a second kernel, so the result is bit-reproducible
// (the reference's cuDNN algorithms are not guaranteed to be).  fp32 CUDA cores: first correct version of
// the training path, not tuned.
#include "common.cuh"

namespace otp {
namespace {
constexpr int kWgT = 32;          // tile edge in output / input channels
constexpr int kWgPx = 128;        // pixels per staged chunk (a whole image row for W <= 128)
constexpr int kWgThreads = 256;   // 16 x 16 threads, 2 x 2 outputs each

__global__ void __launch_bounds__(kWgThreads)
conv_wgrad_kernel(const float *__restrict__ x, long long x_bs, const float *__restrict__ dy, long long dy_bs,
                  float *__restrict__ part, int B, int Cin, int H, int W, int Cout, int K, int dil, int tiles_c,
                  int rows_per_slice) {
  __shared__ __align__(16) float ds[kWgT][kWgPx + 4];   // dY tile   [o][px]
  __shared__ __align__(16) float xs[kWgT][kWgPx + 4];   // x tile    [c][px], shifted by the tap, zero padded
  const int o0 = (blockIdx.x / tiles_c) * kWgT, c0 = (blockIdx.x % tiles_c) * kWgT;
  const int tap = blockIdx.y, di = (tap / K - K / 2) * dil, dj = (tap % K - K / 2) * dil;
  const int to = threadIdx.x >> 4, tc = threadIdx.x & 15;
  const int P = H * W;
  const int row_begin = blockIdx.z * rows_per_slice, row_end = min(B * H, row_begin + rows_per_slice);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int row = row_begin; row < row_end; ++row) {
    const int b = row / H, h = row % H, hh = h + di;
    if (hh < 0 || hh >= H) continue;   // the whole shifted row is padding (uniform across the CTA)
    const float *dyr = dy + (size_t)b * dy_bs + (size_t)h * W;
    const float *xr = x + (size_t)b * x_bs + (size_t)hh * W;
    for (int w0 = 0; w0 < W; w0 += kWgPx) {
      const int npx = (min(kWgPx, W - w0) + 3) & ~3;   // staged / contracted pixels of this chunk (zero padded to 4)
      __syncthreads();
      for (int e = threadIdx.x; e < kWgT * npx; e += kWgThreads) {
        const int r = e / npx, px = e % npx, w = w0 + px, ww = w + dj;
        ds[r][px] = (o0 + r < Cout && w < W) ? __ldg(dyr + (size_t)(o0 + r) * P + w) : 0.f;
        xs[r][px] = (c0 + r < Cin && w < W && ww >= 0 && ww < W) ? __ldg(xr + (size_t)(c0 + r) * P + ww) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int px = 0; px < npx; px += 4) {
        const float4 d0 = *reinterpret_cast<const float4 *>(&ds[2 * to][px]);
        const float4 d1 = *reinterpret_cast<const float4 *>(&ds[2 * to + 1][px]);
        const float4 x0 = *reinterpret_cast<const float4 *>(&xs[2 * tc][px]);
        const float4 x1 = *reinterpret_cast<const float4 *>(&xs[2 * tc + 1][px]);
        acc[0][0] = fmaf(d0.x, x0.x, fmaf(d0.y, x0.y, fmaf(d0.z, x0.z, fmaf(d0.w, x0.w, acc[0][0]))));
        acc[0][1] = fmaf(d0.x, x1.x, fmaf(d0.y, x1.y, fmaf(d0.z, x1.z, fmaf(d0.w, x1.w, acc[0][1]))));
        acc[1][0] = fmaf(d1.x, x0.x, fmaf(d1.y, x0.y, fmaf(d1.z, x0.z, fmaf(d1.w, x0.w, acc[1][0]))));
        acc[1][1] = fmaf(d1.x, x1.x, fmaf(d1.y, x1.y, fmaf(d1.z, x1.z, fmaf(d1.w, x1.w, acc[1][1]))));
      }
    }
  }
  // partial of this slice: part[slice][o][c][tap]
  float *pp = part + (size_t)blockIdx.z * Cout * Cin * K * K;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int o = o0 + 2 * to + a, ci = c0 + 2 * tc + c;
      if (o < Cout && ci < Cin) pp[((size_t)o * Cin + ci) * K * K + tap] = acc[a][c];
    }
}

__global__ void wgrad_reduce_kernel(const float *__restrict__ part, int n, int nslice, float *__restrict__ dw,
                                    int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int k = 0; k < nslice; ++k) s += part[(size_t)k * n + e];   // fixed order
  dw[e] = accumulate ? dw[e] + s : s;
}

// grad_bias[o]: one CTA per output channel, fixed-order tree over (b, pixel)
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float *__restrict__ dy, long long dy_bs, int B, int P, float *__restrict__ out,
                   int accumulate) {
  __shared__ float red[256];
  const int o = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float *r = dy + (size_t)b * dy_bs + (size_t)o * P;
    for (int p = threadIdx.x; p < P; p += 256) s += __ldg(r + p);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[o] = accumulate ? out[o] + red[0] : red[0];
}

int wgrad_slices(int b, int cin, int h, int cout, int k) {
  const int ctas = ceil_div(cout, kWgT) * ceil_div(cin, kWgT) * k * k;
  int n = ceil_div(4 * num_sms(), ctas);
  n = n < 1 ? 1 : (n > 64 ? 64 : n);
  return n > b * h ? b * h : n;
}
}  // namespace
}  // namespace otp

using namespace otp;

extern "C" size_t otp_conv2d_wgrad_workspace_bytes(int b, int cin, int h, int w, int cout, int k) {
  if (b <= 0 || cin <= 0 || h <= 0 || w <= 0 || cout <= 0 || (k != 1 && k != 3)) return 0;
  return (size_t)wgrad_slices(b, cin, h, cout, k) * cout * cin * k * k * sizeof(float);
}

extern "C" int otp_conv2d_wgrad(const float *x, long long x_bstride, const float *grad_out, long long go_bstride,
                                float *grad_weight, float *grad_bias, int b, int cin, int h, int w, int cout, int k,
                                int dilation, int accumulate, void *workspace, size_t workspace_bytes,
                                otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && cin > 0 && h > 0 && w > 0 && cout > 0 && dilation > 0 && b <= 65535);
  if (k != 1 && k != 3) {
    set_error("otp_conv2d_wgrad: kernel size %d unsupported (1 or 3)", k);
    return OTP_ERR_UNSUPPORTED;
  }
  OTP_REQUIRE(grad_weight != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  const int n = cout * cin * k * k;
  if (b == 0) {   // empty batch: the gradients are zero (or unchanged when accumulating)
    if (!accumulate) {
      cudaMemsetAsync(grad_weight, 0, (size_t)n * sizeof(float), st);
      if (grad_bias) cudaMemsetAsync(grad_bias, 0, (size_t)cout * sizeof(float), st);
    }
    return check_launch("otp_conv2d_wgrad");
  }
  OTP_REQUIRE(x && grad_out && workspace);
  const size_t need = otp_conv2d_wgrad_workspace_bytes(b, cin, h, w, cout, k);
  if (workspace_bytes < need) {
    set_error("otp_conv2d_wgrad: workspace of %zu B, need %zu B", workspace_bytes, need);
    return OTP_ERR_WORKSPACE;
  }
  const int nslice = wgrad_slices(b, cin, h, cout, k);
  const int tiles_c = ceil_div(cin, kWgT);
  const int rps = ceil_div(b * h, nslice);
  float *part = static_cast<float *>(workspace);
  LaunchScope ls(K_CONV_BWD, st, grad_bias ? 3 : 2);
  conv_wgrad_kernel<<<dim3(ceil_div(cout, kWgT) * tiles_c, k * k, nslice), kWgThreads, 0, st>>>(
      x, x_bstride, grad_out, go_bstride, part, b, cin, h, w, cout, k, dilation, tiles_c, rps);
  wgrad_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>(part, n, nslice, grad_weight, accumulate);
  if (grad_bias) channel_sum_kernel<<<cout, 256, 0, st>>>(grad_out, go_bstride, b, h * w, grad_bias, accumulate);
  return check_launch("otp_conv2d_wgrad");
}
