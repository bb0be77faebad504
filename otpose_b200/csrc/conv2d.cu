// a7 / a8: small-channel Conv2d (stride 1, k in {1,3}, dilation d, padding d*(k/2))
// with fused input add, bias (eval BatchNorm pre-folded by the host), residual and
// ReLU.  Replaces the nn.Conv2d/BatchNorm2d/ReLU triples of model/RSB.py:106-139 and
// the dilated offset / mask convs of model/OTPose.py:168-177 on the fp32 path.
//
// Direct convolution: one thread per output pixel, CO output channels in
// registers, the (Cin*k*k, CO) weight tile in shared memory read as warp
// broadcasts; input reads are coalesced row segments.  Channel counts here are
// 6..51 in / 6..306 out -- far below a tensor-core tile in K for the RSB chains;
// the 32->459 offset/mask convs get a tcgen05 implicit-GEMM path of their own.
#include "common.cuh"

namespace otp {

constexpr int kCvThreads = 128;
constexpr int kCvCO = 16;

template <int K>
__global__ void __launch_bounds__(kCvThreads)
conv2d_kernel(const float *__restrict__ x, long long x_bs, const float *__restrict__ xa, long long xa_bs,
              const float *__restrict__ weight, const float *__restrict__ bias,
              const float *__restrict__ res, long long res_bs, float *__restrict__ y, long long y_bs,
              int Cin, int H, int W, int Cout, int dil, int relu) {
  extern __shared__ float ws[];  // [Cin*K*K][kCvCO]
  constexpr int K2 = K * K;
  const int co0 = blockIdx.y * kCvCO;
  const int CK = Cin * K2;
  for (int e = threadIdx.x; e < CK * kCvCO; e += kCvThreads) {
    int o = e % kCvCO, ck = e / kCvCO;
    ws[e] = (co0 + o < Cout) ? __ldg(weight + (size_t)(co0 + o) * CK + ck) : 0.f;
  }
  __syncthreads();
  const int P = H * W;
  const int p = blockIdx.x * kCvThreads + threadIdx.x;
  if (p >= P) return;
  const int b = blockIdx.z;
  const int h = p / W, w = p % W;
  const float *__restrict__ xb = x + (size_t)b * x_bs;
  const float *__restrict__ xab = xa ? xa + (size_t)b * xa_bs : nullptr;

  float acc[kCvCO];
#pragma unroll
  for (int o = 0; o < kCvCO; ++o) acc[o] = 0.f;

#pragma unroll 4
  for (int ci = 0; ci < Cin; ++ci) {
    const float *__restrict__ xc = xb + (size_t)ci * P;
    const float *__restrict__ xac = xab ? xab + (size_t)ci * P : nullptr;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int hh = h + (i - K / 2) * dil;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int ww = w + (j - K / 2) * dil;
        float v = 0.f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          v = __ldg(xc + hh * W + ww);
          if (xac) v += __ldg(xac + hh * W + ww);
        }
        const float4 *wr = reinterpret_cast<const float4 *>(ws + (ci * K2 + i * K + j) * kCvCO);
#pragma unroll
        for (int q = 0; q < kCvCO / 4; ++q) {
          float4 w4 = wr[q];
          acc[4 * q + 0] = fmaf(w4.x, v, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(w4.y, v, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(w4.z, v, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(w4.w, v, acc[4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kCvCO; ++o) {
    const int co = co0 + o;
    if (co < Cout) {
      float v = acc[o] + (bias ? __ldg(bias + co) : 0.f);
      if (res) v += __ldg(res + (size_t)b * res_bs + (size_t)co * P + p);
      if (relu) v = fmaxf(v, 0.f);
      y[(size_t)b * y_bs + (size_t)co * P + p] = v;
    }
  }
}

// 3x3, dilation 1, few channels (the ten dense-connected branch convs of an RSB block):
// a strip of TH image rows (+1 halo row each side) of every input channel is staged in
// shared memory by coalesced loads that are all in flight at once, then each thread
// produces every output channel of its pixels from shared memory.
constexpr int kStripThreads = 256;

template <int CO>
__global__ void __launch_bounds__(kStripThreads)
conv3x3_strip_kernel(const float *__restrict__ x, long long x_bs, const float *__restrict__ xa, long long xa_bs,
                     const float *__restrict__ weight, const float *__restrict__ bias,
                     const float *__restrict__ res, long long res_bs, float *__restrict__ y, long long y_bs,
                     int Cin, int H, int W, int Cout, int TH, int relu) {
  extern __shared__ float sm[];
  const int WP = W + 2, plane = (TH + 2) * WP;
  float *xs = sm;                       // [Cin][TH+2][W+2]
  float *ws = sm + ((Cin * plane + 3) & ~3);   // [Cin*9][CO]
  const int h0 = blockIdx.x * TH, b = blockIdx.y;
  const int P = H * W;
  for (int e = threadIdx.x; e < Cin * 9 * CO; e += kStripThreads) {
    const int o = e % CO, ck = e / CO;
    ws[e] = o < Cout ? __ldg(weight + (size_t)o * Cin * 9 + ck) : 0.f;
  }
  const float *xb = x + (size_t)b * x_bs;
  const float *xab = xa ? xa + (size_t)b * xa_bs : nullptr;
  constexpr int U = 8;   // independent loads in flight per thread (a load->store loop serialises on latency)
  for (int base = threadIdx.x; base < Cin * plane; base += U * kStripThreads) {
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * kStripThreads;
      const int ci = e / plane, r = (e % plane) / WP, c = e % WP;
      const int hh = h0 - 1 + r, ww = c - 1;
      float val = 0.f;
      if (e < Cin * plane && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        val = __ldg(xb + (size_t)ci * P + hh * W + ww);
        if (xab) val += __ldg(xab + (size_t)ci * P + hh * W + ww);
      }
      v[u] = val;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = base + u * kStripThreads;
      if (e < Cin * plane) xs[e] = v[u];
    }
  }
  __syncthreads();
  const int rows = min(TH, H - h0);
  for (int pix = threadIdx.x; pix < rows * W; pix += kStripThreads) {
    const int ty = pix / W, tx = pix % W;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float *xp = xs + ci * plane + ty * WP + tx;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float v = xp[i * WP + j];
          const float4 *wr = reinterpret_cast<const float4 *>(ws + (ci * 9 + i * 3 + j) * CO);
#pragma unroll
          for (int q = 0; q < CO / 4; ++q) {
            const float4 w4 = wr[q];
            acc[4 * q + 0] = fmaf(w4.x, v, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(w4.y, v, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w4.z, v, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w4.w, v, acc[4 * q + 3]);
          }
        }
    }
    const int p = (h0 + ty) * W + tx;
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      if (o < Cout) {
        float v = acc[o] + (bias ? __ldg(bias + o) : 0.f);
        if (res) v += __ldg(res + (size_t)b * res_bs + (size_t)o * P + p);
        if (relu) v = fmaxf(v, 0.f);
        y[(size_t)b * y_bs + (size_t)o * P + p] = v;
      }
    }
  }
}

template <int CO>
static bool launch_strip(const float *x, long long x_bs, const float *xa, long long xa_bs, const float *weight,
                         const float *bias, const float *res, long long res_bs, float *y, long long y_bs, int b,
                         int cin, int h, int w, int cout, int relu, cudaStream_t st) {
  // strip height: fill the 256 threads' pixel rounds as evenly as possible
  int th = 2;
  float best = 0.f;
  for (int t = 2; t <= 8; ++t) {
    const int px = t * w;
    const float eff = (float)px / (float)(ceil_div(px, kStripThreads) * kStripThreads) + 0.01f * t;
    if (eff > best) {
      best = eff;
      th = t;
    }
  }
  const size_t smem = ((size_t)((cin * (th + 2) * (w + 2) + 3) & ~3) + (size_t)cin * 9 * CO) * sizeof(float);
  if (smem > 160 * 1024) return false;
  static size_t attr = 0;
  if (smem > attr) {
    cudaFuncSetAttribute(conv3x3_strip_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr = smem;
  }
  conv3x3_strip_kernel<CO><<<dim3(ceil_div(h, th), b), kStripThreads, smem, st>>>(
      x, x_bs, xa, xa_bs, weight, bias, res, res_bs, y, y_bs, cin, h, w, cout, th, relu);
  return true;
}

}  // namespace otp

using namespace otp;

extern "C" int otp_conv2d(const float *x, long long x_bstride, const float *x_add,
                          long long x_add_bstride, const float *weight, const float *bias,
                          const float *residual, long long residual_bstride, float *y,
                          long long y_bstride, int b, int cin, int h, int w, int cout, int k,
                          int dilation, int relu, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && cin > 0 && h > 0 && w > 0 && cout > 0 && dilation > 0 && b <= 65535);
  if (k != 1 && k != 3) {
    set_error("otp_conv2d: kernel size %d unsupported (1 or 3)", k);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && weight && y);
  size_t smem = (size_t)cin * k * k * kCvCO * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("otp_conv2d: weight tile of %zu B does not fit shared memory", smem);
    return OTP_ERR_UNSUPPORTED;
  }
  dim3 grid(ceil_div(h * w, kCvThreads), ceil_div(cout, kCvCO), b);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_CONV2D, st);
  if (k == 3 && dilation == 1 && cin <= 32 && cout <= 24 && w <= 510) {
    bool ok = false;
    if (cout <= 8)
      ok = launch_strip<8>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual, residual_bstride, y,
                           y_bstride, b, cin, h, w, cout, relu, st);
    else if (cout <= 16)
      ok = launch_strip<16>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual, residual_bstride, y,
                            y_bstride, b, cin, h, w, cout, relu, st);
    else
      ok = launch_strip<24>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual, residual_bstride, y,
                            y_bstride, b, cin, h, w, cout, relu, st);
    if (ok) return check_launch("conv3x3_strip_kernel");
  }
  if (k == 1) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv2d_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv2d_kernel<1><<<grid, kCvThreads, smem, st>>>(x, x_bstride, x_add, x_add_bstride, weight, bias,
                                                      residual, residual_bstride, y, y_bstride, cin, h,
                                                      w, cout, dilation, relu);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(conv2d_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv2d_kernel<3><<<grid, kCvThreads, smem, st>>>(x, x_bstride, x_add, x_add_bstride, weight, bias,
                                                      residual, residual_bstride, y, y_bstride, cin, h,
                                                      w, cout, dilation, relu);
  }
  return check_launch("conv2d_kernel");
}
