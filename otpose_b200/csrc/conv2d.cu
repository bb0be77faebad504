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

// 1x1 conv, four consecutive pixels per thread (128-bit loads / stores) x 16 output channels:
// one LDG.128 + four weight LDS.128 feed 64 FMAs, so the loop is FMA-bound.
constexpr int kV1Threads = 128;

__global__ void __launch_bounds__(kV1Threads)
conv1x1_vec_kernel(const float *__restrict__ x, long long x_bs, const float *__restrict__ xa, long long xa_bs,
                   const float *__restrict__ weight, const float *__restrict__ bias,
                   const float *__restrict__ res, long long res_bs, float *__restrict__ y, long long y_bs, int Cin,
                   int P, int Cout, int relu) {
  extern __shared__ float ws[];  // [Cin][kCvCO]
  const int co0 = blockIdx.y * kCvCO, b = blockIdx.z;
  for (int e = threadIdx.x; e < Cin * kCvCO; e += kV1Threads) {
    const int o = e % kCvCO, ci = e / kCvCO;
    ws[e] = (co0 + o < Cout) ? __ldg(weight + (size_t)(co0 + o) * Cin + ci) : 0.f;
  }
  __syncthreads();
  const int p4 = blockIdx.x * kV1Threads + threadIdx.x;
  if (4 * p4 >= P) return;
  const float4 *xb = reinterpret_cast<const float4 *>(x + (size_t)b * x_bs) + p4;
  const float4 *xab = xa ? reinterpret_cast<const float4 *>(xa + (size_t)b * xa_bs) + p4 : nullptr;
  const int P4 = P / 4;
  float acc[4][kCvCO];
#pragma unroll
  for (int o = 0; o < kCvCO; ++o) acc[0][o] = acc[1][o] = acc[2][o] = acc[3][o] = 0.f;
#pragma unroll 4
  for (int ci = 0; ci < Cin; ++ci) {
    float4 v = __ldg(xb + (size_t)ci * P4);
    if (xab) {
      const float4 a = __ldg(xab + (size_t)ci * P4);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    const float4 *wr = reinterpret_cast<const float4 *>(ws + ci * kCvCO);
#pragma unroll
    for (int q = 0; q < kCvCO / 4; ++q) {
      const float4 w4 = wr[q];
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[0][4 * q + e] = fmaf(wv[e], v.x, acc[0][4 * q + e]);
        acc[1][4 * q + e] = fmaf(wv[e], v.y, acc[1][4 * q + e]);
        acc[2][4 * q + e] = fmaf(wv[e], v.z, acc[2][4 * q + e]);
        acc[3][4 * q + e] = fmaf(wv[e], v.w, acc[3][4 * q + e]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kCvCO; ++o) {
    const int co = co0 + o;
    if (co < Cout) {
      const float bo = bias ? __ldg(bias + co) : 0.f;
      float4 r = make_float4(acc[0][o] + bo, acc[1][o] + bo, acc[2][o] + bo, acc[3][o] + bo);
      if (res) {
        const float4 rr = __ldg(reinterpret_cast<const float4 *>(res + (size_t)b * res_bs + (size_t)co * P) + p4);
        r.x += rr.x; r.y += rr.y; r.z += rr.z; r.w += rr.w;
      }
      if (relu) {
        r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
      }
      reinterpret_cast<float4 *>(y + (size_t)b * y_bs + (size_t)co * P)[p4] = r;
    }
  }
}

// 3x3, dilation 1, few channels (the ten dense-connected branch convs of an RSB block):
// a strip of TH image rows (+1 halo row each side) of every input channel is staged in
// shared memory by coalesced loads that are all in flight at once, then each thread
// produces every output channel of its pixels from shared memory.
constexpr int kStripThreads = 160;   // >= 4 x 36 blocks of 2x2 pixels in an 8-row strip of a 72-wide map

template <int CO>
__global__ void __launch_bounds__(kStripThreads, 3)
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
  // staging: one (channel, row) line per warp iteration, four lines in flight per warp (a plain
  // load->store loop would serialise on load latency); no per-element integer division
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nlines = Cin * (TH + 2);
    constexpr int U = 4, NC = 3;   // 4 lines x 3 column rounds = 12 (24 with x_add) loads in flight per thread
    for (int l0 = warp; l0 < nlines; l0 += U * (kStripThreads / 32)) {
      for (int cb = 0; cb < WP; cb += 32 * NC) {
        float v[U][NC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int line = l0 + u * (kStripThreads / 32);
          const int ci = line / (TH + 2), r = line - ci * (TH + 2);
          const int hh = h0 - 1 + r;
          const bool lok = line < nlines && hh >= 0 && hh < H;
          const size_t ro = (size_t)ci * P + hh * W;
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int ww = cb + lane + 32 * k - 1;
            float val = 0.f;
            if (lok && ww >= 0 && ww < W) {
              val = __ldg(xb + ro + ww);
              if (xab) val += __ldg(xab + ro + ww);
            }
            v[u][k] = val;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int line = l0 + u * (kStripThreads / 32);
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int c0 = cb + lane + 32 * k;
            if (line < nlines && c0 < WP) xs[line * WP + c0] = v[u][k];
          }
        }
      }
    }
  }
  __syncthreads();
  // Each thread produces a 2x2 block of pixels.  The weight vectors (warp-broadcast LDS.128, the
  // scarce resource: a 128-bit shared load costs four LSU cycles) are then shared by four
  // pixels and the 4x4 input window by nine taps, which balances the LSU against the FMA pipe.
  const int rows = min(TH, H - h0), WH = (W + 1) / 2, RH = (rows + 1) / 2;
  for (int item = threadIdx.x; item < RH * WH; item += kStripThreads) {
    const int ty = 2 * (item / WH), tx = 2 * (item % WH);
    const bool c1 = tx + 1 < W, r1 = ty + 1 < rows;
    float acc[4][CO];   // pixel (dy, dx) -> acc[2*dy + dx]
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[0][o] = acc[1][o] = acc[2][o] = acc[3][o] = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float *xp = xs + ci * plane + ty * WP + tx;
      float win[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) win[i][j] = ((i < 3 || r1) && (j < 3 || c1)) ? xp[i * WP + j] : 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 *wr = reinterpret_cast<const float4 *>(ws + (ci * 9 + i * 3 + j) * CO);
#pragma unroll
          for (int q = 0; q < CO / 4; ++q) {
            const float4 w4 = wr[q];
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc[0][4 * q + e] = fmaf(wv[e], win[i][j], acc[0][4 * q + e]);
              acc[1][4 * q + e] = fmaf(wv[e], win[i][j + 1], acc[1][4 * q + e]);
              acc[2][4 * q + e] = fmaf(wv[e], win[i + 1][j], acc[2][4 * q + e]);
              acc[3][4 * q + e] = fmaf(wv[e], win[i + 1][j + 1], acc[3][4 * q + e]);
            }
          }
        }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
      if (o < Cout) {
        const float bo = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (((k & 1) == 0 || c1) && ((k & 2) == 0 || r1)) {
            const int p = (h0 + ty + (k >> 1)) * W + tx + (k & 1);
            float v = acc[k][o] + bo;
            if (res) v += __ldg(res + (size_t)b * res_bs + (size_t)o * P + p);
            if (relu) v = fmaxf(v, 0.f);
            y[(size_t)b * y_bs + (size_t)o * P + p] = v;
          }
        }
      }
    }
  }
}

template <int CO>
static bool launch_strip(const float *x, long long x_bs, const float *xa, long long xa_bs, const float *weight,
                         const float *bias, const float *res, long long res_bs, float *y, long long y_bs, int b,
                         int cin, int h, int w, int cout, int relu, cudaStream_t st) {
  const int th = 8;   // strip height (2x2 pixel blocks per thread; the halo costs 10/8)
  const size_t smem = ((size_t)((cin * (th + 2) * (w + 2) + 3) & ~3) + (size_t)cin * 9 * CO) * sizeof(float);
  if (smem > 160 * 1024) return false;
  static size_t attr[kMaxDevices] = {};   // largest opt-in so far, per device
  const int dev = current_device();
  if (smem > attr[dev]) {
    if (!set_max_smem(conv3x3_strip_kernel<CO>, smem, "conv3x3_strip_kernel")) return false;
    attr[dev] = smem;
  }
  conv3x3_strip_kernel<CO><<<dim3(ceil_div(h, th), b), kStripThreads, smem, st>>>(
      x, x_bs, xa, xa_bs, weight, bias, res, res_bs, y, y_bs, cin, h, w, cout, th, relu);
  return true;
}

}  // namespace otp

using namespace otp;

__global__ void conv_bn_fold_kernel(const float *__restrict__ w, const float *__restrict__ b, const float *__restrict__ gamma,
                                    const float *__restrict__ beta, const float *__restrict__ mean,
                                    const float *__restrict__ var, float eps, int has_bn, int cout, int per_out,
                                    float *__restrict__ w_out, float *__restrict__ b_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cout * per_out + cout) return;
  const int o = e < cout * per_out ? e / per_out : e - cout * per_out;
  const float g = has_bn ? gamma[o] / sqrtf(var[o] + eps) : 1.f;
  if (e < cout * per_out) {
    w_out[e] = w[e] * g;
  } else {
    const float bb = b ? b[o] : 0.f;
    b_out[o] = has_bn ? (bb - mean[o]) * g + beta[o] : bb;
  }
}

extern "C" int otp_conv_bn_fold(const float *weight, const float *bias, const float *gamma, const float *beta,
                                const float *running_mean, const float *running_var, float eps, int has_bn, int cout,
                                int per_out, float *weight_out, float *bias_out, otp_stream_t stream) {
  OTP_REQUIRE(cout > 0 && per_out > 0 && weight && weight_out && bias_out);
  OTP_REQUIRE(!has_bn || (gamma && beta && running_mean && running_var));
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PACK, st);
  const int n = cout * per_out + cout;
  conv_bn_fold_kernel<<<ceil_div(n, 256), 256, 0, st>>>(weight, bias, gamma, beta, running_mean, running_var, eps, has_bn,
                                                        cout, per_out, weight_out, bias_out);
  return check_launch("conv_bn_fold_kernel");
}

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
    else if (cout <= 20)
      ok = launch_strip<20>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual, residual_bstride, y,
                            y_bstride, b, cin, h, w, cout, relu, st);
    else
      ok = launch_strip<24>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual, residual_bstride, y,
                            y_bstride, b, cin, h, w, cout, relu, st);
    if (ok) return check_launch("conv3x3_strip_kernel");
  }
  const int P = h * w;
  auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (k == 1 && P % 4 == 0 && x_bstride % 4 == 0 && y_bstride % 4 == 0 && al16(x) && al16(y) &&
      (!x_add || (x_add_bstride % 4 == 0 && al16(x_add))) && (!residual || (residual_bstride % 4 == 0 && al16(residual))) &&
      smem <= 48 * 1024) {
    dim3 g4(ceil_div(P / 4, kV1Threads), ceil_div(cout, kCvCO), b);
    conv1x1_vec_kernel<<<g4, kV1Threads, smem, st>>>(x, x_bstride, x_add, x_add_bstride, weight, bias, residual,
                                                      residual_bstride, y, y_bstride, cin, P, cout, relu);
    return check_launch("conv1x1_vec_kernel");
  }
  if (k == 1) {
    if (smem > 48 * 1024 && !set_max_smem(conv2d_kernel<1>, smem, "conv2d_kernel")) return OTP_ERR_CUDA;
    conv2d_kernel<1><<<grid, kCvThreads, smem, st>>>(x, x_bstride, x_add, x_add_bstride, weight, bias,
                                                      residual, residual_bstride, y, y_bstride, cin, h,
                                                      w, cout, dilation, relu);
  } else {
    if (smem > 48 * 1024 && !set_max_smem(conv2d_kernel<3>, smem, "conv2d_kernel")) return OTP_ERR_CUDA;
    conv2d_kernel<3><<<grid, kCvThreads, smem, st>>>(x, x_bstride, x_add, x_add_bstride, weight, bias,
                                                      residual, residual_bstride, y, y_bstride, cin, h,
                                                      w, cout, dilation, relu);
  }
  return check_launch("conv2d_kernel");
}
