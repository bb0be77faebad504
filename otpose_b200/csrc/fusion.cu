// a1 / a6 and the encoder glue ops: fusion prologue, positional embedding add,
// linear upsampling, and the 3-scale pyramid 1x1 conv.  All HBM-bound streaming
// kernels: one thread per (clip, token), channel loop inside, so every global
// access is a coalesced row segment of a (B,C,T) tensor.
#include "common.cuh"

namespace otp {

constexpr int kFuThreads = 256;

// model/OTPose.py:324-326 -- total_b = cur+prev+next+pprev+nnext, squeezed = sum_j total_b.
// NP = number of (prev_k, next_k) frame pairs around the current frame: the reference
// hard-codes NP = 2 (`supplement = 5`, OTPose.py:188, 317-321); NP = 1 / 3 are the T=3 / T=7
// window extension of BASELINE config 5, frames ordered cur, prev1, next1, prev2, next2, ...
// and summed left to right, so NP = 2 is the reference's expression bit for bit.
template <int NP>
__global__ void __launch_bounds__(kFuThreads)
fusion_sum_kernel(const float *__restrict__ rough, int B, int J, int T, float *__restrict__ total_b,
                  float *__restrict__ squeezed) {
  const int t = blockIdx.x * kFuThreads + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const size_t fs = (size_t)B * J * T;  // frame stride
  const float *__restrict__ r = rough + (size_t)b * J * T + t;
  float sq = 0.f;
  for (int j = 0; j < J; ++j) {
    const float *rj = r + (size_t)j * T;
    float v[2 * NP + 1];
#pragma unroll
    for (int f = 0; f < 2 * NP + 1; ++f) v[f] = __ldg(rj + f * fs);
    float tot = v[0];
#pragma unroll
    for (int f = 1; f < 2 * NP + 1; ++f) tot += v[f];
    total_b[((size_t)b * J + j) * T + t] = tot;
    sq += tot;
  }
  squeezed[(size_t)b * T + t] = sq;
}

// model/OTPose.py:330, 339-359 (+ pos_embd add of ConvVideoTransformer.py:140-155)
//   prev_b  = cur + (prev1' + prev2' + ...)        next_b = cur + (next1' + next2' + ...)
//   close_b = cur + (next1' + prev1')              far_b  = cur + sum_{k>=2} (next_k' + prev_k')
// with x' = x / (margin + 1); NP = 2 gives the reference's lines 343-354, NP = 1 leaves far_b = cur.
template <int NP>
__global__ void __launch_bounds__(kFuThreads)
fusion_stack_kernel(const float *__restrict__ rough, const int64_t *__restrict__ margin,
                    const float *__restrict__ squeezed, const float *__restrict__ context,
                    const float *__restrict__ pe1, const float *__restrict__ pe2, int pe_stride, int B,
                    int J, int T, float *__restrict__ x1, float *__restrict__ x2,
                    float *__restrict__ intersection, float *__restrict__ prev_b_out) {
  const int t = blockIdx.x * kFuThreads + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const size_t fs = (size_t)B * J * T;
  float den[2 * NP];
#pragma unroll
  for (int f = 0; f < 2 * NP; ++f) den[f] = (float)(margin[2 * NP * b + f] + 1);
  const float sq = __ldg(squeezed + (size_t)b * T + t);
  for (int j = 0; j < J; ++j) {
    const size_t o17 = ((size_t)b * J + j) * T + t;
    const float *rj = rough + o17;
    float v[2 * NP + 1];
#pragma unroll
    for (int f = 0; f < 2 * NP + 1; ++f) v[f] = __ldg(rj + f * fs);
    const float cur = v[0];
    float tot = cur;
#pragma unroll
    for (int f = 1; f < 2 * NP + 1; ++f) tot += v[f];
    const float inter = tot * sq;
    const float ctx = __ldg(context + o17);
#pragma unroll
    for (int f = 0; f < 2 * NP; ++f) v[f + 1] = v[f + 1] / den[f];
    // v[2k-1] = prev_k', v[2k] = next_k'
    float ps = v[1], ns = v[2];
#pragma unroll
    for (int k = 2; k <= NP; ++k) {
      ps += v[2 * k - 1];
      ns += v[2 * k];
    }
    const float prev_b = cur + ps, next_b = cur + ns;
    const float close_b = cur + (v[2] + v[1]);
    float far_b = cur;
    if (NP >= 2) {
      float fsum = v[4] + v[3];
#pragma unroll
      for (int k = 3; k <= NP; ++k) fsum += v[2 * k] + v[2 * k - 1];
      far_b = cur + fsum;
    }
    float v1[8] = {inter, ctx, prev_b, far_b, close_b, prev_b * sq, far_b * sq, close_b * sq};
    float v2[8] = {inter, ctx, next_b, close_b, far_b, next_b * sq, close_b * sq, far_b * sq};
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int c = j * 8 + m;
      const size_t o = ((size_t)b * J * 8 + c) * T + t;
      x1[o] = v1[m] + (pe1 ? __ldg(pe1 + (size_t)c * pe_stride + t) : 0.f);
      x2[o] = v2[m] + (pe2 ? __ldg(pe2 + (size_t)c * pe_stride + t) : 0.f);
    }
    if (intersection) intersection[o17] = inter;
    if (prev_b_out) prev_b_out[o17] = prev_b;
  }
}

__global__ void __launch_bounds__(kFuThreads)
add_pe_kernel(const float *__restrict__ x, const float *__restrict__ pe, int pe_stride,
              float *__restrict__ y, int C, int T) {
  const int t = blockIdx.x * kFuThreads + threadIdx.x;
  if (t >= T) return;
  const int c = blockIdx.y, b = blockIdx.z;
  const size_t o = ((size_t)b * C + c) * T + t;
  y[o] = __ldg(x + o) + __ldg(pe + (size_t)c * pe_stride + t);
}

// upsample_linear1d, align_corners=False, integer scale factor
__device__ __forceinline__ void lerp_index(int d, float inv_scale, int L, int &i0, int &i1, float &l1) {
  float src = inv_scale * ((float)d + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 + (i0 < L - 1 ? 1 : 0);
  l1 = src - (float)i0;
}

__global__ void __launch_bounds__(kFuThreads)
upsample_kernel(const float *__restrict__ x, float *__restrict__ y, int Tin, int scale) {
  const int Tout = Tin * scale;
  const int d = blockIdx.x * kFuThreads + threadIdx.x;
  if (d >= Tout) return;
  const size_t row = (size_t)blockIdx.z * gridDim.y + blockIdx.y;
  int i0, i1;
  float l1;
  lerp_index(d, 1.0f / (float)scale, Tin, i0, i1, l1);
  const float *xr = x + row * Tin;
  y[row * Tout + d] = (1.f - l1) * __ldg(xr + i0) + l1 * __ldg(xr + i1);
}

// model/OTPose.py:362-373: out[o,t] = bias[o] + sum_{s,c} W[o, s*C+c] * up_s(src_s)[c,t]
// Weight rows are padded to a multiple of 4 outputs and read as float4 broadcasts, and every thread
// produces TWO tokens (t and t + 256) per weight read: the kernel is bound by the shared-memory
// weight traffic per FMA, not by HBM.
template <int CO>
__global__ void __launch_bounds__(kFuThreads)
pyramid_conv_kernel(const float *__restrict__ s0, const float *__restrict__ s1,
                    const float *__restrict__ s2, int C, int T, int T1, int T2,
                    const float *__restrict__ weight, const float *__restrict__ bias, int Cout,
                    float *__restrict__ out, long long out_bstride) {
  constexpr int LD = (CO + 3) / 4 * 4;
  extern __shared__ __align__(16) float ws[];  // [3C][LD]
  const int co0 = blockIdx.z * CO;
  for (int e = threadIdx.x; e < 3 * C * LD; e += kFuThreads) {
    int o = e % LD, k = e / LD;
    ws[e] = (o < CO && co0 + o < Cout) ? __ldg(weight + (size_t)(co0 + o) * 3 * C + k) : 0.f;
  }
  __syncthreads();
  const int b = blockIdx.y;
  int tk[2];
  tk[0] = blockIdx.x * 2 * kFuThreads + threadIdx.x;
  tk[1] = tk[0] + kFuThreads;
  if (tk[0] >= T) return;
  const bool two = tk[1] < T;
  float acc[2][LD];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int o = 0; o < LD; ++o) acc[u][o] = 0.f;
  int a0[2], a1[2], b0[2], b1[2];
  float la[2], lb[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int t = two || u == 0 ? tk[u] : tk[0];
    lerp_index(t, 0.5f, T1, a0[u], a1[u], la[u]);
    lerp_index(t, 0.25f, T2, b0[u], b1[u], lb[u]);
    tk[u] = t;
  }
  const float *p0 = s0 + (size_t)b * C * T;
  const float *p1 = s1 + (size_t)b * C * T1;
  const float *p2 = s2 + (size_t)b * C * T2;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    float v0[2], v1[2], v2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      v0[u] = __ldg(p0 + tk[u]);
      v1[u] = (1.f - la[u]) * __ldg(p1 + a0[u]) + la[u] * __ldg(p1 + a1[u]);
      v2[u] = (1.f - lb[u]) * __ldg(p2 + b0[u]) + lb[u] * __ldg(p2 + b1[u]);
    }
    p0 += T;
    p1 += T1;
    p2 += T2;
    const float4 *w0 = reinterpret_cast<const float4 *>(ws + c * LD);
    const float4 *w1 = reinterpret_cast<const float4 *>(ws + (C + c) * LD);
    const float4 *w2 = reinterpret_cast<const float4 *>(ws + (2 * C + c) * LD);
#pragma unroll
    for (int q = 0; q < LD / 4; ++q) {
      const float4 x0 = w0[q], x1 = w1[q], x2 = w2[q];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        acc[u][4 * q] = fmaf(x0.x, v0[u], fmaf(x1.x, v1[u], fmaf(x2.x, v2[u], acc[u][4 * q])));
        acc[u][4 * q + 1] = fmaf(x0.y, v0[u], fmaf(x1.y, v1[u], fmaf(x2.y, v2[u], acc[u][4 * q + 1])));
        acc[u][4 * q + 2] = fmaf(x0.z, v0[u], fmaf(x1.z, v1[u], fmaf(x2.z, v2[u], acc[u][4 * q + 2])));
        acc[u][4 * q + 3] = fmaf(x0.w, v0[u], fmaf(x1.w, v1[u], fmaf(x2.w, v2[u], acc[u][4 * q + 3])));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (u == 1 && !two) break;
#pragma unroll
    for (int o = 0; o < CO; ++o)
      if (co0 + o < Cout)
        out[(size_t)b * out_bstride + (size_t)(co0 + o) * T + tk[u]] = acc[u][o] + (bias ? __ldg(bias + co0 + o) : 0.f);
  }
}

}  // namespace otp

using namespace otp;

extern "C" int otp_fusion_sum_frames(const float *rough, int frames, int b, int j, int t, float *total_b,
                                     float *squeezed, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && j > 0 && t > 0 && b <= 65535);
  if (frames != 3 && frames != 5 && frames != 7) {
    set_error("otp_fusion_sum_frames: window of %d frames not built (3, 5 or 7)", frames);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(rough && total_b && squeezed);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ceil_div(t, kFuThreads), b);
  LaunchScope ls(K_FUSION_SUM, st);
  if (frames == 3) fusion_sum_kernel<1><<<grid, kFuThreads, 0, st>>>(rough, b, j, t, total_b, squeezed);
  else if (frames == 5) fusion_sum_kernel<2><<<grid, kFuThreads, 0, st>>>(rough, b, j, t, total_b, squeezed);
  else fusion_sum_kernel<3><<<grid, kFuThreads, 0, st>>>(rough, b, j, t, total_b, squeezed);
  return check_launch("fusion_sum_kernel");
}

extern "C" int otp_fusion_sum(const float *rough, int b, int j, int t, float *total_b, float *squeezed,
                              otp_stream_t stream) {
  return otp_fusion_sum_frames(rough, 5, b, j, t, total_b, squeezed, stream);
}

extern "C" int otp_fusion_stack_frames(const float *rough, const int64_t *margin, const float *squeezed,
                                       const float *context, const float *pe1, const float *pe2,
                                       int pe_stride, int frames, int b, int j, int t, float *x1, float *x2,
                                       float *intersection, float *prev_b, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && j > 0 && t > 0 && b <= 65535);
  if (frames != 3 && frames != 5 && frames != 7) {
    set_error("otp_fusion_stack_frames: window of %d frames not built (3, 5 or 7)", frames);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(rough && margin && squeezed && context && x1 && x2);
  OTP_REQUIRE((!pe1 && !pe2) || pe_stride >= t);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ceil_div(t, kFuThreads), b);
  LaunchScope ls(K_FUSION_STACK, st);
#define OTP_STACK(NP)                                                                                         \
  fusion_stack_kernel<NP><<<grid, kFuThreads, 0, st>>>(rough, margin, squeezed, context, pe1, pe2, pe_stride, \
                                                       b, j, t, x1, x2, intersection, prev_b)
  if (frames == 3) OTP_STACK(1);
  else if (frames == 5) OTP_STACK(2);
  else OTP_STACK(3);
#undef OTP_STACK
  return check_launch("fusion_stack_kernel");
}

extern "C" int otp_fusion_stack(const float *rough, const int64_t *margin, const float *squeezed,
                                const float *context, const float *pe1, const float *pe2, int pe_stride,
                                int b, int j, int t, float *x1, float *x2, float *intersection,
                                float *prev_b, otp_stream_t stream) {
  return otp_fusion_stack_frames(rough, margin, squeezed, context, pe1, pe2, pe_stride, 5, b, j, t, x1, x2,
                                 intersection, prev_b, stream);
}

extern "C" int otp_add_pos_embd(const float *x, const float *pe, int pe_stride, float *y, int b, int c,
                                int t, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && t > 0 && b <= 65535 && c <= 65535 && pe_stride >= t);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && pe && y);
  LaunchScope ls(K_ADD_PE, (cudaStream_t)stream);
  add_pe_kernel<<<dim3(ceil_div(t, kFuThreads), c, b), kFuThreads, 0, (cudaStream_t)stream>>>(
      x, pe, pe_stride, y, c, t);
  return check_launch("add_pe_kernel");
}

extern "C" int otp_upsample_linear(const float *x, float *y, int b, int c, int t_in, int scale,
                                   otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && t_in > 0 && scale >= 1 && b <= 65535 && c <= 65535);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && y);
  LaunchScope ls(K_UPSAMPLE, (cudaStream_t)stream);
  upsample_kernel<<<dim3(ceil_div(t_in * scale, kFuThreads), c, b), kFuThreads, 0,
                    (cudaStream_t)stream>>>(x, y, t_in, scale);
  return check_launch("upsample_kernel");
}

extern "C" int otp_pyramid_conv1x1(const float *s0, const float *s1, const float *s2, int b, int c,
                                   int t, int t1, int t2, const float *weight, const float *bias,
                                   int cout, float *out, long long out_bstride, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && t > 0 && cout > 0 && b <= 65535);
  OTP_REQUIRE(t1 * 2 == t && t2 * 4 == t);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(s0 && s1 && s2 && weight && out);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PYRAMID, st);
  if (cout % 17 == 0) {
    size_t smem = (size_t)3 * c * 20 * sizeof(float);
    if (smem > 48 * 1024 && !set_max_smem(pyramid_conv_kernel<17>, smem, "pyramid_conv_kernel")) return OTP_ERR_CUDA;
    pyramid_conv_kernel<17><<<dim3(ceil_div(t, 2 * kFuThreads), b, cout / 17), kFuThreads, smem, st>>>(
        s0, s1, s2, c, t, t1, t2, weight, bias, cout, out, out_bstride);
  } else {
    size_t smem = (size_t)3 * c * 16 * sizeof(float);
    if (smem > 48 * 1024 && !set_max_smem(pyramid_conv_kernel<16>, smem, "pyramid_conv_kernel")) return OTP_ERR_CUDA;
    pyramid_conv_kernel<16><<<dim3(ceil_div(t, 2 * kFuThreads), b, ceil_div(cout, 16)), kFuThreads, smem, st>>>(
        s0, s1, s2, c, t, t1, t2, weight, bias, cout, out, out_bstride);
  }
  return check_launch("pyramid_conv_kernel");
}
