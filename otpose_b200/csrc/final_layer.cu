// a0 + a1 (the backbone -> head hand-off): HRNet.final_layer, a 1x1 conv Cin -> J on
// the backbone's last feature map (reference model/HRNet.py:108-114, 150), fused with the first fusion
// pass of the head (model/OTPose.py:324-326).  The feature map is read ONCE -- fp32 NCHW as the reference's
// backbone produces it, or bf16 / fp16 channels-last as a cuDNN channels-last backbone emits it -- and the
// kernel writes rough_heatmaps (OTPose.forward returns them), total_b and squeezed.  Separately the two
// steps read the features (fp32: 4 * Cin B per pixel and frame), write rough and read it again.
//
// HBM-bound streaming kernel: one thread per TWO pixels of a clip (every weight row fetched from shared
// memory serves both), frames in the reference order cur, prev1, next1, ...; the frame sum runs left to
// right and the joint sum in channel order, so total_b / squeezed are bit-identical to otp_fusion_sum
// applied to the rough heat maps this kernel writes.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace otp {
namespace {
constexpr int kFlJ = 17;         // joints (OTPose: 17; every head kernel is built for it)
constexpr int kFlJP = 20;        // weight row stride in shared memory: 5 x float4
constexpr int kFlThreads = 128;
constexpr int kFlMaxCin = 64;    // HRNet-W32 / W48 last-stage widths: 32 / 48

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// features: NHWC == false: (frames*B, Cin, P);  NHWC == true: (frames*B, P, Cin)  (P = H*W pixels)
template <typename T, bool NHWC>
__global__ void __launch_bounds__(kFlThreads)
final_layer_sum_kernel(const T *__restrict__ feats, const float *__restrict__ weight, const float *__restrict__ bias,
                       int frames, int B, int Cin, int P, float *__restrict__ rough, float *__restrict__ total_b,
                       float *__restrict__ squeezed) {
  __shared__ __align__(16) float ws[kFlMaxCin * kFlJP];   // ws[c][j] = weight[j][c]
  __shared__ float bs[kFlJP];
  for (int e = threadIdx.x; e < Cin * kFlJP; e += kFlThreads) {
    const int c = e / kFlJP, j = e % kFlJP;
    ws[e] = j < kFlJ ? __ldg(weight + (size_t)j * Cin + c) : 0.f;
  }
  if (threadIdx.x < kFlJP) bs[threadIdx.x] = (threadIdx.x < kFlJ && bias) ? __ldg(bias + threadIdx.x) : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * (2 * kFlThreads) + threadIdx.x;
  const int pp[2] = {p0, p0 + kFlThreads};
  const bool ok[2] = {pp[0] < P, pp[1] < P};
  float tot[2][kFlJ];
  for (int f = 0; f < frames; ++f) {
    const size_t n = (size_t)f * B + b;
    float acc[2][kFlJ];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < kFlJ; ++j) acc[q][j] = bs[j];
    // 8 input channels at a time: NHWC -> one 16-byte (16-bit types) or two 16-byte (fp32) loads per
    // pixel, NCHW -> 8 coalesced row loads
    for (int c0 = 0; c0 < Cin; c0 += 8) {
      float x[2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!ok[q]) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[q][e] = 0.f;
          continue;
        }
        if constexpr (NHWC) {
          const T *src = feats + (n * P + pp[q]) * Cin + c0;
          if constexpr (sizeof(T) == 2) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(src));
            const T *h = reinterpret_cast<const T *>(&raw);
#pragma unroll
            for (int e = 0; e < 8; ++e) x[q][e] = to_f32<T>(h[e]);
          } else {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(src));
            const float4 c = __ldg(reinterpret_cast<const float4 *>(src) + 1);
            x[q][0] = a.x, x[q][1] = a.y, x[q][2] = a.z, x[q][3] = a.w;
            x[q][4] = c.x, x[q][5] = c.y, x[q][6] = c.z, x[q][7] = c.w;
          }
        } else {
          const T *src = feats + (n * Cin + c0) * P + pp[q];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[q][e] = to_f32<T>(__ldg(src + (size_t)e * P));
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 *wr = reinterpret_cast<const float4 *>(ws + (c0 + e) * kFlJP);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 w4 = wr[g];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            acc[q][4 * g] = fmaf(w4.x, x[q][e], acc[q][4 * g]);
            acc[q][4 * g + 1] = fmaf(w4.y, x[q][e], acc[q][4 * g + 1]);
            acc[q][4 * g + 2] = fmaf(w4.z, x[q][e], acc[q][4 * g + 2]);
            acc[q][4 * g + 3] = fmaf(w4.w, x[q][e], acc[q][4 * g + 3]);
          }
        }
        const float w16 = ws[(c0 + e) * kFlJP + 16];
#pragma unroll
        for (int q = 0; q < 2; ++q) acc[q][16] = fmaf(w16, x[q][e], acc[q][16]);
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
      float *r = rough + n * kFlJ * P + pp[q];
#pragma unroll
      for (int j = 0; j < kFlJ; ++j) {
        r[(size_t)j * P] = acc[q][j];
        tot[q][j] = f == 0 ? acc[q][j] : tot[q][j] + acc[q][j];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (!ok[q]) continue;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kFlJ; ++j) {
      total_b[((size_t)b * kFlJ + j) * P + pp[q]] = tot[q][j];
      sq += tot[q][j];
    }
    squeezed[(size_t)b * P + pp[q]] = sq;
  }
}
}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_final_layer_fusion_sum(const void *features, int feat_dtype, int channels_last,
                                          const float *weight, const float *bias, int frames, int b, int cin,
                                          int joints, int t, float *rough, float *total_b, float *squeezed,
                                          otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && t > 0 && b <= 65535);
  OTP_REQUIRE(frames == 3 || frames == 5 || frames == 7);
  OTP_REQUIRE(feat_dtype == OTP_PREC_FP32 || feat_dtype == OTP_PREC_BF16 || feat_dtype == OTP_PREC_FP16);
  if (joints != kFlJ || cin <= 0 || cin > kFlMaxCin || cin % 8 != 0) {
    set_error("otp_final_layer_fusion_sum: built for 17 joints and Cin %% 8 == 0, Cin <= %d (got J=%d, Cin=%d)",
              kFlMaxCin, joints, cin);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(features && weight && rough && total_b && squeezed);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(features) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ceil_div(t, 2 * kFlThreads), b);
  LaunchScope ls(K_FINAL_LAYER, st);
#define OTP_FL(T, NHWC)                                                                                      \
  final_layer_sum_kernel<T, NHWC><<<grid, kFlThreads, 0, st>>>(static_cast<const T *>(features), weight, bias, \
                                                               frames, b, cin, t, rough, total_b, squeezed)
  if (feat_dtype == OTP_PREC_FP32) {
    if (channels_last) OTP_FL(float, true); else OTP_FL(float, false);
  } else if (feat_dtype == OTP_PREC_BF16) {
    if (channels_last) OTP_FL(__nv_bfloat16, true); else OTP_FL(__nv_bfloat16, false);
  } else {
    if (channels_last) OTP_FL(__half, true); else OTP_FL(__half, false);
  }
#undef OTP_FL
  return check_launch("final_layer_sum_kernel");
}
