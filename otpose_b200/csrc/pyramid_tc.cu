// a6 (16-bit tensor-core path): stack/view + final_layer1/2 of model/OTPose.py:362-373 -- the x2 / x4
// linear upsampling of the two branch outputs, the channel stack and the 1x1 conv 408 -> 17 -- as one
// implicit GEMM on tcgen05: out[o][t] = bias[o] + sum_k W[o][k] * A[t][k], k = (scale, channel).
//
// The CUDA-core kernel (fusion.cu) spends 51 FMAs and their shared-memory weight reads per token
// and channel (23 % of its FMA floor, 15 % of its HBM roofline).  Here a 128-token tile of the three
// sources is converted to 16 bit once -- the upsampled sources with the exact interpolation weights
// of upsample_linear1d (align_corners=False: 0.25/0.75 for x2, 0.125..0.875 for x4, border clamp) --
// into a token-contiguous (MN-major) operand tile of 8-channel x 8-token core matrices, and 26 UMMAs
// (M128 x N32 x K16) contract it with the packed weight image.  Persistent CTA per SM.
#include "common.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;
namespace {

constexpr int kPyThreads = 1024;
constexpr int kPyTM = 128;
constexpr int kPyN = 32;   // UMMA N (17 outputs padded)

struct PyShape {
  int kdim, cg;                 // 3C padded to 16, channel groups of 8
  uint32_t w_bytes, a_bytes;
  size_t smem;
};
__host__ __device__ inline PyShape py_shape(int c) {
  PyShape s;
  s.kdim = (3 * c + 15) / 16 * 16;
  s.cg = s.kdim / 8;
  s.w_bytes = (uint32_t)kPyN * s.kdim * 2;
  s.a_bytes = (uint32_t)(kPyTM / 8) * s.cg * 128;
  s.smem = (size_t)s.w_bytes + s.a_bytes;
  return s;
}

template <bool F16>
__global__ void __launch_bounds__(kPyThreads, 1)
pyramid_tc_kernel(const float *__restrict__ s0, const float *__restrict__ s1, const float *__restrict__ s2, int B, int C,
                  int T, const uint8_t *__restrict__ wimg, const float *__restrict__ bias, int cout,
                  float *__restrict__ out, long long out_bs, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[kPyN];
  const PyShape S = py_shape(C);
  uint8_t *wsm = smem;
  uint8_t *stg = smem + S.w_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T1 = T / 2, T2 = T / 4;

  for (uint32_t o = threadIdx.x * 16; o < S.w_bytes; o += kPyThreads * 16) cp_async16(wsm + o, wimg + o);
  cp_async_commit();
  if (threadIdx.x < kPyN) sbias[threadIdx.x] = (threadIdx.x < cout && bias) ? bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  uint32_t phase = 0;
  // K padding (channel groups beyond 3C): zero once, never rewritten
  for (int i = threadIdx.x; i < (kPyTM / 8) * (S.cg - 3 * (C / 8)) * 8; i += kPyThreads) {
    const int j = i / ((S.cg - 3 * (C / 8)) * 8), r = i % ((S.cg - 3 * (C / 8)) * 8);
    *reinterpret_cast<uint4 *>(stg + (size_t)j * S.cg * 128 + 3 * (C / 8) * 128 + r * 16) = make_uint4(0, 0, 0, 0);
  }

  for (int g = blockIdx.x; g < B * tiles; g += gridDim.x) {
    const int b = g / tiles, t0 = (g % tiles) * kPyTM;
    // ---- stage the [128 tokens][3C] operand tile (token-contiguous core matrices), one source at a
    //      time: item = (channel, 8-token chunk), 8 lanes = 8 consecutive channels = one core matrix;
    //      all loads of a source are issued before the first conversion ----
    constexpr int kU = (17 * 128 + kPyThreads - 1) / kPyThreads;   // items per thread and source (C <= 136)
    const int nsrc = (C / 8) * 128;                                  // C % 8 == 0
#pragma unroll 1
    for (int src = 0; src < 3; ++src) {
      float raw[kU][8];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int it = threadIdx.x + u * kPyThreads;
        const int j = (it >> 3) & 15, c = ((it >> 7) << 3) | (it & 7);
        const int t = t0 + 8 * j;   // first token of the chunk (T % 8 == 0: entirely inside or outside)
#pragma unroll
        for (int e = 0; e < 8; ++e) raw[u][e] = 0.f;
        if (it < nsrc && t < T) {
          if (src == 0) {
            const float4 *p = reinterpret_cast<const float4 *>(s0 + ((size_t)b * C + c) * T + t);
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1);
            raw[u][0] = a0.x, raw[u][1] = a0.y, raw[u][2] = a0.z, raw[u][3] = a0.w;
            raw[u][4] = a1.x, raw[u][5] = a1.y, raw[u][6] = a1.z, raw[u][7] = a1.w;
          } else if (src == 1) {
            const float *row = s1 + ((size_t)b * C + c) * T1;
            const int m0 = t >> 1;
            const float4 x = __ldg(reinterpret_cast<const float4 *>(row + m0));
            raw[u][0] = __ldg(row + max(m0 - 1, 0));
            raw[u][1] = x.x, raw[u][2] = x.y, raw[u][3] = x.z, raw[u][4] = x.w;
            raw[u][5] = __ldg(row + min(m0 + 4, T1 - 1));
          } else {
            const float *row = s2 + ((size_t)b * C + c) * T2;
            const int m0 = t >> 2;
            const float2 x = __ldg(reinterpret_cast<const float2 *>(row + m0));
            raw[u][0] = __ldg(row + max(m0 - 1, 0));
            raw[u][1] = x.x, raw[u][2] = x.y;
            raw[u][3] = __ldg(row + min(m0 + 2, T2 - 1));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int it = threadIdx.x + u * kPyThreads;
        if (it >= nsrc) continue;
        const int j = (it >> 3) & 15, cgi = it >> 7;
        float v[8];
        const float *q = raw[u];
        if (src == 1) {
          // x2: token 2m -> 0.25 * x[m-1] + 0.75 * x[m], token 2m+1 -> 0.75 * x[m] + 0.25 * x[m+1] (clamped)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[2 * i] = (1.f - 0.75f) * q[i] + 0.75f * q[i + 1];
            v[2 * i + 1] = (1.f - 0.25f) * q[i + 1] + 0.25f * q[i + 2];
          }
        } else if (src == 2) {
          // x4: tokens 4m+r -> weights 0.625, 0.875 on x[m] (with x[m-1]); 0.125, 0.375 on x[m+1] (with x[m])
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            v[4 * i] = (1.f - 0.625f) * q[i] + 0.625f * q[i + 1];
            v[4 * i + 1] = (1.f - 0.875f) * q[i] + 0.875f * q[i + 1];
            v[4 * i + 2] = (1.f - 0.125f) * q[i + 1] + 0.125f * q[i + 2];
            v[4 * i + 3] = (1.f - 0.375f) * q[i + 1] + 0.375f * q[i + 2];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = q[e];
        }
        *reinterpret_cast<uint4 *>(stg + (size_t)j * S.cg * 128 + (src * (C / 8) + cgi) * 128 + (it & 7) * 16) =
            pack16x8<F16>(v);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      const uint32_t idesc = make_idesc_16(kPyN, true, false, F16 ? 0u : 1u);
      const uint32_t sa = smem_u32(stg), sw = smem_u32(wsm);
      const uint32_t row = (uint32_t)S.cg * 128;   // A: next 8 tokens; W: next 8 output rows
      for (int s = 0; s < S.kdim / 16; ++s) {
        const uint64_t ad = make_desc(sa + s * 256, 128, row);
        const uint64_t bd = make_desc(sw + s * 256, 128, row);
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(s > 0))
            : "memory");
      }
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "elect.sync _|q, 0xffffffff;\n\t"
          "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
          ::"r"(smem_u32(&bar))
          : "memory");
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (warp < 4) {   // epilogue: lane == token, column == output channel
      const int t = t0 + warp * 32 + lane;
      float *op = out + (size_t)b * out_bs + t;
      for (int n0 = 0; n0 < kPyN; n0 += 16) {
        float acc[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, acc);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n0 + e < cout && t < T) op[(size_t)(n0 + e) * T] = acc[e] + sbias[n0 + e];
      }
    }
    tc_fence_before();
    __syncthreads();   // the staged tile and the accumulator are rewritten by the next tile
  }
  if (warp == 0) tmem_dealloc(tmem_slot, 32);
}

// weight (cout, 3C) fp32 -> image[n][k] (K-major core-matrix layout), zero padded to 32 x kdim
template <bool F16>
__global__ void pyramid_tc_pack_kernel(const float *__restrict__ w, int c, int cout, uint8_t *__restrict__ img) {
  const PyShape S = py_shape(c);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kPyN * S.kdim) return;
  const int n = e / S.kdim, k = e % S.kdim;
  const float v = (n < cout && k < 3 * c) ? w[(size_t)n * 3 * c + k] : 0.f;
  *reinterpret_cast<unsigned short *>(img + cm_offset(n, k, (S.kdim / 8) * 128, 128)) = to16<F16>(v);
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_pyramid_conv1x1_tc_supported(int c, int t, int cout) {
  if (c <= 0 || c > 136 || c % 8 != 0 || t <= 0 || cout <= 0 || cout > kPyN || t % 8 != 0) return 0;
  return py_shape(c).smem <= 200 * 1024 ? 1 : 0;
}

extern "C" size_t otp_pyramid_conv1x1_tc_pack_bytes(int c) { return c > 0 ? py_shape(c).w_bytes : 0; }

extern "C" int otp_pyramid_conv1x1_tc_pack(const float *weight, int c, int cout, int precision, void *packed,
                                           size_t packed_bytes, otp_stream_t stream) {
  OTP_REQUIRE(weight && packed && c > 0 && cout > 0 && cout <= kPyN);
  OTP_REQUIRE(precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  const PyShape S = py_shape(c);
  if (packed_bytes < S.w_bytes) {
    set_error("otp_pyramid_conv1x1_tc_pack: buffer of %zu B, need %u B", packed_bytes, S.w_bytes);
    return OTP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PACK, st);
  const int n = kPyN * S.kdim;
  if (precision == OTP_PREC_FP16)
    pyramid_tc_pack_kernel<true><<<ceil_div(n, 256), 256, 0, st>>>(weight, c, cout, static_cast<uint8_t *>(packed));
  else
    pyramid_tc_pack_kernel<false><<<ceil_div(n, 256), 256, 0, st>>>(weight, c, cout, static_cast<uint8_t *>(packed));
  return check_launch("pyramid_tc_pack_kernel");
}

extern "C" int otp_pyramid_conv1x1_tc(const float *s0, const float *s1, const float *s2, int b, int c, int t,
                                      const void *packed, const float *bias, int cout, float *out,
                                      long long out_bstride, int precision, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && c > 0 && t > 0 && cout > 0);
  OTP_REQUIRE(precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  if (!otp_pyramid_conv1x1_tc_supported(c, t, cout)) {
    set_error("otp_pyramid_conv1x1_tc: c=%d t=%d cout=%d not built (t %% 8 == 0, cout <= 32)", c, t, cout);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(s0 && s1 && s2 && packed && out);
  auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  OTP_REQUIRE(al16(s0) && al16(s1) && al16(s2));
  const PyShape S = py_shape(c);
  const int tiles = ceil_div(t, kPyTM);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr;
  if (attr.first()) {
    if (!set_max_smem(pyramid_tc_kernel<true>, 200 * 1024, "pyramid_tc_kernel") ||
        !set_max_smem(pyramid_tc_kernel<false>, 200 * 1024, "pyramid_tc_kernel"))
      return OTP_ERR_CUDA;
  }
  LaunchScope ls(K_PYRAMID, st);
  const int grid = min(b * tiles, num_sms());
  if (precision == OTP_PREC_FP16)
    pyramid_tc_kernel<true><<<grid, kPyThreads, S.smem, st>>>(s0, s1, s2, b, c, t, static_cast<const uint8_t *>(packed),
                                                              bias, cout, out, out_bstride, tiles);
  else
    pyramid_tc_kernel<false><<<grid, kPyThreads, S.smem, st>>>(s0, s1, s2, b, c, t, static_cast<const uint8_t *>(packed),
                                                               bias, cout, out, out_bstride, tiles);
  return check_launch("pyramid_tc_kernel");
}
