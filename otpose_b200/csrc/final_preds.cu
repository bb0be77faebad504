// a11: heat-map argmax + quarter-pixel refinement + back-projection.
// Replaces the NumPy path utils/heatmap.py:108-171 (+ utils/transform.py:76-126).
//
// One CTA per (sample, joint) map: a single coalesced 128-bit pass over the
// H*W floats, per-thread (max, first index) kept in registers, warp-shuffle
// then cross-warp reduction.  HBM-bound: algorithmic bytes = H*W*4 per map
// read + 28 B written.
#include <math_constants.h>

#include "common.cuh"

namespace otp {

struct MaxIdx {
  float v;
  int i;
};

__device__ __forceinline__ void take(MaxIdx &a, float v, int i) {
  // np.argmax semantics: strictly larger wins, equal keeps the lower index
  if (v > a.v || (v == a.v && i < a.i)) {
    a.v = v;
    a.i = i;
  }
}

__device__ __forceinline__ MaxIdx warp_argmax(MaxIdx a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float v = __shfl_xor_sync(0xffffffffu, a.v, o);
    int i = __shfl_xor_sync(0xffffffffu, a.i, o);
    take(a, v, i);
  }
  return a;
}

constexpr int kFpThreads = 256;

__global__ void __launch_bounds__(kFpThreads)
final_preds_kernel(const float *__restrict__ hm, int J, int H, int W, const float *__restrict__ center,
                   const float *__restrict__ scale, int32_t *__restrict__ out_idx,
                   float *__restrict__ out_coords, float *__restrict__ out_preds,
                   float *__restrict__ out_maxvals) {
  const int map = blockIdx.x;  // n*J + j
  const int T = H * W;
  const float *__restrict__ p = hm + (size_t)map * T;
  MaxIdx best{-CUDART_INF_F, 0x7fffffff};

  if ((T & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const float4 *p4 = reinterpret_cast<const float4 *>(p);
    for (int i = threadIdx.x; i < (T >> 2); i += kFpThreads) {
      float4 v = __ldg(p4 + i);
      take(best, v.x, 4 * i);
      take(best, v.y, 4 * i + 1);
      take(best, v.z, 4 * i + 2);
      take(best, v.w, 4 * i + 3);
    }
  } else {
    for (int i = threadIdx.x; i < T; i += kFpThreads) take(best, __ldg(p + i), i);
  }
  best = warp_argmax(best);
  __shared__ MaxIdx s[kFpThreads / 32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x < 32) {
    MaxIdx b = threadIdx.x < kFpThreads / 32 ? s[threadIdx.x] : MaxIdx{-CUDART_INF_F, 0x7fffffff};
    b = warp_argmax(b);
    if (threadIdx.x == 0) {
      int idx = b.i == 0x7fffffff ? 0 : b.i;  // only when every element is NaN
      float maxval = b.v;
      // get_max_preds: x = idx % W, y = floor(idx / W), zeroed when maxval <= 0
      float x = (float)(idx % W), y = (float)(idx / W);
      if (!(maxval > 0.0f)) {
        x = 0.f;
        y = 0.f;
      }
      // get_final_preds: +-0.25 toward the larger neighbour, interior maxima only
      int px = (int)floorf(x + 0.5f), py = (int)floorf(y + 0.5f);
      if (1 < px && px < W - 1 && 1 < py && py < H - 1) {
        float dx = p[py * W + px + 1] - p[py * W + px - 1];
        float dy = p[(py + 1) * W + px] - p[(py - 1) * W + px];
        x += (dx > 0.f ? 0.25f : (dx < 0.f ? -0.25f : 0.f));
        y += (dy > 0.f ? 0.25f : (dy < 0.f ? -0.25f : 0.f));
      }
      if (out_idx) out_idx[map] = idx;
      if (out_maxvals) out_maxvals[map] = maxval;
      if (out_coords) {
        out_coords[2 * map] = x;
        out_coords[2 * map + 1] = y;
      }
      if (out_preds) {
        // transform_preds with rot = 0: isotropic scale k = scale[0]*200/W about
        // the box centre (get_affine_transform(..., inv=1), utils/transform.py:76-105)
        int n = map / J;
        double k = (double)(__ldg(scale + 2 * n) * 200.0f) / (double)W;
        double cx = (double)__ldg(center + 2 * n), cy = (double)__ldg(center + 2 * n + 1);
        out_preds[2 * map] = (float)(cx + k * ((double)x - 0.5 * W));
        out_preds[2 * map + 1] = (float)(cy + k * ((double)y - 0.5 * H));
      }
    }
  }
}

}  // namespace otp

extern "C" int otp_final_preds(const float *heatmaps, int n, int j, int h, int w, const float *center,
                               const float *scale, int32_t *out_idx, float *out_coords,
                               float *out_preds, float *out_maxvals, otp_stream_t stream) {
  OTP_REQUIRE(n >= 0 && j > 0 && h > 0 && w > 0);
  OTP_REQUIRE((long long)h * w < 0x7fffffffLL);
  if (n == 0) return OTP_OK;
  OTP_REQUIRE(heatmaps != nullptr);
  OTP_REQUIRE(out_preds == nullptr || (center != nullptr && scale != nullptr));
  otp::LaunchScope ls(otp::K_FINAL_PREDS, (cudaStream_t)stream);
  otp::final_preds_kernel<<<n * j, otp::kFpThreads, 0, (cudaStream_t)stream>>>(
      heatmaps, j, h, w, center, scale, out_idx, out_coords, out_preds, out_maxvals);
  return otp::check_launch("final_preds_kernel");
}
