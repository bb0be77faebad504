// Per-clip softmax + fold between the front and apply passes of a TransformerBlock
// (reference model/blocks.py:432-440):
//   S   = fixed-order sum of the partial channel Grams          (deterministic)
//   A_h = softmax_rows(S_h)                                      (hs x hs per head)
//   W_eff[n][c] = sum_j A[n][j] * W_v[h(n)*hs + j][c],  b_eff[n] = sum_j A[n][j] * b_v[h(n)*hs + j]
// so that att @ (W_v vn + b_v) == W_eff vn + b_eff and v is never formed.
//
// One CTA per (clip, block of NB attention rows): it reduces only its own rows of the
// partial Grams, so no work is duplicated and B * C/NB CTAs fill the machine; in the
// fold loop consecutive threads own consecutive input channels c (coalesced W_v reads,
// A rows as shared-memory broadcasts, NB independent accumulators).
//
// Output format OUT: 0 = fp32 transposed [c][NPAD] (CUDA-core apply pass),
//                    1 = bf16 / 2 = fp16 UMMA operand image [n][c] (tensor-core apply pass).
#pragma once
#include "block_common.cuh"
#include "tc_common.cuh"

namespace otp {

constexpr int kFoldThreads = 160;   // >= padded channel count (144)

template <int C_>
struct FoldCfg {
  static constexpr int NB = (BlockCfg<C_>::HS % 4 == 0) ? 4 : 1;   // rows per CTA, never straddles a head
  static constexpr int NBLK = (C_ + NB - 1) / NB;
};

template <int C_, int OUT>
__global__ void __launch_bounds__(kFoldThreads)
block_fold_kernel(const float *__restrict__ wv, const float *__restrict__ bv,
                  const float *__restrict__ gram_part, int nchunk, void *__restrict__ weff_out,
                  float *__restrict__ beff, int cpad, int beff_stride) {
  using Cfg = BlockCfg<C_>;
  constexpr int C = Cfg::C, HS = Cfg::HS, NB = FoldCfg<C_>::NB, LDS = HS + 1;
  __shared__ float S[NB * LDS];
  const int b = blockIdx.x, nb = blockIdx.y;
  const int n0 = nb * NB, h = n0 / HS;
  const float *gp = gram_part + (size_t)b * nchunk * C * HS + (size_t)n0 * HS;
  for (int e = threadIdx.x; e < NB * HS; e += kFoldThreads) {
    float s = 0.f;
#pragma unroll 8
    for (int ch = 0; ch < nchunk; ++ch) s += __ldg(gp + (size_t)ch * C * HS + e);  // fixed order
    S[(e / HS) * LDS + e % HS] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < NB) {
    float *row = S + warp * LDS;
    float m = -3.402823466e38f;
    for (int j = lane; j < HS; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < HS; j += 32) {
      float e = expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < HS; j += 32) row[j] *= inv;
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c < cpad) {
    float acc[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) acc[i] = 0.f;
    if (c < C) {
      const float *wp = wv + (size_t)(h * HS) * C + c;
#pragma unroll
      for (int j = 0; j < HS; ++j) {   // fully unrolled: all HS loads of W_v in flight at once
        const float w = __ldg(wp + (size_t)j * C);
#pragma unroll
        for (int i = 0; i < NB; ++i) acc[i] = fmaf(S[i * LDS + j], w, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int n = n0 + i;
      if constexpr (OUT == 0) {
        if (c < C) static_cast<float *>(weff_out)[((size_t)b * C + c) * Cfg::NPAD + n] = acc[i];
      } else {
        uint8_t *img = static_cast<uint8_t *>(weff_out) + (size_t)b * (cpad / 8) * (cpad / 8) * 128;
        *reinterpret_cast<unsigned short *>(img + tc::cm_offset(n, c, (cpad / 8) * 128, 128)) =
            tc::to16<OUT == 2>(acc[i]);
      }
    }
  }
  // b_eff of this CTA's rows
  if (threadIdx.x < NB) {
    const int i = threadIdx.x;
    float acc = 0.f;
    for (int j = 0; j < HS; ++j) acc = fmaf(S[i * LDS + j], __ldg(bv + h * HS + j), acc);
    beff[(size_t)b * beff_stride + n0 + i] = acc;
  }
  // the last CTA of a clip also zeroes the padding rows n in [C, npad)
  if (nb == FoldCfg<C_>::NBLK - 1) {
    if constexpr (OUT == 0) {
      for (int idx = threadIdx.x; idx < C * (Cfg::NPAD - C); idx += kFoldThreads) {
        const int cc = idx / (Cfg::NPAD - C), n = C + idx % (Cfg::NPAD - C);
        static_cast<float *>(weff_out)[((size_t)b * C + cc) * Cfg::NPAD + n] = 0.f;
      }
    } else {
      uint8_t *img = static_cast<uint8_t *>(weff_out) + (size_t)b * (cpad / 8) * (cpad / 8) * 128;
      for (int idx = threadIdx.x; idx < cpad * (cpad - C); idx += kFoldThreads) {
        const int cc = idx / (cpad - C), n = C + idx % (cpad - C);
        *reinterpret_cast<unsigned short *>(img + tc::cm_offset(n, cc, (cpad / 8) * 128, 128)) = 0;
      }
    }
    for (int n = C + threadIdx.x; n < beff_stride; n += kFoldThreads) beff[(size_t)b * beff_stride + n] = 0.f;
  }
}

}  // namespace otp
