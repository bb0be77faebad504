// a2-a5 (16-bit tensor-core path): the TransformerBlock passes of block_simt.cu with
// every pointwise-conv GEMM and the channel-Gram on tcgen05 tensor cores
// (UTCHMMA, accumulators in TMEM, 16-bit operands staged in shared memory in the
// core-matrix interleaved layout of tc_common.cuh).  C = 136, 2 heads.
//
//   tc_front  x -> LN1 -> {dw_q, dw_k, dw_v} -> {LN_q, LN_k, LN_v}        (CUDA cores, fp32)
//             q = Wq qn + bq, k = Wk kn + bk : two M128 x N144 x K144 UMMAs    (tensor)
//             S_h += q_h k_h^T over the tile's 128 tokens: the [token][channel] 16-bit q/k
//             tiles re-read as MN-major operands, accumulated in TMEM over the CTA's
//             whole token chunk                                                 (tensor)
//             vn (16-bit, already in operand layout) is written for tc_apply.
//   fold      softmax + W_eff = softmax(S) W_v, emitted as an operand image.
//   tc_apply  o = W_eff vn + b_eff : one M128 x N144 x K144 UMMA per tile, stored
//             token-major 16-bit == the reference's scramble buffer.
//   tc_back   proj UMMA -> u = skip + s_a(.) -> LN2 (registers) -> 6 x { W1 chunk UMMA
//             (N96) -> bias + erf-GELU -> H tile -> W2 chunk UMMA accumulating in TMEM }
//             -> y = u + s_m(.)
//
// fp32 is kept for the residual stream, all LayerNorm statistics, the Gram
// accumulation, softmax and every epilogue; only UMMA operands are 16-bit
// (template F16: false = bfloat16, true = IEEE half -- same speed, 8x finer rounding;
// every operand here is O(1)..O(100), far inside the half range).
#include "block_common.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;

namespace {
constexpr int kC = 136, kHS = 68, kKP = 144;
constexpr int kTM = 128;        // tokens per tile == UMMA M
constexpr int kTcThreads = 256; // thread (q4, lane, half): token = 32*q4+lane, channel half
constexpr uint32_t kCS = 128;                 // byte stride between 8-element K chunks
constexpr uint32_t kRS144 = (kKP / 8) * 128;  // byte stride between 8-row groups, K = 144
constexpr uint32_t kTile144 = (kTM / 8) * kRS144;  // 36864  [128][144]
constexpr uint32_t kW144 = (kKP / 8) * kRS144;     // 41472  [144][144]
constexpr int kNH = 96, kNChunk = 6, kHidPad = kNH * kNChunk;  // hidden 544 -> 576
constexpr uint32_t kRS96 = (kNH / 8) * 128;        // 1536
constexpr uint32_t kW1c = (kNH / 8) * kRS144;      // 27648  [96][144]
constexpr uint32_t kW2c = (kKP / 8) * kRS96;       // 27648  [144][96]
constexpr uint32_t kWc = kW1c + kW2c;              // 55296
constexpr uint32_t kHTile = (kTM / 8) * kRS96;     // 24576  [128][96]
constexpr int kXLD = 140;   // fp32 staging row stride; 136 staged tokens [s*ob-4, s*ob+132)
constexpr int kXOff = 3;    // staged index of input token s*ob-1 (first tap of output token ob)
constexpr int kNI = 130;    // tokens the tile's taps touch

// ---- packed tensor-core weights of one block (bytes) ----
struct TcPack {
  size_t wq, wk, wp, wc /* [6] x {W1 chunk, W2 chunk} */, img_bytes /* one operand format */;
  size_t b1f /* fp32 [576], after both formats */, total;
};
constexpr TcPack tc_pack_layout() {
  TcPack p{};
  p.wq = 0;
  p.wk = p.wq + kW144;
  p.wp = p.wk + kW144;
  p.wc = p.wp + kW144;
  p.img_bytes = p.wc + (size_t)kNChunk * kWc;
  p.b1f = 2 * p.img_bytes;  // images: [bf16 set][fp16 set]
  p.total = p.b1f + kHidPad * 4;
  return p;
}

struct TcWorkspace {
  size_t gram_part, beff, weff, vn, obuf, total;
  int tiles, tiles_per_chunk, nchunk, tout;
};
TcWorkspace tc_workspace(int b, int t, int stride) {
  TcWorkspace w{};
  w.tout = stride == 1 ? t : (t - 1) / 2 + 1;
  w.tiles = ceil_div(w.tout, kTM);
  w.tiles_per_chunk = (int)(((long long)w.tiles * b + 295) / 296);
  if (w.tiles_per_chunk < 1) w.tiles_per_chunk = 1;
  w.nchunk = ceil_div(w.tiles, w.tiles_per_chunk);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += align_up(bytes, 256);
    return r;
  };
  w.gram_part = take((size_t)b * w.nchunk * kC * kHS * 4);
  w.beff = take((size_t)b * kKP * 4);
  w.weff = take((size_t)b * kW144);
  w.vn = take((size_t)b * w.tiles * kTile144);
  w.obuf = take((size_t)b * kC * w.tout * 2 + 64);
  w.total = o;
  return w;
}

// erf-GELU with the Abramowitz-Stegun 7.1.26 rational/exponential form
// (|erf error| <= 1.5e-7): 1 MUFU.RCP + 1 MUFU.EX2 + ~12 FMA-pipe ops.
__device__ __forceinline__ float gelu_as(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, az, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-az * az);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  const float erf_z = copysignf(erf_abs, z);
  return 0.5f * x * (1.0f + erf_z);
}

// all threads: copy `bytes` (multiple of 16) global -> shared with cp.async
__device__ __forceinline__ void cp_async_block(uint8_t *dst, const uint8_t *src, uint32_t bytes) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += kTcThreads * 16) cp_async16(dst + o, src + o);
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

struct TcIds {
  int q4, lane, half, tok, warp;
};
__device__ __forceinline__ TcIds tc_ids() {
  TcIds t;
  t.warp = threadIdx.x >> 5;
  t.lane = threadIdx.x & 31;
  t.q4 = t.warp & 3;
  t.half = t.warp >> 2;
  t.tok = t.q4 * 32 + t.lane;
  return t;
}
// TMEM address of (this thread's lane quarter, column col)
__device__ __forceinline__ uint32_t tcol(uint32_t base, const TcIds &id, int col) {
  return base + ((uint32_t)(id.q4 * 32) << 16) + (uint32_t)col;
}

// ------------------------------------------------------------------ tc_front
struct FrontVec {  // small fp32 parameters staged in shared memory
  float ln1w[kC], ln1b[kC];
  float4 dwq[kC], dwk[kC], dwv[kC];
  float qnw[kC], qnb[kC], knw[kC], knb[kC], vnw[kC], vnb[kC];
  float bq[kKP], bk[kKP];
  float part[6][2][kTM];
};

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_front_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const float *__restrict__ x,
                float *__restrict__ gram_part, uint8_t *__restrict__ vn_img, int T, int Tout, int stride,
                int tiles, int tiles_per_chunk, int nchunk, float qscale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float *xs = reinterpret_cast<float *>(smem);  // [136][140] fp32 staging, later the Wk image
  uint8_t *aq = smem + kC * kXLD * 4;
  uint8_t *ak = aq + kTile144;
  uint8_t *wq = ak + kTile144;
  FrontVec *V = reinterpret_cast<FrontVec *>(wq + kW144);
  __shared__ uint64_t bar_mma, bar_gram;
  __shared__ uint32_t tmem_slot;

  const TcIds id = tc_ids();
  const int b = blockIdx.y, chunk = blockIdx.x;
  constexpr TcPack L = tc_pack_layout();

  cp_async_block(wq, tcw + L.wq, kW144);
  cp_async_commit();
  for (int c = threadIdx.x; c < kC; c += kTcThreads) {
    V->ln1w[c] = P.ln1_w[c]; V->ln1b[c] = P.ln1_b[c];
    V->dwq[c] = make_float4(P.dwq[3 * c], P.dwq[3 * c + 1], P.dwq[3 * c + 2], 0.f);
    V->dwk[c] = make_float4(P.dwk[3 * c], P.dwk[3 * c + 1], P.dwk[3 * c + 2], 0.f);
    V->dwv[c] = make_float4(P.dwv[3 * c], P.dwv[3 * c + 1], P.dwv[3 * c + 2], 0.f);
    V->qnw[c] = P.qn_w[c]; V->qnb[c] = P.qn_b[c]; V->knw[c] = P.kn_w[c]; V->knb[c] = P.kn_b[c];
    V->vnw[c] = P.vn_w[c]; V->vnb[c] = P.vn_b[c];
  }
  for (int c = threadIdx.x; c < kKP; c += kTcThreads) {
    V->bq[c] = P.bq[c];
    V->bk[c] = P.bk[c];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_gram, 1);
    fence_mbar_init();
  }
  if (id.warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_q = tm, t_k = tm + 144, t_g0 = tm + 288, t_g1 = tm + 368;
  constexpr uint32_t kFmt = F16 ? 0u : 1u;
  const uint32_t idesc_qk = make_idesc_16(kKP, false, false, kFmt);
  const uint32_t idesc_gram = make_idesc_16(80, true, true, kFmt);
  uint32_t ph_mma = 0, ph_gram = 0;
  bool gram_pending = false;

  const float *xb = x + (size_t)b * kC * T;
  const int c_lo = id.half * 72, c_hi = id.half ? kC : 72;
  const int tile_begin = chunk * tiles_per_chunk;
  const int tile_end = min(tiles, tile_begin + tiles_per_chunk);
  const int ntok = kTM / stride;
  const bool vec_ok = (T & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const float *xsh = xs + kXOff;   // xsh[c*kXLD + i] = input token (s*ob - 1 + i)

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int t0 = tile * kTM;
    const int nvalid = min(kTM, Tout - t0);
    uint8_t *vn_tile = vn_img + ((size_t)b * tiles + tile) * kTile144;
    for (int round = 0; round < stride; ++round) {
      const int ob = t0 + round * ntok;    // first output token of this round
      const int ib = stride * ob - 1;      // input token of its first tap
      const int ab = stride * ob - 4;      // first staged token (16-byte aligned row offset)
      // ---- stage the fp32 input tile: asynchronous copies, everything in flight at once ----
      if (vec_ok) {
        for (int e = threadIdx.x; e < kC * 34; e += kTcThreads) {
          const int c = e / 34, q = e % 34, tk = ab + 4 * q;
          float *dst = xs + c * kXLD + 4 * q;
          if (tk >= 0 && tk + 3 < T) cp_async16(dst, xb + (size_t)c * T + tk);
          else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int e = threadIdx.x; e < kC * 136; e += kTcThreads) {
          const int c = e / 136, i = e % 136, tk = ab + i;
          float *dst = xs + c * kXLD + i;
          if (tk >= 0 && tk < T) cp_async4(dst, xb + (size_t)c * T + tk);
          else *dst = 0.f;
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      // ---- LN1 over channels per staged token (two-pass, fp32) ----
      const int extra = (id.tok < kNI - kTM) ? kTM + id.tok : -1;  // tokens 128, 129
      {
        float s0 = 0.f, s1 = 0.f;
        for (int c = c_lo; c < c_hi; ++c) {
          s0 += xsh[c * kXLD + id.tok];
          if (extra >= 0) s1 += xsh[c * kXLD + extra];
        }
        V->part[0][id.half][id.tok] = s0;
        if (extra >= 0) V->part[1][id.half][id.tok] = s1;
      }
      __syncthreads();
      const float mu0 = (V->part[0][0][id.tok] + V->part[0][1][id.tok]) * (1.0f / kC);
      const float mu1 = extra >= 0 ? (V->part[1][0][id.tok] + V->part[1][1][id.tok]) * (1.0f / kC) : 0.f;
      {
        float s0 = 0.f, s1 = 0.f;
        for (int c = c_lo; c < c_hi; ++c) {
          float d = xsh[c * kXLD + id.tok] - mu0;
          s0 = fmaf(d, d, s0);
          if (extra >= 0) {
            float e = xsh[c * kXLD + extra] - mu1;
            s1 = fmaf(e, e, s1);
          }
        }
        V->part[2][id.half][id.tok] = s0;
        if (extra >= 0) V->part[3][id.half][id.tok] = s1;
      }
      __syncthreads();
      {
        const float r0 = 1.0f / sqrtf((V->part[2][0][id.tok] + V->part[2][1][id.tok]) * (1.0f / kC) + 1e-5f);
        const bool v0 = (ib + id.tok >= 0) && (ib + id.tok < T);
        float r1 = 0.f;
        bool v1 = false;
        if (extra >= 0) {
          r1 = 1.0f / sqrtf((V->part[3][0][id.tok] + V->part[3][1][id.tok]) * (1.0f / kC) + 1e-5f);
          v1 = (ib + extra >= 0) && (ib + extra < T);
        }
        float *xw = xs + kXOff;
        for (int c = c_lo; c < c_hi; ++c) {
          const float w = V->ln1w[c], bb = V->ln1b[c];
          float *p0 = xw + c * kXLD + id.tok;
          *p0 = v0 ? fmaf((*p0 - mu0) * r0, w, bb) : 0.f;   // zero == conv zero padding
          if (extra >= 0) {
            float *p1 = xw + c * kXLD + extra;
            *p1 = v1 ? fmaf((*p1 - mu1) * r1, w, bb) : 0.f;
          }
        }
      }
      __syncthreads();
      // ---- depthwise convs + LN_q / LN_k / LN_v statistics (one pass: sum, sum of squares) ----
      const bool active = id.tok < ntok;
      const int xi = stride * id.tok;
      float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (active) {
        for (int c = c_lo; c < c_hi; ++c) {
          const float *xr = xsh + c * kXLD + xi;
          const float y0 = xr[0], y1 = xr[1], y2 = xr[2];
          const float4 a = V->dwq[c], bb = V->dwk[c], cc = V->dwv[c];
          const float dq = fmaf(a.z, y2, fmaf(a.y, y1, a.x * y0));
          const float dk = fmaf(bb.z, y2, fmaf(bb.y, y1, bb.x * y0));
          const float dv = fmaf(cc.z, y2, fmaf(cc.y, y1, cc.x * y0));
          st[0] += dq; st[1] = fmaf(dq, dq, st[1]);
          st[2] += dk; st[3] = fmaf(dk, dk, st[3]);
          st[4] += dv; st[5] = fmaf(dv, dv, st[5]);
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) V->part[k][id.half][id.tok] = st[k];
      __syncthreads();
      float mean[3], rs[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float s = V->part[2 * k][0][id.tok] + V->part[2 * k][1][id.tok];
        const float ss = V->part[2 * k + 1][0][id.tok] + V->part[2 * k + 1][1][id.tok];
        mean[k] = s * (1.0f / kC);
        const float var = fmaxf(ss * (1.0f / kC) - mean[k] * mean[k], 0.f);
        rs[k] = 1.0f / sqrtf(var + 1e-5f);
      }
      // the previous tile's Gram UMMAs still read aq/ak: wait before overwriting them
      if (gram_pending) {
        mbar_wait(&bar_gram, ph_gram);
        ph_gram ^= 1;
        gram_pending = false;
      }
      // ---- normalise, round to 16 bit, store operand tiles (16-byte, conflict-free) ----
      if (active) {
        const int row = round * ntok + id.tok;
        const int ch_lo = id.half ? 9 : 0, ch_hi = id.half ? 18 : 9;
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
          float q8[8], k8[8], v8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            if (c < kC) {
              const float *xr = xsh + c * kXLD + xi;
              const float y0 = xr[0], y1 = xr[1], y2 = xr[2];
              const float4 a = V->dwq[c], bb = V->dwk[c], cc = V->dwv[c];
              const float dq = fmaf(a.z, y2, fmaf(a.y, y1, a.x * y0));
              const float dk = fmaf(bb.z, y2, fmaf(bb.y, y1, bb.x * y0));
              const float dv = fmaf(cc.z, y2, fmaf(cc.y, y1, cc.x * y0));
              q8[e] = fmaf((dq - mean[0]) * rs[0], V->qnw[c], V->qnb[c]);
              k8[e] = fmaf((dk - mean[1]) * rs[1], V->knw[c], V->knb[c]);
              v8[e] = fmaf((dv - mean[2]) * rs[2], V->vnw[c], V->vnb[c]);
            } else {
              q8[e] = k8[e] = v8[e] = 0.f;
            }
          }
          const uint32_t off = cm_offset(row, ch * 8, kRS144, kCS);
          *reinterpret_cast<uint4 *>(aq + off) = pack16x8<F16>(q8);
          *reinterpret_cast<uint4 *>(ak + off) = pack16x8<F16>(k8);
          *reinterpret_cast<uint4 *>(vn_tile + off) = pack16x8<F16>(v8);
        }
      }
      __syncthreads();  // xs is re-staged by the next round / reused for Wk
    }
    // ---- Wk image -> the (now dead) staging region; q and k projections on the tensor cores ----
    cp_async_block(reinterpret_cast<uint8_t *>(xs), tcw + L.wk, kW144);
    cp_async_commit();
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_q = smem_u32(aq), a_k = smem_u32(ak), w_q = smem_u32(wq), w_k = smem_u32(xs);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s) {
        umma_bf16(t_q, make_desc(a_q + s * 2 * kCS, kCS, kRS144), make_desc(w_q + s * 2 * kCS, kCS, kRS144),
                  idesc_qk, s > 0);
      }
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s) {
        umma_bf16(t_k, make_desc(a_k + s * 2 * kCS, kCS, kRS144), make_desc(w_k + s * 2 * kCS, kCS, kRS144),
                  idesc_qk, s > 0);
      }
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    // ---- epilogue: + bias, * 1/sqrt(hs) for q, 16 bit, back into aq / ak as [token][channel] ----
    {
      const bool live = id.tok < nvalid;   // padded tokens must not reach the Gram
#pragma unroll 1
      for (int g = 0; g < 9; ++g) {
        const int col = id.half * 72 + g * 8;
        float vq[8], vk[8];
        tmem_ld8(tcol(t_q, id, col), vq);
        tmem_ld8(tcol(t_k, id, col), vk);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          vq[e] = live ? (vq[e] + V->bq[col + e]) * qscale : 0.f;
          vk[e] = live ? (vk[e] + V->bk[col + e]) : 0.f;
        }
        const uint32_t off = cm_offset(id.tok, col, kRS144, kCS);
        *reinterpret_cast<uint4 *>(aq + off) = pack16x8<F16>(vq);
        *reinterpret_cast<uint4 *>(ak + off) = pack16x8<F16>(vk);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- channel Gram over this tile's tokens: MN-major views of the q / k tiles ----
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_q = smem_u32(aq), a_k = smem_u32(ak);
      const bool first = tile == tile_begin;
#pragma unroll
      for (int s = 0; s < kTM / 16; ++s) {
        const uint32_t ko = s * 2 * kRS144;
        // head 0: rows = q channels 0..127, cols = k channels 0..79
        umma_bf16(t_g0, make_desc(a_q + ko, kRS144, kCS), make_desc(a_k + ko, kRS144, kCS), idesc_gram,
                  !(first && s == 0));
        // head 1: rows = q channels 8..135, cols = k channels 64..143
        umma_bf16(t_g1, make_desc(a_q + ko + kCS, kRS144, kCS), make_desc(a_k + ko + 8 * kCS, kRS144, kCS),
                  idesc_gram, !(first && s == 0));
      }
      umma_commit(&bar_gram);
    }
    gram_pending = true;
  }
  if (gram_pending) {
    mbar_wait(&bar_gram, ph_gram);
    ph_gram ^= 1;
  }
  tc_fence_after();
  // ---- flush the partial Gram: TMEM lane == q channel (row), column == k channel ----
  {
    float *gp = gram_part + (size_t)(b * nchunk + chunk) * kC * kHS;
    const int row_ch = id.half ? 8 + id.tok : id.tok;          // q channel of this lane
    const bool row_ok = id.half ? (row_ch >= kHS && row_ch < kC) : (row_ch < kHS);
    const int col0 = id.half ? 4 : 0;                            // first useful column
    const uint32_t tg = id.half ? t_g1 : t_g0;
#pragma unroll 1
    for (int g = 0; g < 10; ++g) {
      float v[8];
      tmem_ld8(tcol(tg, id, g * 8), v);
      if (row_ok) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = g * 8 + e - col0;
          if (j >= 0 && j < kHS) gp[(size_t)row_ch * kHS + j] = v[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (id.warp == 0) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------ fold (operand image)
template <bool F16>
__global__ void __launch_bounds__(256)
tc_fold_kernel(BlockPack P, const float *__restrict__ gram_part, int nchunk, uint8_t *__restrict__ weff_img,
               float *__restrict__ beff) {
  constexpr int LDS = kHS + 1;
  __shared__ float S[kC * LDS];
  const int b = blockIdx.x, cs = blockIdx.y;
  const float *gp = gram_part + (size_t)b * nchunk * kC * kHS;
  for (int e = threadIdx.x; e < kC * kHS; e += 256) {
    float s = 0.f;
    for (int ch = 0; ch < nchunk; ++ch) s += __ldg(gp + (size_t)ch * kC * kHS + e);  // fixed order
    S[(e / kHS) * LDS + e % kHS] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < kC; r += 8) {
    float *row = S + r * LDS;
    float m = -3.402823466e38f;
    for (int j = lane; j < kHS; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < kHS; j += 32) {
      float e = expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < kHS; j += 32) row[j] *= inv;
  }
  __syncthreads();
  // W_eff[n][c] = sum_j A[n][j] * Wv[h(n)*hs + j][c]  -> image row n, K index c
  uint8_t *img = weff_img + (size_t)b * kW144;
  const int c_lo = (kKP * cs) / 4, c_hi = (kKP * (cs + 1)) / 4;
  for (int idx = threadIdx.x; idx < (c_hi - c_lo) * kKP; idx += 256) {
    const int n = idx % kKP, c = c_lo + idx / kKP;
    float acc = 0.f;
    if (n < kC && c < kC) {
      const int h = n / kHS;
      const float *a = S + n * LDS;
      const float *wv = P.wv + (size_t)(h * kHS) * kC + c;
      for (int j = 0; j < kHS; ++j) acc = fmaf(a[j], __ldg(wv + (size_t)j * kC), acc);
    }
    *reinterpret_cast<unsigned short *>(img + cm_offset(n, c, kRS144, kCS)) = to16<F16>(acc);
  }
  if (cs == 0) {
    for (int n = threadIdx.x; n < kKP; n += 256) {
      float acc = 0.f;
      if (n < kC) {
        const int h = n / kHS;
        for (int j = 0; j < kHS; ++j) acc = fmaf(S[n * LDS + j], __ldg(P.bv + h * kHS + j), acc);
      }
      beff[(size_t)b * kKP + n] = acc;
    }
  }
}

// ------------------------------------------------------------------ tc_apply
template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_apply_kernel(const uint8_t *__restrict__ vn_img, const uint8_t *__restrict__ weff_img,
                const float *__restrict__ beff, unsigned short *__restrict__ obuf, int Tout, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *a = smem;
  uint8_t *w = smem + kTile144;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sb[kKP];
  const TcIds id = tc_ids();
  const int b = blockIdx.y, tile = blockIdx.x;
  const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
  cp_async_block(a, vn_img + ((size_t)b * tiles + tile) * kTile144, kTile144);
  cp_async_block(w, weff_img + (size_t)b * kW144, kW144);
  cp_async_commit();
  for (int n = threadIdx.x; n < kKP; n += kTcThreads) sb[n] = beff[(size_t)b * kKP + n];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (id.warp == 0) tmem_alloc(&tmem_slot, 256);
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_16(kKP, false, false, F16 ? 0u : 1u);
    const uint32_t aa = smem_u32(a), ww = smem_u32(w);
#pragma unroll
    for (int s = 0; s < kKP / 16; ++s)
      umma_bf16(tm, make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(ww + s * 2 * kCS, kCS, kRS144), idesc, s > 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // token-major store obuf[b][head][token][68] (16 bit) == the reference's scramble buffer
  unsigned short *ob = obuf + (size_t)b * kC * Tout;
#pragma unroll 1
  for (int g = 0; g < 9; ++g) {
    const int col = id.half * 72 + g * 8;
    float v[8];
    tmem_ld8(tcol(tm, id, col), v);
    if (id.tok < nvalid) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n4 = col + 4 * q;
        if (n4 < kC) {
          const int h = n4 >= kHS ? 1 : 0, cp = n4 - h * kHS;
          uint2 pk;
          pk.x = pack16x2<F16>(v[4 * q] + sb[n4], v[4 * q + 1] + sb[n4 + 1]);
          pk.y = pack16x2<F16>(v[4 * q + 2] + sb[n4 + 2], v[4 * q + 3] + sb[n4 + 3]);
          *reinterpret_cast<uint2 *>(ob + ((size_t)h * Tout + t0 + id.tok) * kHS + cp) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (id.warp == 0) tmem_dealloc(tm, 256);
}

// ------------------------------------------------------------------ tc_back
struct BackVec {
  float bp[kKP], sa[kKP], b2[kKP], sm[kKP], ln2w[kKP], ln2b[kKP];
  float b1f[kHidPad];
  float part[2][2][kTM];
};

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_back_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const float *__restrict__ b1f,
               const float *__restrict__ x, const unsigned short *__restrict__ obuf, float *__restrict__ y,
               int B, int T, int Tout, int stride, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *a = smem;                    // out2 tile, then LN2(u)
  uint8_t *hbuf = a + kTile144;         // GELU(hidden chunk)
  uint8_t *wp = hbuf + kHTile;
  uint8_t *wb0 = wp + kW144;
  uint8_t *wb1 = wb0 + kWc;
  BackVec *V = reinterpret_cast<BackVec *>(wb1 + kWc);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const TcIds id = tc_ids();
  constexpr TcPack L = tc_pack_layout();

  cp_async_block(wp, tcw + L.wp, kW144);
  cp_async_commit();
  for (int n = threadIdx.x; n < kKP; n += kTcThreads) {
    V->bp[n] = P.bp[n]; V->sa[n] = P.sa[n]; V->b2[n] = P.b2[n]; V->sm[n] = P.sm[n];
    V->ln2w[n] = n < kC ? P.ln2_w[n] : 0.f;
    V->ln2b[n] = n < kC ? P.ln2_b[n] : 0.f;
  }
  for (int n = threadIdx.x; n < kHidPad; n += kTcThreads) V->b1f[n] = b1f[n];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (id.warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_y = tm, t_a = tm + 144;   // y accumulator | proj accumulator, then hidden chunk
  constexpr uint32_t kFmt = F16 ? 0u : 1u;
  const uint32_t idesc144 = make_idesc_16(kKP, false, false, kFmt);
  const uint32_t idesc96 = make_idesc_16(kNH, false, false, kFmt);
  uint32_t ph = 0;
  const int col_lo = id.half * 72;

  for (int g = blockIdx.x; g < B * tiles; g += gridDim.x) {
    const int b = g / tiles, tile = g % tiles;
    const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
    const bool live = id.tok < nvalid;
    // weight chunks 0 and 1 in flight while the activations are staged
    cp_async_block(wb0, tcw + L.wc, kWc);
    cp_async_commit();
    cp_async_block(wb1, tcw + L.wc + kWc, kWc);
    cp_async_commit();
    // ---- the (nh, T', hs) 16-bit buffer re-read as (C, T'): A tile [token][channel].
    //      All 72 loads of a thread are issued before the first use (memory-level parallelism).
    {
      const unsigned short *ob = obuf + (size_t)b * kC * Tout + t0 + id.tok;
      unsigned short ov[72];
#pragma unroll
      for (int i = 0; i < 72; ++i) {
        const int c = col_lo + i;
        ov[i] = (live && c < kC) ? __ldg(ob + (size_t)c * Tout) : (unsigned short)0;
      }
#pragma unroll
      for (int gg = 0; gg < 9; ++gg) {
        uint4 w4;
        w4.x = (uint32_t)ov[gg * 8 + 0] | ((uint32_t)ov[gg * 8 + 1] << 16);
        w4.y = (uint32_t)ov[gg * 8 + 2] | ((uint32_t)ov[gg * 8 + 3] << 16);
        w4.z = (uint32_t)ov[gg * 8 + 4] | ((uint32_t)ov[gg * 8 + 5] << 16);
        w4.w = (uint32_t)ov[gg * 8 + 6] | ((uint32_t)ov[gg * 8 + 7] << 16);
        *reinterpret_cast<uint4 *>(a + cm_offset(id.tok, col_lo + gg * 8, kRS144, kCS)) = w4;
      }
    }
    // ---- skip path: pool_skip(x) prefetched into the registers that will hold u ----
    float u[72];
    {
      const float *xb = x + (size_t)b * kC * T;
      const int tt = t0 + id.tok;
#pragma unroll
      for (int i = 0; i < 72; ++i) {
        const int n = col_lo + i;
        float skip = 0.f;
        if (n < kC && live) {
          const float *xr = xb + (size_t)n * T;
          if (stride == 1) {
            skip = __ldg(xr + tt);
          } else {   // MaxPool1d(3, 2, 1)
            const int c0 = 2 * tt;
            skip = __ldg(xr + c0);
            if (c0 - 1 >= 0) skip = fmaxf(skip, __ldg(xr + c0 - 1));
            if (c0 + 1 < T) skip = fmaxf(skip, __ldg(xr + c0 + 1));
          }
        }
        u[i] = skip;
      }
    }
    cp_async_wait<2>();   // Wp (first tile); no-op afterwards
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(a), ww = smem_u32(wp);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_bf16(t_a, make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(ww + s * 2 * kCS, kCS, kRS144), idesc144,
                  s > 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    // ---- u = skip(x) + s_a * (proj + b_p); LN2 statistics from registers ----
    {
      float s = 0.f;
#pragma unroll
      for (int gg = 0; gg < 9; ++gg) {
        float v[8];
        tmem_ld8(tcol(t_a, id, col_lo + gg * 8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = col_lo + gg * 8 + e;
          const float val = (n < kC && live) ? fmaf(V->sa[n], v[e] + V->bp[n], u[gg * 8 + e]) : 0.f;
          u[gg * 8 + e] = val;
          s += val;
        }
      }
      V->part[0][id.half][id.tok] = s;
    }
    __syncthreads();
    const float mu = (V->part[0][0][id.tok] + V->part[0][1][id.tok]) * (1.0f / kC);
    {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 72; ++i) {
        const float d = (col_lo + i < kC) ? u[i] - mu : 0.f;
        s = fmaf(d, d, s);
      }
      V->part[1][id.half][id.tok] = s;
    }
    __syncthreads();
    {
      const float rstd = 1.0f / sqrtf((V->part[1][0][id.tok] + V->part[1][1][id.tok]) * (1.0f / kC) + 1e-5f);
#pragma unroll
      for (int gg = 0; gg < 9; ++gg) {
        float h8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = col_lo + gg * 8 + e;
          h8[e] = fmaf((u[gg * 8 + e] - mu) * rstd, V->ln2w[n], V->ln2b[n]);   // pad channels: w = b = 0
        }
        *reinterpret_cast<uint4 *>(a + cm_offset(id.tok, col_lo + gg * 8, kRS144, kCS)) = pack16x8<F16>(h8);
      }
    }
    // ---- MLP: hidden chunk j on the tensor cores, GELU on the CUDA cores, W2 accumulates in TMEM ----
    cp_async_wait<1>();   // W chunk 0
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(a), w1 = smem_u32(wb0);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_bf16(t_a, make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(w1 + s * 2 * kCS, kCS, kRS144), idesc96,
                  s > 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int j = 0; j < kNChunk; ++j) {
      // GELU(hidden chunk j) -> 16-bit H tile
#pragma unroll 1
      for (int gg = 0; gg < 6; ++gg) {
        const int col = id.half * 48 + gg * 8;
        float v[8];
        tmem_ld8(tcol(t_a, id, col), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = gelu_as(v[e] + V->b1f[j * kNH + col + e]);
        *reinterpret_cast<uint4 *>(hbuf + cm_offset(id.tok, col, kRS96, kCS)) = pack16x8<F16>(v);
      }
      cp_async_wait<0>();   // W chunk j+1 (issued one chunk ago)
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint8_t *wcur = (j & 1) ? wb1 : wb0, *wnext = (j & 1) ? wb0 : wb1;
        const uint32_t hh = smem_u32(hbuf), w2 = smem_u32(wcur) + kW1c;
#pragma unroll
        for (int s = 0; s < kNH / 16; ++s)
          umma_bf16(t_y, make_desc(hh + s * 2 * kCS, kCS, kRS96), make_desc(w2 + s * 2 * kCS, kCS, kRS96), idesc144,
                    (j > 0 || s > 0));
        if (j + 1 < kNChunk) {
          const uint32_t aa = smem_u32(a), w1 = smem_u32(wnext);
#pragma unroll
          for (int s = 0; s < kKP / 16; ++s)
            umma_bf16(t_a, make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(w1 + s * 2 * kCS, kCS, kRS144),
                      idesc96, s > 0);
        }
        umma_commit(&bar);
      }
      mbar_wait(&bar, ph);
      ph ^= 1;
      tc_fence_after();
      // chunk j's weight buffer is free: prefetch chunk j+2 into it
      if (j + 2 < kNChunk) cp_async_block((j & 1) ? wb1 : wb0, tcw + L.wc + (size_t)(j + 2) * kWc, kWc);
      cp_async_commit();
    }
    // ---- y = u + s_m * (mlp + b_2) ----
    {
      float *yb = y + (size_t)b * kC * Tout + t0 + id.tok;
#pragma unroll
      for (int gg = 0; gg < 9; ++gg) {
        float v[8];
        tmem_ld8(tcol(t_y, id, col_lo + gg * 8), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int n = col_lo + gg * 8 + e;
          if (n < kC && live) yb[(size_t)n * Tout] = fmaf(V->sm[n], v[e] + V->b2[n], u[gg * 8 + e]);
        }
      }
    }
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();   // TMEM / smem reuse by the next tile
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (id.warp == 0) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------ weight packing
template <bool F16>
__global__ void pack_image_kernel(const float *__restrict__ src, int ld, int row0, int col0, int rows_valid,
                                  int cols_valid, uint8_t *__restrict__ dst, int rows_pad, int cols_pad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows_pad * cols_pad) return;
  const int r = e / cols_pad, k = e % cols_pad;
  const float v = (r < rows_valid && k < cols_valid) ? src[(size_t)(row0 + r) * ld + col0 + k] : 0.f;
  *reinterpret_cast<unsigned short *>(dst + cm_offset(r, k, (cols_pad / 8) * 128, 128)) = to16<F16>(v);
}
__global__ void pack_b1_kernel(const float *__restrict__ src, float *__restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < kHidPad) dst[e] = e < 4 * kC ? src[e] : 0.f;
}

void pack_image(bool f16, const float *src, int ld, int row0, int col0, int rv, int cv, uint8_t *dst, int rp,
                int cp, cudaStream_t st) {
  if (f16)
    pack_image_kernel<true><<<ceil_div(rp * cp, 256), 256, 0, st>>>(src, ld, row0, col0, rv, cv, dst, rp, cp);
  else
    pack_image_kernel<false><<<ceil_div(rp * cp, 256), 256, 0, st>>>(src, ld, row0, col0, rv, cv, dst, rp, cp);
}

constexpr size_t kFrontSmem = (size_t)kC * kXLD * 4 + 2 * kTile144 + kW144 + sizeof(FrontVec);
constexpr size_t kApplySmem = (size_t)kTile144 + kW144;
constexpr size_t kBackSmem = (size_t)kTile144 + kHTile + kW144 + 2 * kWc + sizeof(BackVec);
static_assert(kFrontSmem <= 226 * 1024, "tc_front shared memory");
static_assert(kBackSmem <= 226 * 1024, "tc_back shared memory");

template <bool F16>
int forward_tc_t(const void *packed_fp32, const void *packed_tc, const float *x, float *y, int b, int t, int stride,
                 void *ws_tc, cudaStream_t st) {
  const TcWorkspace W = tc_workspace(b, t, stride);
  const BlockPack P = block_pack_view(packed_fp32, kC);
  constexpr TcPack L = tc_pack_layout();
  const uint8_t *tcw = static_cast<const uint8_t *>(packed_tc) + (F16 ? L.img_bytes : 0);
  const float *b1f = reinterpret_cast<const float *>(static_cast<const uint8_t *>(packed_tc) + L.b1f);
  uint8_t *ws = static_cast<uint8_t *>(ws_tc);
  float *gram = reinterpret_cast<float *>(ws + W.gram_part);
  float *beff = reinterpret_cast<float *>(ws + W.beff);
  uint8_t *weff = ws + W.weff;
  uint8_t *vn = ws + W.vn;
  unsigned short *obuf = reinterpret_cast<unsigned short *>(ws + W.obuf);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tc_front_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrontSmem);
    cudaFuncSetAttribute(tc_apply_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplySmem);
    cudaFuncSetAttribute(tc_back_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBackSmem);
    attr_done = true;
  }
  {
    LaunchScope ls(K_TC_FRONT, st);
    tc_front_kernel<F16><<<dim3(W.nchunk, b), kTcThreads, kFrontSmem, st>>>(
        P, tcw, x, gram, vn, t, W.tout, stride, W.tiles, W.tiles_per_chunk, W.nchunk, 1.0f / sqrtf((float)kHS));
  }
  {
    LaunchScope ls(K_BLOCK_FOLD, st);
    tc_fold_kernel<F16><<<dim3(b, 4), 256, 0, st>>>(P, gram, W.nchunk, weff, beff);
  }
  {
    LaunchScope ls(K_TC_APPLY, st);
    tc_apply_kernel<F16><<<dim3(W.tiles, b), kTcThreads, kApplySmem, st>>>(vn, weff, beff, obuf, W.tout, W.tiles);
  }
  {
    LaunchScope ls(K_TC_BACK, st);
    const int total = b * W.tiles;
    tc_back_kernel<F16><<<min(total, num_sms()), kTcThreads, kBackSmem, st>>>(P, tcw, b1f, x, obuf, y, b, t, W.tout,
                                                                             stride, W.tiles);
  }
  return check_launch("block_forward_tc");
}
}  // namespace

bool block_tc_built() { return true; }

size_t block_tc_packed_bytes(int c) { return c == kC ? align_up(tc_pack_layout().total, 1024) : 0; }

int block_tc_pack(const otp_block_params *p, int c, void *packed_tc, cudaStream_t st) {
  if (c != kC) return OTP_OK;
  constexpr TcPack L = tc_pack_layout();
  LaunchScope ls(K_PACK, st, 31);
  for (int f = 0; f < 2; ++f) {
    uint8_t *d = static_cast<uint8_t *>(packed_tc) + f * L.img_bytes;
    pack_image(f, p->q_w, kC, 0, 0, kC, kC, d + L.wq, kKP, kKP, st);
    pack_image(f, p->k_w, kC, 0, 0, kC, kC, d + L.wk, kKP, kKP, st);
    pack_image(f, p->proj_w, kC, 0, 0, kC, kC, d + L.wp, kKP, kKP, st);
    for (int j = 0; j < kNChunk; ++j) {
      const int hv = min(kNH, 4 * kC - j * kNH);
      pack_image(f, p->mlp0_w, kC, j * kNH, 0, hv, kC, d + L.wc + (size_t)j * kWc, kNH, kKP, st);
      pack_image(f, p->mlp3_w, 4 * kC, 0, j * kNH, kC, hv, d + L.wc + (size_t)j * kWc + kW1c, kKP, kNH, st);
    }
  }
  pack_b1_kernel<<<ceil_div(kHidPad, 256), 256, 0, st>>>(
      p->mlp0_b, reinterpret_cast<float *>(static_cast<uint8_t *>(packed_tc) + L.b1f));
  return check_launch("block_tc_pack");
}

size_t block_tc_workspace_bytes(int b, int c, int t, int stride) {
  return c == kC ? tc_workspace(b, t, stride).total : 0;
}

int block_forward_tc(const void *packed_fp32, const void *packed_tc, const float *x, float *y, int b, int c,
                     int t, int stride, int f16, void *ws_tc, cudaStream_t st) {
  (void)c;
  return f16 ? forward_tc_t<true>(packed_fp32, packed_tc, x, y, b, t, stride, ws_tc, st)
             : forward_tc_t<false>(packed_fp32, packed_tc, x, y, b, t, stride, ws_tc, st);
}

}  // namespace otp
