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
//   fold      softmax + W_eff = softmax(S) W_v, emitted as an operand image (block_fold.cuh).
//   tc_apply  o = W_eff vn + b_eff : one M128 x N144 x K144 UMMA per tile, stored
//             token-major 16-bit == the reference's scramble buffer.
//   tc_back   proj UMMA -> u = skip + s_a(.) -> LN2 (registers) -> 6 x { W1 chunk UMMA
//             (N96, double-buffered in TMEM so chunk j+1 runs under chunk j's GELU) ->
//             bias + GELU -> H tile -> W2 chunk UMMA accumulating in TMEM } -> y = u + s_m(.)
//
// fp32 is kept for the residual stream, all LayerNorm statistics, the Gram accumulation,
// softmax and every epilogue; only UMMA operands are 16-bit (template F16: false =
// bfloat16, true = IEEE half -- same speed, 8x finer rounding; every operand here is
// O(1)..O(100), far inside the half range).  The affine part of LN_q/LN_k/LN_v/LN2 is
// folded into the following GEMM's weights and bias when the weights are packed, so the
// staged operands are plain (x - mean) * rstd.
//
// Thread map of the 384-thread CTAs: warp w -> TMEM lane quarter q4 = w % 4 (a hardware
// rule of tcgen05.ld), token = 32*q4 + lane, and channel third = w / 4 (48 channels =
// six 16-byte operand chunks), so a thread's registers, its LayerNorm partial sums and
// its TMEM columns all refer to the same (token, channel-third).
#include "block_common.cuh"
#include "block_fold.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;

namespace {
constexpr int kC = 136, kHS = 68, kKP = 144;
constexpr int kTM = 128;        // tokens per tile == UMMA M
constexpr int kTcThreads = 384;
constexpr int kApplyThreads = 256;
constexpr uint32_t kCS = 128;                 // byte stride between 8-element K chunks
constexpr uint32_t kRS144 = (kKP / 8) * 128;  // byte stride between 8-row groups, K = 144
constexpr uint32_t kTile144 = (kTM / 8) * kRS144;  // 36864  [128][144]
constexpr uint32_t kW144 = (kKP / 8) * kRS144;     // 41472  [144][144]
constexpr int kNH = 64, kNChunk = 9, kHidPad = kNH * kNChunk;  // hidden 544 -> 576, 9 chunks of 64
constexpr uint32_t kRS96 = (kNH / 8) * 128;        // 768  (row-group stride of a K = 48 tile)
constexpr uint32_t kW1c = (kNH / 8) * kRS144;      // 13824  [48][144]
constexpr uint32_t kW2c = (kKP / 8) * kRS96;       // 13824  [144][48]
constexpr uint32_t kHTile = (kTM / 8) * kRS96;     // 12288  [128][48]
constexpr int kXLD = 140;   // fp32 staging row stride; 136 staged tokens [s*ob-4, s*ob+132)
constexpr int kXOff = 3;    // staged index of input token s*ob-1 (first tap of output token ob)
constexpr int kNI = 130;    // tokens the tile's taps touch

// ---- packed tensor-core weights of one block (bytes) ----
struct TcPack {
  size_t wq, wk, wp, w1 /* [6] */, w2 /* [6] */, img_bytes /* one operand format */;
  size_t wvp /* fp32 [136][136] Wv * g_v */, bvp /* [144] */, bqp, bkp /* [144] */, b1p /* [576] */, total;
};
constexpr TcPack tc_pack_layout() {
  TcPack p{};
  p.wq = 0;
  p.wk = p.wq + kW144;
  p.wp = p.wk + kW144;
  p.w1 = p.wp + kW144;
  p.w2 = p.w1 + (size_t)kNChunk * kW1c;
  p.img_bytes = p.w2 + (size_t)kNChunk * kW2c;
  p.wvp = 2 * p.img_bytes;  // images: [bf16 set][fp16 set], then the fp32 folded vectors
  p.bvp = p.wvp + (size_t)kC * kC * 4;
  p.bqp = p.bvp + kKP * 4;
  p.bkp = p.bqp + kKP * 4;
  p.b1p = p.bkp + kKP * 4;
  p.total = p.b1p + kHidPad * 4;
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

// GELU of two pre-activations, returned as a packed 16-bit operand pair.
// tanh form 0.5 x (1 + tanh(x (a + b x^2))) with (a, b) refitted to the erf GELU the
// reference uses (nn.GELU default): max |error| 2.7e-4, below the half spacing of the
// O(1) outputs it is rounded to.  F16: evaluated in packed half2 (7 instructions + one
// MUFU per PAIR); bf16: fp32 arithmetic, rounded at the end.
template <bool F16>
__device__ __forceinline__ uint32_t gelu_pair(float a, float b) {
  constexpr float kA = 0.80015708f, kB = 0.03470089f;
  if constexpr (F16) {
    const __half2 x = __floats2half2_rn(a, b);
    const __half2 x2 = __hmul2(x, x);
    const __half2 p = __hfma2(x2, __float2half2_rn(kB), __float2half2_rn(kA));
    const __half2 t = h2tanh_approx(__hmul2(x, p));
    const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
    const __half2 o = __hfma2(hx, t, hx);
    return *reinterpret_cast<const uint32_t *>(&o);
  } else {
    float r[2] = {a, b};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float x = r[i];
      float t;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * fmaf(kB, x * x, kA)));
      r[i] = fmaf(0.5f * x, t, 0.5f * x);
    }
    return pack16x2<false>(r[0], r[1]);
  }
}

__device__ __forceinline__ void cp_async_block(uint8_t *dst, const uint8_t *src, uint32_t bytes, int nthreads) {
  for (uint32_t o = threadIdx.x * 16; o < bytes; o += nthreads * 16) cp_async16(dst + o, src + o);
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

// tcgen05.ld without the wait, plus a wait that the loaded registers depend on.
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void reg_fence16(uint32_t (&r)[16]) {   // orders uses of r[] after the wait above
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
               "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
               "+r"(r[15])::"memory");
}
// 48 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld48(uint32_t taddr, float (&v)[48]) {
  uint32_t r0[16], r1[16], r2[16];
  tmem_ld16_nw(taddr, r0);
  tmem_ld16_nw(taddr + 16, r1);
  tmem_ld16_nw(taddr + 32, r2);
  tmem_wait_ld();
  reg_fence16(r0);
  reg_fence16(r1);
  reg_fence16(r2);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __uint_as_float(r0[i]);
    v[16 + i] = __uint_as_float(r1[i]);
    v[32 + i] = __uint_as_float(r2[i]);
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r0[16], r1[16];
  tmem_ld16_nw(taddr, r0);
  tmem_ld16_nw(taddr + 16, r1);
  tmem_wait_ld();
  reg_fence16(r0);
  reg_fence16(r1);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __uint_as_float(r0[i]);
    v[16 + i] = __uint_as_float(r1[i]);
  }
}

struct TcIds {
  int q4, lane, third, tok, warp;
};
__device__ __forceinline__ TcIds tc_ids() {
  TcIds t;
  t.warp = threadIdx.x >> 5;
  t.lane = threadIdx.x & 31;
  t.q4 = t.warp & 3;
  t.third = t.warp >> 2;
  t.tok = t.q4 * 32 + t.lane;
  return t;
}
// TMEM address of (this thread's lane quarter, column col)
__device__ __forceinline__ uint32_t tcol(uint32_t base, int q4, int col) {
  return base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)col;
}

// ------------------------------------------------------------------ tc_front
struct FrontVec {  // small fp32 parameters staged in shared memory
  float ln1w[kC], ln1b[kC];
  float4 dw[3][kC];          // depthwise taps of q, k, v
  float bq[kKP], bk[kKP];    // folded biases
  float part[2][3][kTM];     // per-third partial (sum, sum of squares) per token
  float parte[2][3][2];      // same for the two halo tokens 128, 129
};

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_front_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const float *__restrict__ bqp,
                const float *__restrict__ bkp, const float *__restrict__ x, float *__restrict__ gram_part,
                uint8_t *__restrict__ vn_img, int T, int Tout, int stride, int tiles, int tiles_per_chunk,
                int nchunk, float qscale) {
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

  cp_async_block(wq, tcw + L.wq, kW144, kTcThreads);
  cp_async_commit();
  for (int c = threadIdx.x; c < kC; c += kTcThreads) {
    V->ln1w[c] = P.ln1_w[c];
    V->ln1b[c] = P.ln1_b[c];
    V->dw[0][c] = make_float4(P.dwq[3 * c], P.dwq[3 * c + 1], P.dwq[3 * c + 2], 0.f);
    V->dw[1][c] = make_float4(P.dwk[3 * c], P.dwk[3 * c + 1], P.dwk[3 * c + 2], 0.f);
    V->dw[2][c] = make_float4(P.dwv[3 * c], P.dwv[3 * c + 1], P.dwv[3 * c + 2], 0.f);
  }
  for (int c = threadIdx.x; c < kKP; c += kTcThreads) {
    V->bq[c] = bqp[c];
    V->bk[c] = bkp[c];
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
  const int c_lo = id.third * 48, c_hi = min(kC, c_lo + 48);
  const int tile_begin = chunk * tiles_per_chunk;
  const int tile_end = min(tiles, tile_begin + tiles_per_chunk);
  const int ntok = kTM / stride;
  const bool vec_ok = (T & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  float *xsh = xs + kXOff;   // xsh[c*kXLD + i] = input token (s*ob - 1 + i)

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
      const bool has_e = id.tok < kNI - kTM;   // this thread also covers halo token 128 + tok
      const int extra = kTM + id.tok;
      {
        float s0 = 0.f, s1 = 0.f;
        for (int c = c_lo; c < c_hi; ++c) {
          s0 += xsh[c * kXLD + id.tok];
          if (has_e) s1 += xsh[c * kXLD + extra];
        }
        V->part[0][id.third][id.tok] = s0;
        if (has_e) V->parte[0][id.third][id.tok] = s1;
      }
      __syncthreads();
      const float mu0 = (V->part[0][0][id.tok] + V->part[0][1][id.tok] + V->part[0][2][id.tok]) * (1.0f / kC);
      const float mu1 = has_e ? (V->parte[0][0][id.tok] + V->parte[0][1][id.tok] + V->parte[0][2][id.tok]) * (1.0f / kC)
                              : 0.f;
      {
        float s0 = 0.f, s1 = 0.f;
        for (int c = c_lo; c < c_hi; ++c) {
          const float d = xsh[c * kXLD + id.tok] - mu0;
          s0 = fmaf(d, d, s0);
          if (has_e) {
            const float e = xsh[c * kXLD + extra] - mu1;
            s1 = fmaf(e, e, s1);
          }
        }
        V->part[1][id.third][id.tok] = s0;
        if (has_e) V->parte[1][id.third][id.tok] = s1;
      }
      __syncthreads();
      {
        const float r0 = 1.0f / sqrtf((V->part[1][0][id.tok] + V->part[1][1][id.tok] + V->part[1][2][id.tok]) *
                                          (1.0f / kC) + 1e-5f);
        const bool v0 = (ib + id.tok >= 0) && (ib + id.tok < T);
        float r1 = 0.f;
        bool v1 = false;
        if (has_e) {
          r1 = 1.0f / sqrtf((V->parte[1][0][id.tok] + V->parte[1][1][id.tok] + V->parte[1][2][id.tok]) * (1.0f / kC) +
                            1e-5f);
          v1 = (ib + extra >= 0) && (ib + extra < T);
        }
        for (int c = c_lo; c < c_hi; ++c) {
          const float w = V->ln1w[c], bb = V->ln1b[c];
          float *p0 = xsh + c * kXLD + id.tok;
          *p0 = v0 ? fmaf((*p0 - mu0) * r0, w, bb) : 0.f;   // zero == conv zero padding
          if (has_e) {
            float *p1 = xsh + c * kXLD + extra;
            *p1 = v1 ? fmaf((*p1 - mu1) * r1, w, bb) : 0.f;
          }
        }
      }
      __syncthreads();
      // ---- q, k, v in turn: depthwise conv (kept in registers) -> statistics -> (d-mean)*rstd ----
      const bool active = id.tok < ntok;
      const int xi = stride * id.tok;
      const int row = round * ntok + id.tok;
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        float d[48];
        float s = 0.f, ss = 0.f;
        if (active) {
#pragma unroll
          for (int i = 0; i < 48; ++i) {
            const int c = c_lo + i;
            float v = 0.f;
            if (c < kC) {
              const float *xr = xsh + c * kXLD + xi;
              const float4 w = V->dw[m][c];
              v = fmaf(w.z, xr[2], fmaf(w.y, xr[1], w.x * xr[0]));
            }
            d[i] = v;
            s += v;
            ss = fmaf(v, v, ss);
          }
        }
        V->part[0][id.third][id.tok] = s;
        V->part[1][id.third][id.tok] = ss;
        __syncthreads();
        const float mean = (V->part[0][0][id.tok] + V->part[0][1][id.tok] + V->part[0][2][id.tok]) * (1.0f / kC);
        const float var = fmaxf((V->part[1][0][id.tok] + V->part[1][1][id.tok] + V->part[1][2][id.tok]) * (1.0f / kC) -
                                    mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        // the previous tile's Gram UMMAs still read aq/ak: wait before overwriting them
        if (m == 0 && gram_pending) {
          mbar_wait(&bar_gram, ph_gram);
          ph_gram ^= 1;
          gram_pending = false;
        }
        if (active) {
          uint8_t *dst = (m == 0 ? aq : (m == 1 ? ak : vn_tile)) + cm_offset(row, c_lo, kRS144, kCS);
#pragma unroll
          for (int g = 0; g < 6; ++g) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = (c_lo + g * 8 + e < kC) ? (d[g * 8 + e] - mean) * rstd : 0.f;
            *reinterpret_cast<uint4 *>(dst + g * kCS) = pack16x8<F16>(o8);
          }
        }
        __syncthreads();   // part[] is rewritten by the next matrix; xs by the next round / Wk
      }
    }
    // ---- Wk image -> the (now dead) staging region; q and k projections on the tensor cores ----
    cp_async_block(reinterpret_cast<uint8_t *>(xs), tcw + L.wk, kW144, kTcThreads);
    cp_async_commit();
    cp_async_wait<0>();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_q = smem_u32(aq), a_k = smem_u32(ak), w_q = smem_u32(wq), w_k = smem_u32(xs);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_bf16(t_q, make_desc(a_q + s * 2 * kCS, kCS, kRS144), make_desc(w_q + s * 2 * kCS, kCS, kRS144),
                  idesc_qk, s > 0);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_bf16(t_k, make_desc(a_k + s * 2 * kCS, kCS, kRS144), make_desc(w_k + s * 2 * kCS, kCS, kRS144),
                  idesc_qk, s > 0);
      umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    // ---- epilogue: + bias, * 1/sqrt(hs) for q, 16 bit, back into aq / ak as [token][channel] ----
    {
      const bool live = id.tok < nvalid;   // padded tokens must not reach the Gram
      const int col = id.third * 48;
      const uint32_t off = cm_offset(id.tok, col, kRS144, kCS);
      float v[48];
      tmem_ld48(tcol(t_q, id.q4, col), v);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = live ? (v[g * 8 + e] + V->bq[col + g * 8 + e]) * qscale : 0.f;
        *reinterpret_cast<uint4 *>(aq + off + g * kCS) = pack16x8<F16>(o8);
      }
      tmem_ld48(tcol(t_k, id.q4, col), v);
#pragma unroll
      for (int g = 0; g < 6; ++g) {
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = live ? v[g * 8 + e] + V->bk[col + g * 8 + e] : 0.f;
        *reinterpret_cast<uint4 *>(ak + off + g * kCS) = pack16x8<F16>(o8);
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
  if (id.third < 2) {
    float *gp = gram_part + (size_t)(b * nchunk + chunk) * kC * kHS;
    const int row_ch = id.third ? 8 + id.tok : id.tok;          // q channel of this lane
    const bool row_ok = id.third ? (row_ch >= kHS && row_ch < kC) : (row_ch < kHS);
    const int col0 = id.third ? 4 : 0;                            // first useful column
    const uint32_t tg = id.third ? t_g1 : t_g0;
#pragma unroll 1
    for (int g = 0; g < 10; ++g) {
      float v[8];
      tmem_ld8(tcol(tg, id.q4, g * 8), v);
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

// ------------------------------------------------------------------ tc_apply
template <bool F16>
__global__ void __launch_bounds__(kApplyThreads, 2)
tc_apply_kernel(const uint8_t *__restrict__ vn_img, const uint8_t *__restrict__ weff_img,
                const float *__restrict__ beff, unsigned short *__restrict__ obuf, int Tout, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *a = smem;
  uint8_t *w = smem + kTile144;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sb[kKP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q4 = warp & 3, half = warp >> 2;
  const int tok = q4 * 32 + lane;
  const int b = blockIdx.y, tile = blockIdx.x;
  const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
  cp_async_block(a, vn_img + ((size_t)b * tiles + tile) * kTile144, kTile144, kApplyThreads);
  cp_async_block(w, weff_img + (size_t)b * kW144, kW144, kApplyThreads);
  cp_async_commit();
  for (int n = threadIdx.x; n < kKP; n += kApplyThreads) sb[n] = beff[(size_t)b * kKP + n];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
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
    const int col = half * 72 + g * 8;
    float v[8];
    tmem_ld8(tcol(tm, q4, col), v);
    if (tok < nvalid) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int n4 = col + 4 * q;
        if (n4 < kC) {
          const int h = n4 >= kHS ? 1 : 0, cp = n4 - h * kHS;
          uint2 pk;
          pk.x = pack16x2<F16>(v[4 * q] + sb[n4], v[4 * q + 1] + sb[n4 + 1]);
          pk.y = pack16x2<F16>(v[4 * q + 2] + sb[n4 + 2], v[4 * q + 3] + sb[n4 + 3]);
          *reinterpret_cast<uint2 *>(ob + ((size_t)h * Tout + t0 + tok) * kHS + cp) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ------------------------------------------------------------------ tc_back
int g_trace_on = 0;   // otp_debug_trace
#include "block_tc_back.cuh"

// ------------------------------------------------------------------ tc_front, stride-1 blocks
#include "block_tc_front.cuh"

// ------------------------------------------------------------------ weight packing
// operand image of src[row0 + r][col0 + k] * colscale[col0 + k]   (zero padded)
template <bool F16>
__global__ void pack_image_kernel(const float *__restrict__ src, int ld, int row0, int col0, int rows_valid,
                                  int cols_valid, const float *__restrict__ colscale, uint8_t *__restrict__ dst,
                                  int rows_pad, int cols_pad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows_pad * cols_pad) return;
  const int r = e / cols_pad, k = e % cols_pad;
  float v = 0.f;
  if (r < rows_valid && k < cols_valid) {
    v = src[(size_t)(row0 + r) * ld + col0 + k];
    if (colscale) v *= colscale[col0 + k];
  }
  *reinterpret_cast<unsigned short *>(dst + cm_offset(r, k, (cols_pad / 8) * 128, 128)) = to16<F16>(v);
}
// column `col` of an operand image <- bias[row0 + r] (r < rows_valid): the bias rides in the MMA
// against a ones column of the activation tile
template <bool F16>
__global__ void pack_bias_col_kernel(const float *__restrict__ bias, int row0, int rows_valid, uint8_t *__restrict__ dst,
                                     int cols_pad, int col) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows_valid)
    *reinterpret_cast<unsigned short *>(dst + cm_offset(r, col, (cols_pad / 8) * 128, 128)) = to16<F16>(bias[row0 + r]);
}
// out[n] = bias[n] + sum_c w[n][c] * lnb[c]  (n < rows), 0 for the padding
__global__ void fold_bias_kernel(const float *__restrict__ w, const float *__restrict__ bias,
                                 const float *__restrict__ lnb, float *__restrict__ out, int rows, int cols,
                                 int rows_pad) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows_pad) return;
  float acc = 0.f;
  if (n < rows) {
    acc = bias[n];
    for (int c = 0; c < cols; ++c) acc = fmaf(w[(size_t)n * cols + c], lnb[c], acc);
  }
  out[n] = acc;
}
__global__ void scale_cols_kernel(const float *__restrict__ w, const float *__restrict__ g, float *__restrict__ out,
                                  int rows, int cols) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < rows * cols) out[e] = w[e] * g[e % cols];
}

void pack_image(bool f16, const float *src, int ld, int row0, int col0, int rv, int cv, const float *colscale,
                uint8_t *dst, int rp, int cp, cudaStream_t st) {
  if (f16)
    pack_image_kernel<true><<<ceil_div(rp * cp, 256), 256, 0, st>>>(src, ld, row0, col0, rv, cv, colscale, dst, rp, cp);
  else
    pack_image_kernel<false><<<ceil_div(rp * cp, 256), 256, 0, st>>>(src, ld, row0, col0, rv, cv, colscale, dst, rp, cp);
}

constexpr size_t kFrontSmem = (size_t)kC * kXLD * 4 + 2 * kTile144 + kW144 + sizeof(FrontVec);
constexpr size_t kApplySmem = (size_t)kTile144 + kW144;
constexpr size_t kBackSmem = (size_t)2 * kTile144 + kW144 + kBackTH * kHTile + kBackSlots * kW1c + sizeof(BackVec);
static_assert(kFrontSmem <= 226 * 1024, "tc_front shared memory");
static_assert(kBackSmem + 1024 <= 227 * 1024, "tc_back shared memory");

template <bool F16>
int forward_tc_t(const void *packed_fp32, const void *packed_tc, const float *x, float *y, int b, int t, int stride,
                 void *ws_tc, cudaStream_t st) {
  const TcWorkspace W = tc_workspace(b, t, stride);
  const BlockPack P = block_pack_view(packed_fp32, kC);
  constexpr TcPack L = tc_pack_layout();
  const uint8_t *base = static_cast<const uint8_t *>(packed_tc);
  const uint8_t *tcw = base + (F16 ? L.img_bytes : 0);
  const float *wvp = reinterpret_cast<const float *>(base + L.wvp);
  const float *bvp = reinterpret_cast<const float *>(base + L.bvp);
  const float *bqp = reinterpret_cast<const float *>(base + L.bqp);
  const float *bkp = reinterpret_cast<const float *>(base + L.bkp);
  uint8_t *ws = static_cast<uint8_t *>(ws_tc);
  float *gram = reinterpret_cast<float *>(ws + W.gram_part);
  float *beff = reinterpret_cast<float *>(ws + W.beff);
  uint8_t *weff = ws + W.weff;
  uint8_t *vn = ws + W.vn;
  unsigned short *obuf = reinterpret_cast<unsigned short *>(ws + W.obuf);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tc_front_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrontSmem);
    cudaFuncSetAttribute(tc_front1_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFront1Smem);
    cudaFuncSetAttribute(tc_apply_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplySmem);
    cudaFuncSetAttribute(tc_back_kernel<F16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBackSmem);
    cudaFuncSetAttribute(tc_back_kernel<F16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBackSmem);
    attr_done = true;
  }
  {
    LaunchScope ls(K_TC_FRONT, st);
    if (stride == 1)
      tc_front1_kernel<F16><<<dim3(W.nchunk, b), kFrThreads, kFront1Smem, st>>>(
          P, tcw, bqp, bkp, x, gram, vn, t, W.tiles, W.tiles_per_chunk, W.nchunk, 1.0f / sqrtf((float)kHS));
    else
      tc_front_kernel<F16><<<dim3(W.nchunk, b), kTcThreads, kFrontSmem, st>>>(
          P, tcw, bqp, bkp, x, gram, vn, t, W.tout, stride, W.tiles, W.tiles_per_chunk, W.nchunk,
          1.0f / sqrtf((float)kHS));
  }
  {
    LaunchScope ls(K_BLOCK_FOLD, st);
    block_fold_kernel<kC, F16 ? 2 : 1><<<dim3(b, FoldCfg<kC>::NBLK), kFoldThreads, 0, st>>>(wvp, bvp, gram, W.nchunk,
                                                                                           weff, beff, kKP, kKP);
  }
  {
    LaunchScope ls(K_TC_APPLY, st);
    tc_apply_kernel<F16><<<dim3(W.tiles, b), kApplyThreads, kApplySmem, st>>>(vn, weff, beff, obuf, W.tout, W.tiles);
  }
  {
    LaunchScope ls(K_TC_BACK, st);
    const int total = b * W.tiles;
    const int grid = min(total, num_sms());
    if (stride == 1)
      tc_back_kernel<F16, false><<<grid, kBackThreads, kBackSmem, st>>>(P, tcw, x, obuf, y, b, t, W.tout, W.tiles, g_trace_on);
    else
      tc_back_kernel<F16, true><<<grid, kBackThreads, kBackSmem, st>>>(P, tcw, x, obuf, y, b, t, W.tout, W.tiles, g_trace_on);
  }
  return check_launch("block_forward_tc");
}
}  // namespace

bool block_tc_built() { return true; }

void block_tc_trace(int on) { g_trace_on = on; }
int block_tc_trace_read(unsigned long long *out, int n) {
  if (n < 3 * kTraceLen) return OTP_ERR_ARG;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_back_trace, sizeof(unsigned long long) * 3 * kTraceLen);
  if (e != cudaSuccess) {
    set_error("block_tc_trace_read: %s", cudaGetErrorString(e));
    return OTP_ERR_CUDA;
  }
  return OTP_OK;
}

size_t block_tc_packed_bytes(int c) { return c == kC ? align_up(tc_pack_layout().total, 1024) : 0; }

int block_tc_pack(const otp_block_params *p, int c, void *packed_tc, cudaStream_t st) {
  if (c != kC) return OTP_OK;
  constexpr TcPack L = tc_pack_layout();
  uint8_t *base = static_cast<uint8_t *>(packed_tc);
  LaunchScope ls(K_PACK, st, 61);
  for (int f = 0; f < 2; ++f) {
    uint8_t *d = base + f * L.img_bytes;
    // LN_q / LN_k / LN2 affine folded in: W' = W diag(g), b' = b + W beta (fp32 vectors below)
    pack_image(f, p->q_w, kC, 0, 0, kC, kC, p->q_norm_w, d + L.wq, kKP, kKP, st);
    pack_image(f, p->k_w, kC, 0, 0, kC, kC, p->k_norm_w, d + L.wk, kKP, kKP, st);
    pack_image(f, p->proj_w, kC, 0, 0, kC, kC, nullptr, d + L.wp, kKP, kKP, st);
    for (int j = 0; j < kNChunk; ++j) {
      const int hv = min(kNH, 4 * kC - j * kNH);
      pack_image(f, p->mlp0_w, kC, j * kNH, 0, hv, kC, p->ln2_w, d + L.w1 + (size_t)j * kW1c, kNH, kKP, st);
      pack_image(f, p->mlp3_w, 4 * kC, 0, j * kNH, kC, hv, nullptr, d + L.w2 + (size_t)j * kW2c, kKP, kNH, st);
    }
  }
  scale_cols_kernel<<<ceil_div(kC * kC, 256), 256, 0, st>>>(p->v_w, p->v_norm_w, reinterpret_cast<float *>(base + L.wvp),
                                                            kC, kC);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->v_w, p->v_b, p->v_norm_b, reinterpret_cast<float *>(base + L.bvp), kC, kC, kKP);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->q_w, p->q_b, p->q_norm_b, reinterpret_cast<float *>(base + L.bqp), kC, kC, kKP);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->k_w, p->k_b, p->k_norm_b, reinterpret_cast<float *>(base + L.bkp), kC, kC, kKP);
  fold_bias_kernel<<<ceil_div(kHidPad, 256), 256, 0, st>>>(p->mlp0_w, p->mlp0_b, p->ln2_b,
                                                           reinterpret_cast<float *>(base + L.b1p), 4 * kC, kC, kHidPad);
  // biases folded into operand column 136 (tc_back): b_p into Wp, b_1 + W_1 beta_2 into the W1 chunks
  const float *b1f = reinterpret_cast<const float *>(base + L.b1p);
  for (int f = 0; f < 2; ++f) {
    uint8_t *d = base + f * L.img_bytes;
    if (f) {
      pack_bias_col_kernel<true><<<1, 256, 0, st>>>(p->proj_b, 0, kC, d + L.wp, kKP, kC);
      for (int j = 0; j < kNChunk; ++j)
        pack_bias_col_kernel<true><<<1, 64, 0, st>>>(b1f, j * kNH, kNH, d + L.w1 + (size_t)j * kW1c, kKP, kC);
    } else {
      pack_bias_col_kernel<false><<<1, 256, 0, st>>>(p->proj_b, 0, kC, d + L.wp, kKP, kC);
      for (int j = 0; j < kNChunk; ++j)
        pack_bias_col_kernel<false><<<1, 64, 0, st>>>(b1f, j * kNH, kNH, d + L.w1 + (size_t)j * kW1c, kKP, kC);
    }
  }
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
