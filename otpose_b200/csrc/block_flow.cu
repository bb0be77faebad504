// a2-a5 for the narrow flow encoder (C = 17, one head; reference model/OTPose.py:214-216,
// model/ConvVideoTransformer.py:123-184, model/blocks.py:264-279, 400-452) in the 16-bit modes:
// the positional embedding and ALL stride-1 blocks of the stem in ONE launch.
//
// One thread-block CLUSTER per clip.  Each CTA of the cluster owns a contiguous run of tokens and
// keeps their fp32 residual stream in shared memory across every block, so the (17, T) map is read
// from HBM once and written once; what remains per block is
//   * the channel Gram (a reduction over all tokens of the clip): per-warp tensor-core partial sums
//     -> fixed-order CTA sum -> fixed-order sum over the cluster through DISTRIBUTED SHARED MEMORY
//     (run-to-run bit-identical), one cluster barrier;
//   * the head re-assembly "scramble" (blocks.py:447: the (T, hs) product re-read as (C, T)), a
//     clip-wide permutation: att @ v goes to a 16-bit scratch (L2 resident, 235 KB per clip) and is
//     re-read after the second cluster barrier.
// Arithmetic: warp tiles of 16 tokens in mma.sync m16n8k16 fragment layout (IEEE half operands,
// fp32 accumulate).  C = 17 is far too narrow for a tcgen05 M128 tile chain (every GEMM here is
// 16 x 24 x 32 per tile; the kernel is bounded by its LayerNorm / depthwise / GELU lanes and the
// two cluster barriers per block, not by tensor throughput), so the warp-level MMA is the fit:
//   Gram-first attention:  G~ = sum_t [a;1][c;1]^T on the tensor core (tokens are the
//   K dimension; the operands are transposed in registers with movmatrix), then
//   S = Wq~ G~ Wk~^T, softmax and W_eff~ = att [Wv' | bv'] in fp32, once per clip and block;
//   o = W_eff~ [vn;1];  u = x + s_a (Wp~ [o2;1]);  y = u + s_m (W2 GELU(W1~ [LN2(u);1]) + b2)
// with every bias riding in the MMA against a ones column and the hidden activations passing from
// the W1 accumulators to the W2 A-operand in registers (accumulator layout == A-operand layout).
// LayerNorms are quad reductions (a token's 17 channels sit in the 4 lanes of a quad); the
// depthwise taps take the neighbour tokens from the neighbouring lanes by warp shuffle.
//
// CTA-edge tokens: each CTA also carries 8 halo tokens on either side of its run and updates them
// redundantly (masked out of the Gram and the scratch).  A halo token next to the outer edge has a
// wrong neighbour, so the error creeps inward by one token per block: 8 >= the number of blocks
// keeps every owned token exact and removes a third barrier per block.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "block_common.cuh"
#include "tc_common.cuh"

namespace otp {
namespace {
namespace cg = cooperative_groups;
using tc::pack16x2;

constexpr int FC = 17, FNP = 18;           // channels; row of the fp32 pack
constexpr int FTH = 512, FNW = FTH / 32;   // threads / warps per CTA
constexpr int FXS = 18;                    // floats per token row of the residual stream
constexpr int FHALO = 8;                   // halo tokens either side (>= blocks per launch)
constexpr int FMAXBLK = 8;
constexpr int FMAXPER = 1728;              // owned tokens per CTA (6912 / 4)
constexpr int FG = 18 * 18;                // augmented Gram
constexpr int FHID = 4 * FC;               // 68 hidden units, 9 n8 tiles
constexpr int FSTAGE = 4224;               // floats of a block's raw weight sections (staging area)

struct FlowArgs {
  BlockPack blk[FMAXBLK];
  int nblocks;
  const float *x, *pe;
  int pe_stride;
  float *y;
  float *obuf;            // [B][T * 17]: att @ v, token-major (the scramble buffer)
  int T, per, cs;
};

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
// LayerNorm (no affine) over the 17 channels of a token spread over a quad: slot s = 2j + e holds
// channel 8j + 2q + e; slots 0..3 are real channels, slot 4 is channel 16 on lanes q == 0, the rest
// is padding and comes back as 0.
__device__ __forceinline__ void ln_quad(float (&v)[6], bool q0) {
  const float v4 = q0 ? v[4] : 0.f;
  const float mu = quad_sum(v[0] + v[1] + v[2] + v[3] + v4) * (1.0f / FC);
  float d[5];
#pragma unroll
  for (int s = 0; s < 4; ++s) d[s] = v[s] - mu;
  d[4] = q0 ? v4 - mu : 0.f;
  float ss = 0.f;
#pragma unroll
  for (int s = 0; s < 5; ++s) ss = fmaf(d[s], d[s], ss);
  const float rstd = rsqrtf(quad_sum(ss) * (1.0f / FC) + 1e-5f);   // MUFU.RSQ: 2^-22.9 relative
#pragma unroll
  for (int s = 0; s < 5; ++s) v[s] = d[s] * rstd;
  v[5] = 0.f;
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {   // 8x8 b16 transpose across the warp
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// x = hi + lo as two half pairs (~22 significant bits; the conversions saturate)
__device__ __forceinline__ void split2(float x, float y, uint32_t &hi, uint32_t &lo) {
  hi = pack16x2<true>(x, y);
  const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
  lo = pack16x2<true>(x - f.x, y - f.y);
}
// A-operand fragments of two token rows (slots 0..5 = channels 8j + 2q + e): k-tile 0 = channels 0..15 (a0..a3),
// k-tile 1 = channels 16, 17 (a0, a1; a2 = a3 = 0)
struct AFrag {
  uint32_t k0[4], k1[2];
};
__device__ __forceinline__ void split_rows(const float (&r0)[6], const float (&r1)[6], AFrag &hi, AFrag &lo) {
  split2(r0[0], r0[1], hi.k0[0], lo.k0[0]);
  split2(r1[0], r1[1], hi.k0[1], lo.k0[1]);
  split2(r0[2], r0[3], hi.k0[2], lo.k0[2]);
  split2(r1[2], r1[3], hi.k0[3], lo.k0[3]);
  split2(r0[4], r0[5], hi.k1[0], lo.k1[0]);
  split2(r1[4], r1[5], hi.k1[1], lo.k1[1]);
}
// d += (hi + lo) (Bh + Bl) without the lo lo term: fp32-class products from half operands
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&h)[4], const uint32_t (&l)[4], uint2 bh, uint2 bl) {
  mma16816(d, h[0], h[1], h[2], h[3], bh);
  mma16816(d, h[0], h[1], h[2], h[3], bl);
  mma16816(d, l[0], l[1], l[2], l[3], bh);
}
__device__ __forceinline__ void mma3_k1(float (&d)[4], const uint32_t (&h)[2], const uint32_t (&l)[2], uint2 bh, uint2 bl) {
  mma16816(d, h[0], h[1], 0u, 0u, bh);
  mma16816(d, h[0], h[1], 0u, 0u, bl);
  mma16816(d, l[0], l[1], 0u, 0u, bh);
}
// GELU in fp32, x Phi(x) with the erf of blocks.py:250 to 1.5e-7 (Abramowitz-Stegun 7.1.26):
//   z = |x| / sqrt 2,  t = 1 / (1 + p z),  h = 0.5 (a1 t + ... + a5 t^5) exp(-z^2),  Phi = x < 0 ? h : 1 - h
// (no cancellation on the negative side).  The cheaper tanh / logistic fits are SMOOTH errors of 3e-5 .. 3e-4,
// i.e. the same perturbation for every token: the next block's channel Gram sums it coherently over the clip and
// its softmax exponentiates it (measured: a 1e-5 error after block 1 is 1e-4 after block 2), so the flow encoder
// is kept at fp32-class accuracy throughout -- hi + lo operand pairs in every MMA and this GELU.  Two MUFUs.
__device__ __forceinline__ float gelu_as(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float h = p * t * e;
  return x * (x < 0.f ? h : 1.0f - h);
}
// B-operand fragment image of W[n][k] (n < 8 NT, k < 16 KT): entry (kt, nt, lane) = {W[8nt+g][16kt+2q .. +1],
// W[8nt+g][16kt+2q+8 .. +9]} as half pairs
template <class F>
__device__ __forceinline__ void build_bfrag(uint2 *dst, uint2 *dst_lo, int KT, int NT, F w) {
  auto split = [](float x, float y, uint32_t &hi, uint32_t &lo) {   // x = hi + lo to ~22 bits
    hi = pack16x2<true>(x, y);
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    lo = pack16x2<true>(x - f.x, y - f.y);
  };
  for (int e = threadIdx.x; e < KT * NT * 32; e += FTH) {
    const int lane = e & 31, nt = (e >> 5) % NT, kt = (e >> 5) / NT;
    const int n = 8 * nt + (lane >> 2), k = 16 * kt + 2 * (lane & 3);
    uint2 hi, lo;
    split(w(n, k), w(n, k + 1), hi.x, lo.x);
    split(w(n, k + 8), w(n, k + 9), hi.y, lo.y);
    dst[e] = hi;
    dst_lo[e] = lo;
  }
}

__global__ void __launch_bounds__(FTH, 1) flow_encoder_kernel(const __grid_constant__ FlowArgs A) {
  extern __shared__ __align__(16) float dsm[];
  // per-block weights
  __shared__ float wq_s[FC * FNP], wk_s[FC * FNP], wv_s[FC * FNP];   // augmented [i][l], l == 17: bias
  __shared__ __align__(8) float ln1g[24], ln1b[24], dws[3][3][24], sa_s[24], sm_s[24], b2_s[24];
  __shared__ uint2 wp_f[2 * 3 * 32], weff_f[2 * 3 * 32], w1_f[2 * 9 * 32], w2_f[5 * 3 * 32];
  __shared__ uint2 wp_l[2 * 3 * 32], weff_l[2 * 3 * 32], w1_l[2 * 9 * 32], w2_l[5 * 3 * 32];   // lo terms of the same images
  // Gram reduction / fold
  __shared__ float gpart[FG], Gs[FG], m1_s[FC * FNP], s_s[FC * FNP], weff_s[FC * FNP];

  cg::cluster_group cluster = cg::this_cluster();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const bool q0 = q == 0;
  const int T = A.T, per = A.per, cs = A.cs;
  const int b = blockIdx.x / cs, rank = blockIdx.x % cs;
  const int o0 = rank * per, o1 = min(T, o0 + per);   // owned tokens [o0, o1) (possibly empty)
  const int e0 = o0 - FHALO;                          // token of local row 0
  const int nrows = per + 2 * FHALO, ntile = nrows / 16;
  float *xs = dsm;                                    // [nrows][FXS] (+ one spare row)
  float *wpart = dsm + (size_t)(nrows + 1) * FXS;     // [FNW][FG], then the weight staging area [FSTAGE]

  // ---- residual stream: x + positional embedding (ConvVideoTransformer.py:147-157) ----
  {
    const float *xb = A.x + (size_t)b * FC * T;
    for (int r = threadIdx.x; r < nrows; r += FTH) {   // thread = token: 34 coalesced loads in flight
      const int tok = e0 + r;
      const bool ok = tok >= 0 && tok < T;
      float v[FC], pe[FC];
#pragma unroll
      for (int c = 0; c < FC; ++c) {
        v[c] = ok ? __ldg(xb + (size_t)c * T + tok) : 0.f;
        pe[c] = (ok && A.pe) ? __ldg(A.pe + (size_t)c * A.pe_stride + tok) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < FC; ++c) xs[r * FXS + c] = v[c] + pe[c];
    }
    for (int r = threadIdx.x; r <= nrows; r += FTH) xs[r * FXS + 17] = 0.f;
    if (threadIdx.x < FXS) xs[nrows * FXS + threadIdx.x] = 0.f;
  }

  // LN1(x) of local row r with the affine applied: what the depthwise convs see; 0 outside the clip
  // (their zero padding) and outside the rows this CTA carries
  auto row_y = [&](int r, float (&y)[6]) {
    const int tok = e0 + r;
    const bool ok = r >= 0 && r < nrows && tok >= 0 && tok < T;
    const float *xr = xs + min(max(r, 0), nrows - 1) * FXS + 2 * q;
    float v[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      // slots 4, 5 of the lanes q != 0 are padding: no load (it would reach into the next token's row)
      const float2 t = (j < 2 || q0) ? *reinterpret_cast<const float2 *>(xr + 8 * j) : make_float2(0.f, 0.f);
      v[2 * j] = t.x, v[2 * j + 1] = t.y;
    }
    ln_quad(v, q0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float2 gg = *reinterpret_cast<const float2 *>(ln1g + 8 * j + 2 * q);
      const float2 bb = *reinterpret_cast<const float2 *>(ln1b + 8 * j + 2 * q);
      y[2 * j] = ok ? fmaf(v[2 * j], gg.x, bb.x) : 0.f;
      y[2 * j + 1] = ok ? fmaf(v[2 * j + 1], gg.y, bb.y) : 0.f;
    }
  };
  // the tile's LN1 rows of this lane (rows g and g + 8) and their neighbours: the row before the tile and
  // the row after it are computed by the g == 0 / g == 7 quads, everything else arrives by shuffle
  struct Taps {
    float y0[6], y1[6], p0[6], n0[6], p1[6], n1[6];
  };
  auto load_taps = [&](int tile, Taps &t) {
    const int r0 = tile * 16 + g;
    float yx[6];
    row_y(r0, t.y0);
    row_y(r0 + 8, t.y1);
    row_y(g == 0 ? tile * 16 - 1 : (g == 7 ? tile * 16 + 16 : r0), yx);
    const int up = (lane - 4) & 31, dn = (lane + 4) & 31;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const float a = __shfl_sync(0xffffffffu, t.y0[s], up), bq = __shfl_sync(0xffffffffu, t.y1[s], up);
      const float c = __shfl_sync(0xffffffffu, t.y0[s], dn), d = __shfl_sync(0xffffffffu, t.y1[s], dn);
      t.p0[s] = g > 0 ? a : yx[s];    // row g - 1 (g == 0: the row before the tile)
      t.p1[s] = g > 0 ? bq : a;       // row g + 7 (g == 0: row 7 lives in the g == 7 quad's first half)
      t.n0[s] = g < 7 ? c : d;        // row g + 1 (g == 7: row 8 lives in the g == 0 quad's second half)
      t.n1[s] = g < 7 ? d : yx[s];    // row g + 9 (g == 7: the row after the tile)
    }
  };
  // LayerNorm (no affine) of the depthwise conv `cv` (0 query, 1 key, 2 value) of both rows
  auto dw_ln = [&](const Taps &t, int cv, float (&r0)[6], float (&r1)[6]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float2 w0 = *reinterpret_cast<const float2 *>(&dws[cv][0][8 * j + 2 * q]);
      const float2 w1 = *reinterpret_cast<const float2 *>(&dws[cv][1][8 * j + 2 * q]);
      const float2 w2 = *reinterpret_cast<const float2 *>(&dws[cv][2][8 * j + 2 * q]);
      r0[2 * j] = fmaf(w0.x, t.p0[2 * j], fmaf(w1.x, t.y0[2 * j], w2.x * t.n0[2 * j]));
      r0[2 * j + 1] = fmaf(w0.y, t.p0[2 * j + 1], fmaf(w1.y, t.y0[2 * j + 1], w2.y * t.n0[2 * j + 1]));
      r1[2 * j] = fmaf(w0.x, t.p1[2 * j], fmaf(w1.x, t.y1[2 * j], w2.x * t.n1[2 * j]));
      r1[2 * j + 1] = fmaf(w0.y, t.p1[2 * j + 1], fmaf(w1.y, t.y1[2 * j + 1], w2.y * t.n1[2 * j + 1]));
    }
    ln_quad(r0, q0);
    ln_quad(r1, q0);
  };
  const int own_lo = FHALO, own_hi = FHALO + max(0, o1 - o0);   // owned local rows [own_lo, own_hi)

  // staging area of a block's raw fp32 weight sections (same offsets for every block), filled by 4-byte cp.async
  float *st = wpart + FNW * FG;
  constexpr int o_wq = 0, o_wk = o_wq + FC * FNP, o_wv = o_wk + FC * FNP, o_wp = o_wv + FC * FC, o_w1 = o_wp + FC * FNP;
  constexpr int o_w2 = o_w1 + 4 * FC * FNP, o_l1w = o_w2 + 4 * FC * FNP, o_l1b = o_l1w + FC, o_l2w = o_l1b + FC;
  constexpr int o_l2b = o_l2w + FC, o_qnw = o_l2b + FC, o_qnb = o_qnw + FC, o_knw = o_qnb + FC, o_knb = o_knw + FC;
  constexpr int o_vnw = o_knb + FC, o_vnb = o_vnw + FC, o_dq = o_vnb + FC, o_dk = o_dq + 3 * FC, o_dv = o_dk + 3 * FC;
  constexpr int o_bq = o_dv + 3 * FC, o_bk = o_bq + FNP, o_bv = o_bk + FNP, o_bp = o_bv + FNP, o_b2 = o_bp + FNP;
  constexpr int o_sa = o_b2 + FNP, o_sm = o_sa + FNP, o_b1 = o_sm + FNP;
  static_assert(o_b1 + 4 * FNP <= FSTAGE, "staging area");
  auto stage_block = [&](const BlockPack &Q) {
    auto take = [&](int o, const float *src, int n) {
      for (int i = threadIdx.x; i < n; i += FTH)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(st + o + i)), "l"(src + i) : "memory");
    };
    take(o_wq, Q.wqT, FC * FNP), take(o_wk, Q.wkT, FC * FNP), take(o_wv, Q.wv, FC * FC), take(o_wp, Q.wpT, FC * FNP);
    take(o_w1, Q.w1T, 4 * FC * FNP), take(o_w2, Q.w2T, 4 * FC * FNP);
    take(o_l1w, Q.ln1_w, FC), take(o_l1b, Q.ln1_b, FC), take(o_l2w, Q.ln2_w, FC), take(o_l2b, Q.ln2_b, FC);
    take(o_qnw, Q.qn_w, FC), take(o_qnb, Q.qn_b, FC), take(o_knw, Q.kn_w, FC), take(o_knb, Q.kn_b, FC);
    take(o_vnw, Q.vn_w, FC), take(o_vnb, Q.vn_b, FC);
    take(o_dq, Q.dwq, 3 * FC), take(o_dk, Q.dwk, 3 * FC), take(o_dv, Q.dwv, 3 * FC);
    take(o_bq, Q.bq, FNP), take(o_bk, Q.bk, FNP), take(o_bv, Q.bv, FNP), take(o_bp, Q.bp, FNP);
    take(o_b2, Q.b2, FNP), take(o_sa, Q.sa, FNP), take(o_sm, Q.sm, FNP), take(o_b1, Q.b1, 4 * FNP);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_block(A.blk[0]);

  for (int blk = 0; blk < A.nblocks; ++blk) {
    __syncthreads();   // the previous block's tiles are done with the weights and with xs
    // ---- this block's weights: the raw fp32 sections of the pack were prefetched into the staging area by cp.async
    //      (under the previous block's tile phases); the fp32 vectors, the augmented q / k / v projections and the
    //      B-operand fragments are formed from it ----
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    for (int e = threadIdx.x; e < 24; e += FTH) {
      const bool in = e < FC;
      ln1g[e] = in ? st[o_l1w + e] : 0.f;
      ln1b[e] = in ? st[o_l1b + e] : 0.f;
      sa_s[e] = in ? st[o_sa + e] : 0.f;
      sm_s[e] = in ? st[o_sm + e] : 0.f;
      b2_s[e] = in ? st[o_b2 + e] : 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dws[0][k][e] = in ? st[o_dq + 3 * e + k] : 0.f;
        dws[1][k][e] = in ? st[o_dk + 3 * e + k] : 0.f;
        dws[2][k][e] = in ? st[o_dv + 3 * e + k] : 0.f;
      }
    }
    for (int e = threadIdx.x; e < 3 * FC * FNP; e += FTH) {
      // W~[i][l] = W[i][l] g[l] (l < 17),  W~[i][17] = b[i] + sum_l W[i][l] beta[l]:  W (g * a^ + beta) + b
      const int m = e / (FC * FNP), i = (e / FNP) % FC, l = e % FNP;
      const int gam = m == 0 ? o_qnw : (m == 1 ? o_knw : o_vnw), bet = m == 0 ? o_qnb : (m == 1 ? o_knb : o_vnb);
      const int bias = m == 0 ? o_bq : (m == 1 ? o_bk : o_bv);
      auto w = [&](int c) { return m == 0 ? st[o_wq + c * FNP + i] : (m == 1 ? st[o_wk + c * FNP + i] : st[o_wv + i * FC + c]); };
      float v;
      if (l < FC) {
        v = w(l) * st[gam + l];
      } else {
        v = st[bias + i];
        for (int c = 0; c < FC; ++c) v = fmaf(w(c), st[bet + c], v);
      }
      (m == 0 ? wq_s : (m == 1 ? wk_s : wv_s))[i * FNP + l] = v;
    }
    build_bfrag(wp_f, wp_l, 2, 3, [&](int n, int k) {   // Wp~[n][k]: proj, bias in column 17
      if (n >= FC || k > FC) return 0.f;
      return k < FC ? st[o_wp + k * FNP + n] : st[o_bp + n];
    });
    build_bfrag(w1_f, w1_l, 2, 9, [&](int n, int k) {   // W1~[n][k] = W1[n][k] g2[k]; column 17: b1[n] + sum_c W1[n][c] beta2[c]
      if (n >= FHID || k > FC) return 0.f;
      const float *w1 = st + o_w1 + (n / FC) * FC * FNP + n % FC;   // W1[n][c] = w1[c * FNP]
      if (k < FC) return w1[k * FNP] * st[o_l2w + k];
      float v = st[o_b1 + (n / FC) * FNP + n % FC];
      for (int c = 0; c < FC; ++c) v = fmaf(w1[c * FNP], st[o_l2b + c], v);
      return v;
    });
    build_bfrag(w2_f, w2_l, 5, 3, [&](int n, int k) {   // W2[n][k], k = hidden unit
      if (n >= FC || k >= FHID) return 0.f;
      return st[o_w2 + (k / FC) * FC * FNP + (k % FC) * FNP + n];
    });
    __syncthreads();
    if (blk + 1 < A.nblocks) stage_block(A.blk[blk + 1]);   // lands under this block's tile phases

    // ================= front: G~ = sum over the owned tokens of [a;1][c;1]^T =================
    {
      float G[2][3][4];
#pragma unroll
      for (int i = 0; i < 24; ++i) (&G[0][0][0])[i] = 0.f;
      for (int tile = warp; tile < ntile; tile += FNW) {
        if (tile * 16 + 16 <= own_lo || tile * 16 >= own_hi) continue;   // no owned row (warp-uniform)
        Taps t;
        load_taps(tile, t);
        const int r0 = tile * 16 + g;
        const bool w0 = r0 >= own_lo && r0 < own_hi, w1 = r0 + 8 >= own_lo && r0 + 8 < own_hi;
        // Operands as hi + lo half pairs (a = hi + lo to ~22 bits): the Gram is a sum over thousands of tokens
        // that the softmax then exponentiates, so it is kept at fp32-class accuracy with three MMAs
        // (hi hi + hi lo + lo hi) instead of one; the dropped lo lo term is below 2^-22.
        float a0[6], a1[6];
        uint32_t Ra[2][3], Ral[2][3], Rc[2][3], Rcl[2][3];
        auto split = [](float x, float y, uint32_t &hi, uint32_t &lo) {
          hi = pack16x2<true>(x, y);
          const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
          lo = pack16x2<true>(x - f.x, y - f.y);
        };
        dw_ln(t, 0, a0, a1);
        if (q0) a0[5] = 1.f, a1[5] = 1.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // rows outside the owned run contribute nothing
          uint32_t h0, l0, h1, l1;
          split(a0[2 * j], a0[2 * j + 1], h0, l0);
          split(a1[2 * j], a1[2 * j + 1], h1, l1);
          Ra[0][j] = movm_t(w0 ? h0 : 0u), Ral[0][j] = movm_t(w0 ? l0 : 0u);
          Ra[1][j] = movm_t(w1 ? h1 : 0u), Ral[1][j] = movm_t(w1 ? l1 : 0u);
        }
        dw_ln(t, 1, a0, a1);
        if (q0) a0[5] = 1.f, a1[5] = 1.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t h0, l0, h1, l1;
          split(a0[2 * j], a0[2 * j + 1], h0, l0);
          split(a1[2 * j], a1[2 * j + 1], h1, l1);
          Rc[0][j] = movm_t(h0), Rcl[0][j] = movm_t(l0);
          Rc[1][j] = movm_t(h1), Rcl[1][j] = movm_t(l1);
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const uint2 bh = make_uint2(Rc[0][nt], Rc[1][nt]), bl = make_uint2(Rcl[0][nt], Rcl[1][nt]);
          // channels 0..15 of a
          mma16816(G[0][nt], Ra[0][0], Ra[0][1], Ra[1][0], Ra[1][1], bh);
          mma16816(G[0][nt], Ra[0][0], Ra[0][1], Ra[1][0], Ra[1][1], bl);
          mma16816(G[0][nt], Ral[0][0], Ral[0][1], Ral[1][0], Ral[1][1], bh);
          // channels 16, 17 (ones)
          mma16816(G[1][nt], Ra[0][2], 0u, Ra[1][2], 0u, bh);
          mma16816(G[1][nt], Ra[0][2], 0u, Ra[1][2], 0u, bl);
          mma16816(G[1][nt], Ral[0][2], 0u, Ral[1][2], 0u, bh);
        }
      }
      float *wp = wpart + warp * FG;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 16 * mt + g + 8 * (i >> 1), col = 8 * nt + 2 * q + (i & 1);
            if (row < 18 && col < 18) wp[row * 18 + col] = G[mt][nt][i];
          }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < FG; e += FTH) {   // fixed order over the warps
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < FNW; ++w) s += wpart[w * FG + e];
      gpart[e] = s;
    }
    cluster.sync();   // barrier 1: every CTA's partial Gram is in its shared memory
    for (int e = threadIdx.x; e < FG; e += FTH) {   // fixed order over the cluster (distributed shared memory)
      float s = 0.f;
      for (int r = 0; r < cs; ++r) s += cluster.map_shared_rank(gpart, r)[e];
      Gs[e] = s;
    }
    __syncthreads();
    // ---- fold (every CTA of the clip, redundantly): S = Wq~ G~ Wk~^T / sqrt(hs), softmax, W_eff~ = att Wv~ ----
    if (threadIdx.x < FC * FNP) {
      const int i = threadIdx.x / FNP, m = threadIdx.x % FNP;
      float s = 0.f;
#pragma unroll
      for (int l = 0; l < FNP; ++l) s = fmaf(wq_s[i * FNP + l], Gs[l * 18 + m], s);
      m1_s[i * FNP + m] = s;
    }
    __syncthreads();
    if (threadIdx.x < FC * FC) {
      const int i = threadIdx.x / FC, j = threadIdx.x % FC;
      float s = 0.f;
#pragma unroll
      for (int m = 0; m < FNP; ++m) s = fmaf(m1_s[i * FNP + m], wk_s[j * FNP + m], s);
      s_s[i * FNP + j] = s * 0.24253562503633297f;   // 1 / sqrt(17)
    }
    __syncthreads();
    for (int i = warp; i < FC; i += FNW) {
      float *row = s_s + i * FNP;
      const float v = lane < FC ? row[lane] : -3.402823466e38f;
      const float mx = warp_max(v);
      const float ex = lane < FC ? expf(v - mx) : 0.f;
      const float inv = 1.0f / warp_sum(ex);
      if (lane < FC) row[lane] = ex * inv;
    }
    __syncthreads();
    if (threadIdx.x < FC * FNP) {
      const int i = threadIdx.x / FNP, k = threadIdx.x % FNP;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < FC; ++j) s = fmaf(s_s[i * FNP + j], wv_s[j * FNP + k], s);
      weff_s[i * FNP + k] = s;
    }
    __syncthreads();
    build_bfrag(weff_f, weff_l, 2, 3, [&](int n, int k) { return (n < FC && k < FNP) ? weff_s[n * FNP + k] : 0.f; });
    __syncthreads();

    // ================= apply: o = W_eff~ [vn;1] of the owned tokens -> scramble buffer =================
    {
      float *ob = A.obuf + (size_t)b * T * FC;
      for (int tile = warp; tile < ntile; tile += FNW) {
        if (tile * 16 + 16 <= own_lo || tile * 16 >= own_hi) continue;
        Taps t;
        load_taps(tile, t);
        float v0[6], v1[6];
        dw_ln(t, 2, v0, v1);
        if (q0) v0[5] = 1.f, v1[5] = 1.f;
        AFrag vh, vl;
        split_rows(v0, v1, vh, vl);
        const int r0 = tile * 16 + g;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          float o[4] = {0.f, 0.f, 0.f, 0.f};
          mma3(o, vh.k0, vl.k0, weff_f[(0 * 3 + nt) * 32 + lane], weff_l[(0 * 3 + nt) * 32 + lane]);
          mma3_k1(o, vh.k1, vl.k1, weff_f[(1 * 3 + nt) * 32 + lane], weff_l[(1 * 3 + nt) * 32 + lane]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = r0 + 8 * (i >> 1), col = 8 * nt + 2 * q + (i & 1);
            if (col < FC && r >= own_lo && r < own_hi) ob[(size_t)(e0 + r) * FC + col] = o[i];
          }
        }
      }
    }
    __threadfence();
    cluster.sync();   // barrier 2: the clip's att @ v is complete in the scramble buffer

    // ================= back: proj, residual, LN2, MLP, residual (all carried rows) =================
    {
      const float *ob = A.obuf + (size_t)b * T * FC;
      const bool last = blk + 1 == A.nblocks;
      float *yb = A.y + (size_t)b * FC * T;
      for (int tile = warp; tile < ntile; tile += FNW) {
        const int r0 = tile * 16 + g, r1 = r0 + 8;
        const int t0 = e0 + r0, t1 = e0 + r1;
        const bool ok0 = t0 >= 0 && t0 < T, ok1 = t1 >= 0 && t1 < T;
        // o2[t][c] = flat[c * T + t]  (blocks.py:447); column 17 = 1 (bias)
        float h0[6], h1[6];
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          const int c = 8 * (s >> 1) + 2 * q + (s & 1);
          h0[s] = (ok0 && c < FC) ? __ldcg(ob + (size_t)c * T + t0) : 0.f;
          h1[s] = (ok1 && c < FC) ? __ldcg(ob + (size_t)c * T + t1) : 0.f;
        }
        if (q0) h0[5] = 1.f, h1[5] = 1.f;
        float u0[6], u1[6];
        {
          AFrag oh, ol;
          split_rows(h0, h1, oh, ol);
          const float *x0 = xs + r0 * FXS + 2 * q, *x1 = xs + r1 * FXS + 2 * q;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            float p[4] = {0.f, 0.f, 0.f, 0.f};
            mma3(p, oh.k0, ol.k0, wp_f[(0 * 3 + nt) * 32 + lane], wp_l[(0 * 3 + nt) * 32 + lane]);
            mma3_k1(p, oh.k1, ol.k1, wp_f[(1 * 3 + nt) * 32 + lane], wp_l[(1 * 3 + nt) * 32 + lane]);
            const float2 sa = *reinterpret_cast<const float2 *>(sa_s + 8 * nt + 2 * q);
            const bool real = nt < 2 || q0;   // padding slots: no load (the next token's row belongs to another warp)
            const float2 xa = real ? *reinterpret_cast<const float2 *>(x0 + 8 * nt) : make_float2(0.f, 0.f);
            const float2 xb2 = real ? *reinterpret_cast<const float2 *>(x1 + 8 * nt) : make_float2(0.f, 0.f);
            u0[2 * nt] = fmaf(sa.x, p[0], xa.x), u0[2 * nt + 1] = fmaf(sa.y, p[1], xa.y);
            u1[2 * nt] = fmaf(sa.x, p[2], xb2.x), u1[2 * nt + 1] = fmaf(sa.y, p[3], xb2.y);
          }
        }
        AFrag lh, ll;
        {
          float l0[6], l1[6];
#pragma unroll
          for (int s = 0; s < 6; ++s) l0[s] = u0[s], l1[s] = u1[s];
          ln_quad(l0, q0);
          ln_quad(l1, q0);
          if (q0) l0[5] = 1.f, l1[5] = 1.f;
          split_rows(l0, l1, lh, ll);
        }
        float out[3][4];
#pragma unroll
        for (int i = 0; i < 12; ++i) (&out[0][0])[i] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 5; ++kt) {   // 16 hidden units at a time: W1 -> GELU -> W2 without leaving registers
          uint32_t ha[4] = {0u, 0u, 0u, 0u}, hl[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int nn = 2 * kt + hh;
            if (nn < 9) {
              float hd[4] = {0.f, 0.f, 0.f, 0.f};
              mma3(hd, lh.k0, ll.k0, w1_f[(0 * 9 + nn) * 32 + lane], w1_l[(0 * 9 + nn) * 32 + lane]);
              mma3_k1(hd, lh.k1, ll.k1, w1_f[(1 * 9 + nn) * 32 + lane], w1_l[(1 * 9 + nn) * 32 + lane]);
              split2(gelu_as(hd[0]), gelu_as(hd[1]), ha[2 * hh], hl[2 * hh]);
              split2(gelu_as(hd[2]), gelu_as(hd[3]), ha[2 * hh + 1], hl[2 * hh + 1]);
            }
          }
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) mma3(out[nt], ha, hl, w2_f[(kt * 3 + nt) * 32 + lane], w2_l[(kt * 3 + nt) * 32 + lane]);
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const float2 sm = *reinterpret_cast<const float2 *>(sm_s + 8 * nt + 2 * q);
          const float2 b2 = *reinterpret_cast<const float2 *>(b2_s + 8 * nt + 2 * q);
          const int c = 8 * nt + 2 * q;
          float2 y0 = make_float2(fmaf(sm.x, out[nt][0] + b2.x, u0[2 * nt]), fmaf(sm.y, out[nt][1] + b2.y, u0[2 * nt + 1]));
          float2 y1 = make_float2(fmaf(sm.x, out[nt][2] + b2.x, u1[2 * nt]), fmaf(sm.y, out[nt][3] + b2.y, u1[2 * nt + 1]));
          if (!ok0) y0 = make_float2(0.f, 0.f);
          if (!ok1) y1 = make_float2(0.f, 0.f);
          if (c < FXS) {
            if (c + 1 >= FC) y0.y = 0.f, y1.y = 0.f;   // padding column 17 stays 0
            *reinterpret_cast<float2 *>(xs + r0 * FXS + c) = y0;
            *reinterpret_cast<float2 *>(xs + r1 * FXS + c) = y1;
          }
          if (last) {
            if (ok0 && r0 >= own_lo && r0 < own_hi) {
              if (c < FC) yb[(size_t)c * T + t0] = y0.x;
              if (c + 1 < FC) yb[(size_t)(c + 1) * T + t0] = y0.y;
            }
            if (ok1 && r1 >= own_lo && r1 < own_hi) {
              if (c < FC) yb[(size_t)c * T + t1] = y1.x;
              if (c + 1 < FC) yb[(size_t)(c + 1) * T + t1] = y1.y;
            }
          }
        }
      }
    }
  }
  cluster.sync();   // no CTA leaves while a peer may still read its partial Gram
}

// cluster size: the smallest power of two whose runs fit a CTA, doubled while the grid is short of one wave
// and the runs stay >= 96 tokens
int flow_cluster_size(int b, int t) {
  int cs = 1;
  while (cs < 8 && ceil_div(t, cs) > FMAXPER) cs *= 2;
  if (ceil_div(t, cs) > FMAXPER) return 0;
  while (cs < 8 && (long long)b * cs < 128 && t / (2 * cs) >= 96) cs *= 2;
  return cs;
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_flow_encoder_supported(int c, int n_head, int t, int nblocks) {
  return (c == FC && n_head == 1 && t > 0 && nblocks >= 1 && nblocks <= FMAXBLK && nblocks <= FHALO &&
          flow_cluster_size(1 << 20, t) > 0) ? 1 : 0;
}

extern "C" size_t otp_flow_encoder_workspace_bytes(int b, int t) {
  if (b <= 0 || t <= 0) return 0;
  return align_up((size_t)b * t * FC * 4 + 64, 256);
}

extern "C" int otp_flow_encoder_forward(const void *const *packed_blocks, int nblocks, const float *x, const float *pe,
                                        int pe_stride, float *y, int b, int t, void *workspace,
                                        size_t workspace_bytes, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && t > 0 && packed_blocks != nullptr);
  if (!otp_flow_encoder_supported(FC, 1, t, nblocks)) {
    set_error("otp_flow_encoder_forward: t=%d / %d blocks not built (t <= %d, blocks <= %d)", t, nblocks, 8 * FMAXPER,
              FHALO);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && y && workspace && x != y);
  OTP_REQUIRE(pe == nullptr || pe_stride >= t);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (workspace_bytes < otp_flow_encoder_workspace_bytes(b, t)) {
    set_error("otp_flow_encoder_forward: workspace of %zu B, need %zu B", workspace_bytes,
              otp_flow_encoder_workspace_bytes(b, t));
    return OTP_ERR_WORKSPACE;
  }
  FlowArgs A{};
  for (int i = 0; i < nblocks; ++i) {
    OTP_REQUIRE(packed_blocks[i] != nullptr);
    A.blk[i] = block_pack_view(packed_blocks[i], FC);
  }
  A.nblocks = nblocks;
  A.x = x, A.pe = pe, A.pe_stride = pe_stride, A.y = y;
  A.obuf = static_cast<float *>(workspace);
  A.T = t;
  A.cs = flow_cluster_size(b, t);
  A.per = ceil_div(ceil_div(t, A.cs), 16) * 16;
  const int nrows = A.per + 2 * FHALO;
  const size_t smem = ((size_t)(nrows + 1) * FXS + (size_t)FNW * FG + FSTAGE) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr;
  if (attr.first()) {
    const size_t lim = ((size_t)(FMAXPER + 2 * FHALO + 1) * FXS + (size_t)FNW * FG + FSTAGE) * sizeof(float);
    if (!set_max_smem(flow_encoder_kernel, lim, "flow_encoder_kernel")) return OTP_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(b * A.cs));
  cfg.blockDim = dim3(FTH);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)A.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  LaunchScope ls(K_FLOW_ENCODER, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, flow_encoder_kernel, A);
  if (e != cudaSuccess) {
    set_error("flow_encoder_kernel: %s", cudaGetErrorString(e));
    return OTP_ERR_CUDA;
  }
  return check_launch("flow_encoder_kernel");
}
