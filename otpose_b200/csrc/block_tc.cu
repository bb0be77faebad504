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
//             Warp-specialised (compute / MMA issuer / halo warps): block_tc_front.cuh.
//   fold      softmax + W_eff = softmax(S) W_v, emitted as an operand image (block_fold.cuh).
//   tc_apply  o = W_eff vn + b_eff : one M128 x N144 x K144 UMMA per tile, stored
//             token-major 16-bit == the reference's scramble buffer.
//   tc_back   proj UMMA -> u = skip + s_a(.) -> LN2 (registers) -> 9 x { W1 chunk UMMA (N64) ->
//             GELU -> H tile -> W2 chunk UMMA accumulating in TMEM } -> y = u + s_m(.)
//             Persistent, warp-specialised (TMA producer / MMA issuer / 16 epilogue warps):
//             block_tc_back.cuh.
//
// fp32 is kept for the residual stream, all LayerNorm statistics, the Gram accumulation,
// softmax and every epilogue; only UMMA operands are 16-bit (template F16: false =
// bfloat16, true = IEEE half -- same speed, 8x finer rounding; every operand here is
// O(1)..O(100), far inside the half range).  The affine part of LN_q/LN_k/LN_v/LN2 is
// folded into the following GEMM's weights and bias when the weights are packed, so the
// staged operands are plain (x - mean) * rstd; b_p and b_1 ride in the MMAs against a ones
// column of the activation tile.
//
// TMEM lane rule: warp w reads lanes 32*(w%4).., so a thread's token is 32*(w%4) + lane and the
// warps sharing a lane quarter split the accumulator columns (thirds of 48 in tc_front, quarters
// of 36 in tc_back, halves of 72 in tc_apply).
#include <string.h>

#include "block_common.cuh"
#include "block_fold.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;

namespace {
constexpr int kC = 136, kHS = 68, kKP = 144;
constexpr int kTM = 128;        // tokens per tile == UMMA M
constexpr int kApplyThreads = 256;
constexpr uint32_t kCS = 128;                 // byte stride between 8-element K chunks
constexpr uint32_t kRS144 = (kKP / 8) * 128;  // byte stride between 8-row groups, K = 144
constexpr uint32_t kTile144 = (kTM / 8) * kRS144;  // 36864  [128][144]
constexpr uint32_t kW144 = (kKP / 8) * kRS144;     // 41472  [144][144]
constexpr int kNH = 144, kNChunk = 4, kHidPad = kNH * kNChunk;   // hidden 544 -> 576, 4 chunks of 144
constexpr int kPieceK = 48;                         // K extent of one streamed weight piece (3 UMMA K-steps)
constexpr uint32_t kRS48 = (kPieceK / 8) * 128;     // 768  row-group stride of a [rows][48] piece
constexpr uint32_t kPiece = (kKP / 8) * kRS48;      // 13824  [144][48]
constexpr int kPiecesPerImage = kKP / kPieceK;      // 3
constexpr int kNumPieces = kPiecesPerImage * (1 + 2 * kNChunk);   // Wp, 4 x W1 chunk, 4 x W2 chunk = 27

// ---- packed tensor-core weights of one block (bytes) ----
// Three operand-image sets [bf16 hi | fp16 | bf16 lo], then fp32 folded vectors / matrices.  A set is an
// array of kNumPieces weight PIECES ([144 rows][48 K], the unit tc_back streams through its ring): Wp
// (pieces 0..2), W1 chunk c (3 + 3c ..), W2 chunk c (15 + 3c ..); a piece index addresses the same offset
// in every set.  The bf16 "lo" set holds bf16(w - bf16(w)) of the W1 / W2 pieces only (it starts at piece
// 3): in bfloat16 mode tc_back runs every MLP UMMA chain twice (hi, then lo into the same accumulator)
// because a weight's rounding error is coherent across tokens (see block_tc_front.cuh;
// scripts/emulate_operand_rounding.py).
struct TcPack {
  size_t wp, w1 /* [4][3] */, w2 /* [4][3] */, img_bytes /* one operand format */, lo /* bf16 lo set: W1, W2 only */;
  size_t wvp /* fp32 [136][136] Wv * g_v */, bvp /* [144] */, bqp, bkp /* [144] */, b1p /* [576] */;
  size_t wqaT, wkaT /* fp32 [144][144] transposed augmented Wq~ / Wk~ (gram_project_kernel) */, total;
};
constexpr TcPack tc_pack_layout() {
  TcPack p{};
  p.wp = 0;
  p.w1 = p.wp + (size_t)kPiecesPerImage * kPiece;
  p.w2 = p.w1 + (size_t)kNChunk * kPiecesPerImage * kPiece;
  p.img_bytes = (size_t)kNumPieces * kPiece;
  p.lo = 2 * p.img_bytes;                         // + (piece - 3) * kPiece addresses a lo piece
  p.wvp = p.lo + (size_t)(kNumPieces - kPiecesPerImage) * kPiece;
  p.bvp = p.wvp + (size_t)kC * kC * 4;
  p.bqp = p.bvp + kKP * 4;
  p.bkp = p.bqp + kKP * 4;
  p.b1p = p.bkp + kKP * 4;
  p.wqaT = p.b1p + kHidPad * 4;
  p.wkaT = p.wqaT + (size_t)kKP * kKP * 4;
  p.total = p.wkaT + (size_t)kKP * kKP * 4;
  return p;
}

constexpr int kGramBlocks = 8;   // partial S per clip written by gram_project_kernel (== kGpBlocks)
struct TcWorkspace {
  size_t gram_part, spart, beff, weff, vn, obuf, total;
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
  w.gram_part = take((size_t)b * w.nchunk * kKP * kKP * 4);      // per-chunk partial G~ (137 rows used)
  w.spart = take((size_t)b * kGramBlocks * kC * kHS * 4);        // per-column-block partial S
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
    // saturating conversion: |x| > 65504 -> +-65504; from there x*x and x*p overflow to +-inf, tanh(+-inf) =
    // +-1 and the result is x (or 0): finite for every finite input
    const uint32_t xu = pack16x2<true>(a, b);
    const __half2 x = *reinterpret_cast<const __half2 *>(&xu);
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
// 36 consecutive columns (16 + 16 + 4), one wait
__device__ __forceinline__ void tmem_ld36(uint32_t taddr, float (&v)[36]) {
  uint32_t r0[16], r1[16], r2[4];
  tmem_ld16_nw(taddr, r0);
  tmem_ld16_nw(taddr + 16, r1);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3])
               : "r"(taddr + 32)
               : "memory");
  tmem_wait_ld();
  reg_fence16(r0);
  reg_fence16(r1);
  asm volatile("" : "+r"(r2[0]), "+r"(r2[1]), "+r"(r2[2]), "+r"(r2[3])::"memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __uint_as_float(r0[i]);
    v[16 + i] = __uint_as_float(r1[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[32 + i] = __uint_as_float(r2[i]);
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

// ------------------------------------------------------------------ tc_apply
// One CTA per (clip, run of `tiles_per_cta` consecutive tiles), two resident per SM: W_eff, the TMEM
// accumulator and the barrier are set up once per CTA; the next tile's vn image streams in with cp.async
// while the current tile's epilogue runs (the operand buffer is dead as soon as the UMMAs have completed).
constexpr uint32_t kApplyStage = 2 * kTM * kHS * 2;   // [2 heads][128 tokens][68] 16-bit output tile
template <bool F16>
__global__ void __launch_bounds__(kApplyThreads, 2)
tc_apply_kernel(const uint8_t *__restrict__ vn_img, const uint8_t *__restrict__ weff_img,
                const float *__restrict__ beff, unsigned short *__restrict__ obuf, int Tout, int tiles,
                int tiles_per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *a = smem;
  uint8_t *w = smem + kTile144;
  unsigned short *stage = reinterpret_cast<unsigned short *>(w + kW144);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sb[kKP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q4 = warp & 3, half = warp >> 2;
  const int tok = q4 * 32 + lane;
  const int b = blockIdx.y;
  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(tiles, tile_begin + tiles_per_cta);
  if (tile_begin >= tile_end) return;
  const uint8_t *vb = vn_img + (size_t)b * tiles * kTile144;
  cp_async_block(a, vb + (size_t)tile_begin * kTile144, kTile144, kApplyThreads);
  cp_async_block(w, weff_img + (size_t)b * kW144, kW144, kApplyThreads);
  cp_async_commit();
  for (int n = threadIdx.x; n < kKP; n += kApplyThreads) sb[n] = beff[(size_t)b * kKP + n];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  unsigned short *ob = obuf + (size_t)b * kC * Tout;
  uint32_t tm = 0, phase = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
    cp_async_wait<0>();        // this tile's operand image (and W_eff on the first tile)
    fence_async_smem();
    tc_fence_before();
    __syncthreads();           // also: the previous tile's output tile has left `stage` (thread 0 waited)
    tc_fence_after();
    tm = tmem_slot;
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_16(kKP, false, false, F16 ? 0u : 1u);
      const uint32_t aa = smem_u32(a), ww = smem_u32(w);
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_bf16(tm, make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(ww + s * 2 * kCS, kCS, kRS144), idesc, s > 0);
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (tile + 1 < tile_end) {   // the UMMAs have read `a`: the next tile's image streams in under the epilogue
      cp_async_block(a, vb + (size_t)(tile + 1) * kTile144, kTile144, kApplyThreads);
      cp_async_commit();
    }
    // token-major store obuf[b][head][token][68] (16 bit) == the reference's scramble buffer.  A tile's
    // rows of one head are ONE contiguous block of nvalid * 136 bytes in global memory: the tile is
    // assembled in shared memory and leaves with two bulk (TMA) stores instead of 18 scattered 8-byte
    // stores per thread (each warp store touched 32 separate sectors).
    const bool bulk = ((Tout | nvalid) & 1) == 0;   // 16-byte size / alignment of the blocks
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
            if (bulk)
              *reinterpret_cast<uint2 *>(stage + (h * kTM + tok) * kHS + cp) = pk;
            else
              *reinterpret_cast<uint2 *>(ob + ((size_t)h * Tout + t0 + tok) * kHS + cp) = pk;
          }
        }
      }
    }
    tc_fence_before();         // the accumulator is rewritten by the next tile's UMMAs
    if (bulk) {
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)nvalid * kHS * 2;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ob + ((size_t)h * Tout + t0) * kHS),
                       "r"(smem_u32(stage + h * kTM * kHS)), "r"(bytes)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the shared tile has been read
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
// 16-bit operand element of format FMT: 0 = bfloat16, 1 = IEEE half, 2 = bfloat16 remainder bf16(v - bf16(v))
template <int FMT>
__device__ __forceinline__ unsigned short to_fmt(float v) {
  if constexpr (FMT == 1) return to16<true>(v);
  if constexpr (FMT == 0) return to16<false>(v);
  return to16<false>(v - __bfloat162float(__float2bfloat16_rn(v)));
}
// ALL weight pieces of a block in one launch: blockIdx.y = piece (Wp 0..2, W1 chunk c slice s 3 + 3c + s, W2 chunk
// c slice s 15 + 3c + s), blockIdx.z = set (0 bf16 hi, 1 fp16, 2 bf16 lo: W1 / W2 only).  A piece is the
// [144 rows][48 K] operand image of a K-slice; LayerNorm-2's scale is folded into W1's columns, and column 136
// (= column 40 of K-slice 2) of the Wp / W1 pieces carries the bias that rides in the MMA against the
// activation tile's ones column (b1f = b_1 + W_1 beta_2, computed by fold_bias_kernel before this launch).
struct PiecePackArgs {
  const float *proj_w, *proj_b, *mlp0_w, *mlp3_w, *ln2_w, *b1f;
  uint8_t *base;
  size_t img_bytes, lo;
};
__global__ void pack_pieces_kernel(const PiecePackArgs A) {
  const int idx = blockIdx.y, f = blockIdx.z;
  if (f == 2 && idx < kPiecesPerImage) return;   // no lo term for Wp
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kKP * kPieceK) return;
  const int r = e / kPieceK, k = e % kPieceK;
  const int fam = idx < kPiecesPerImage ? 0 : (idx < kPiecesPerImage * (1 + kNChunk) ? 1 : 2);
  const int c = fam == 0 ? 0 : (idx - kPiecesPerImage * (fam == 1 ? 1 : 1 + kNChunk)) / kPiecesPerImage;
  const int sl = idx % kPiecesPerImage;
  float v = 0.f;
  if (fam == 0) {          // Wp: rows = output channels, K = input channels
    const int col = sl * kPieceK + k;
    if (r < kC && col < kC) v = A.proj_w[(size_t)r * kC + col];
    else if (r < kC && col == kC) v = A.proj_b[r];
  } else if (fam == 1) {   // W1 chunk c: rows = hidden c*144 + r, K = input channels, LN2 scale folded in
    const int hrow = c * kNH + r, col = sl * kPieceK + k;
    if (hrow < 4 * kC && col < kC) v = A.mlp0_w[(size_t)hrow * kC + col] * A.ln2_w[col];
    else if (hrow < kHidPad && col == kC) v = A.b1f[hrow];
  } else {                 // W2 chunk c: rows = output channels, K = hidden c*144 + 48 s + k
    const int hcol = c * kNH + sl * kPieceK + k;
    if (r < kC && hcol < 4 * kC) v = A.mlp3_w[(size_t)r * 4 * kC + hcol];
  }
  uint8_t *dst = f < 2 ? A.base + f * A.img_bytes + (size_t)idx * kPiece
                       : A.base + A.lo + (size_t)(idx - kPiecesPerImage) * kPiece;
  const unsigned short h = f == 1 ? to_fmt<1>(v) : (f == 0 ? to_fmt<0>(v) : to_fmt<2>(v));
  *reinterpret_cast<unsigned short *>(dst + cm_offset(r, k, kRS48, 128)) = h;
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

constexpr size_t kApplySmem = (size_t)kTile144 + kW144 + kApplyStage;
static_assert(2 * (kApplySmem + 2048) <= 228 * 1024, "two tc_apply CTAs per SM");
constexpr size_t kBackSmem = (size_t)2 * kTile144 + (size_t)kBackHB * kTile144 + (size_t)kBackSlots * kPiece + sizeof(BackVec);
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
  const float *wqaT = reinterpret_cast<const float *>(base + L.wqaT);
  const float *wkaT = reinterpret_cast<const float *>(base + L.wkaT);
  const uint8_t *tcw_lo = base + L.lo;   // bf16 lo terms of the W1 / W2 chunk images
  uint8_t *ws = static_cast<uint8_t *>(ws_tc);
  float *gram = reinterpret_cast<float *>(ws + W.gram_part);
  float *spart = reinterpret_cast<float *>(ws + W.spart);
  float *beff = reinterpret_cast<float *>(ws + W.beff);
  uint8_t *weff = ws + W.weff;
  uint8_t *vn = ws + W.vn;
  unsigned short *obuf = reinterpret_cast<unsigned short *>(ws + W.obuf);
  static PerDeviceOnce attr;
  if (attr.first()) {
    if (!set_max_smem(tc_front1_kernel<F16, false>, kFront1Smem, "tc_front1_kernel") ||
        !set_max_smem(tc_front1_kernel<F16, true>, kFront1Smem, "tc_front1_kernel") ||
        !set_max_smem(tc_apply_kernel<F16>, kApplySmem, "tc_apply_kernel") ||
        !set_max_smem(gram_project_kernel, sizeof(GpSmem), "gram_project_kernel") ||
        !set_max_smem(tc_back_kernel<F16, false>, kBackSmem, "tc_back_kernel") ||
        !set_max_smem(tc_back_kernel<F16, true>, kBackSmem, "tc_back_kernel"))
      return OTP_ERR_CUDA;
  }
  {
    LaunchScope ls(K_TC_FRONT, st);
    // x as a 2-D tensor [(clip, channel)][token] read in [136][128] tiles by TMA (stride-1 blocks)
    alignas(64) CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    const int x_tma = (stride == 1 && make_tensor_map_2d_f32(&xmap, x, (uint64_t)b * kC, (uint64_t)t, kC, kTM)) ? 1 : 0;
    if (stride == 1)
      tc_front1_kernel<F16, false><<<dim3(W.nchunk, b), kFrThreads, kFront1Smem, st>>>(
          P, x, xmap, x_tma, gram, vn, t, W.tout, W.tiles, W.tiles_per_chunk, W.nchunk, g_trace_on);
    else
      tc_front1_kernel<F16, true><<<dim3(W.nchunk, b), kFrThreads, kFront1Smem, st>>>(
          P, x, xmap, 0, gram, vn, t, W.tout, W.tiles, W.tiles_per_chunk, W.nchunk, g_trace_on);
  }
  {
    // S = Wq~ G~ Wk~^T in fp32 (8 column blocks per clip), then softmax + W_eff fold
    LaunchScope ls(K_BLOCK_FOLD, st, 2);
    gram_project_kernel<<<dim3(kGpBlocks, b), kGpThreads, sizeof(GpSmem), st>>>(gram, W.nchunk, wqaT, wkaT, spart);
    block_fold_kernel<kC, F16 ? 2 : 1><<<dim3(b, FoldCfg<kC>::NBLK), kFoldThreads, 0, st>>>(wvp, bvp, spart, kGramBlocks,
                                                                                           weff, beff, kKP, kKP);
  }
  {
    LaunchScope ls(K_TC_APPLY, st);
    // runs of consecutive tiles per CTA: ~2 CTAs per SM over the whole batch, at least one tile each
    const int want = max(1, ceil_div(2 * num_sms(), b));
    const int tpc = max(1, ceil_div(W.tiles, min(W.tiles, want)));
    tc_apply_kernel<F16><<<dim3(ceil_div(W.tiles, tpc), b), kApplyThreads, kApplySmem, st>>>(vn, weff, beff, obuf, W.tout,
                                                                                            W.tiles, tpc);
  }
  {
    LaunchScope ls(K_TC_BACK, st);
    const int total = b * W.tiles;
    const int grid = min(total, num_sms());
    // y as a 2-D tensor [(clip, channel)][token] written in [136][128] tiles by TMA
    alignas(64) CUtensorMap ymap;
    memset(&ymap, 0, sizeof(ymap));
    const int use_tma = make_tensor_map_2d_f32(&ymap, y, (uint64_t)b * kC, (uint64_t)W.tout, kC, kTM) ? 1 : 0;
    if (stride == 1)
      tc_back_kernel<F16, false><<<grid, kBackThreads, kBackSmem, st>>>(P, tcw, tcw_lo, x, obuf, y, ymap, use_tma, b, t,
                                                                        W.tout, W.tiles, g_trace_on);
    else
      tc_back_kernel<F16, true><<<grid, kBackThreads, kBackSmem, st>>>(P, tcw, tcw_lo, x, obuf, y, ymap, use_tma, b, t,
                                                                       W.tout, W.tiles, g_trace_on);
  }
  return check_launch("block_forward_tc");
}
}  // namespace

bool block_tc_built() { return true; }

void block_tc_trace(int on) { g_trace_on = on; }
int block_tc_trace_read(unsigned long long *out, int n) {
  if (n < 3 * kTraceLen) return OTP_ERR_ARG;
  const int rows = n >= 4 * kTraceLen ? 4 : 3;   // row 3 (tc_front) only for callers that ask for it
  cudaError_t e = cudaMemcpyFromSymbol(out, g_back_trace, sizeof(unsigned long long) * rows * kTraceLen);
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
  LaunchScope ls(K_PACK, st, 8);
  scale_cols_kernel<<<ceil_div(kC * kC, 256), 256, 0, st>>>(p->v_w, p->v_norm_w, reinterpret_cast<float *>(base + L.wvp),
                                                            kC, kC);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->v_w, p->v_b, p->v_norm_b, reinterpret_cast<float *>(base + L.bvp), kC, kC, kKP);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->q_w, p->q_b, p->q_norm_b, reinterpret_cast<float *>(base + L.bqp), kC, kC, kKP);
  fold_bias_kernel<<<1, 256, 0, st>>>(p->k_w, p->k_b, p->k_norm_b, reinterpret_cast<float *>(base + L.bkp), kC, kC, kKP);
  fold_bias_kernel<<<ceil_div(kHidPad, 256), 256, 0, st>>>(p->mlp0_w, p->mlp0_b, p->ln2_b,
                                                           reinterpret_cast<float *>(base + L.b1p), 4 * kC, kC, kHidPad);
  // augmented, transposed, LayerNorm-folded q / k projections in fp32 (1/sqrt(hs) folded into Wq~)
  pack_aug_T_kernel<<<ceil_div(kKP * kKP, 256), 256, 0, st>>>(p->q_w, p->q_norm_w, reinterpret_cast<const float *>(base + L.bqp),
                                                              1.0f / sqrtf((float)kHS), reinterpret_cast<float *>(base + L.wqaT));
  pack_aug_T_kernel<<<ceil_div(kKP * kKP, 256), 256, 0, st>>>(p->k_w, p->k_norm_w, reinterpret_cast<const float *>(base + L.bkp),
                                                              1.0f, reinterpret_cast<float *>(base + L.wkaT));
  // all 27 weight pieces x {bf16 hi, fp16, bf16 lo} in one launch (b1f is ready: same stream)
  const PiecePackArgs A{p->proj_w, p->proj_b, p->mlp0_w, p->mlp3_w, p->ln2_w, reinterpret_cast<const float *>(base + L.b1p),
                        base, L.img_bytes, L.lo};
  pack_pieces_kernel<<<dim3(ceil_div(kKP * kPieceK, 256), kNumPieces, 3), 256, 0, st>>>(A);
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
