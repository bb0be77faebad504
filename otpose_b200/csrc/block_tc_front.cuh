// tc_front: LN1 -> {dw_q, dw_k, dw_v} -> {LN_q, LN_k, LN_v} -> per-CTA partial AUGMENTED channel Gram of
// one TransformerBlock (model/blocks.py:264-268, 400-440), stride-1 stem blocks and (template S2)
// stride-2 branch blocks.  Included by block_tc.cu inside its anonymous namespace.
//
// "Gram first": the reference forms q = W_q a + b_q, k = W_k c + b_k per token (a = LN_q(dw_q(LN1 x)),
// c = LN_k(dw_k(LN1 x))) and reduces S_h = sum_t q_h k_h^T over ALL tokens of the clip.  Both maps are
// linear in the token, so with a~ = [a; 1], c~ = [c; 1]
//
//     S_h = [W_q | b_q]_h  ( sum_t a~_t c~_t^T )  [W_k | b_k]_h^T  =  Wq~_h  G~  Wk~_h^T ,
//
// and only the 137 x 137 Gram G~ of the NORMALISED depthwise outputs has to be accumulated over the
// tokens; the two projections are applied once per clip, in fp32, by the fold kernels (block_fold.cuh,
// gram_project_kernel below).  This removes two of the three per-token GEMMs of this pass, their
// TMEM -> bias -> 16-bit write-back epilogues and both resident weight images, and -- the reason it was
// done -- W_q / W_k are never rounded to 16 bits: a rounded weight is the same perturbation for every
// token, so its error adds up coherently over the 6912-token reduction and is then amplified by the
// softmax (scripts/emulate_operand_rounding.py: 1.1e-2 of the 1.3e-2 bfloat16 feature error), whereas
// the rounding of a_t / c_t is independent per token and averages out.
//
// One warp-specialised CTA per (clip, token chunk), 448 threads:
//   warps 0-11  COMPUTE: thread (q4, lane, third) owns token 32*q4 + lane and 48 channels.  The
//               token's x values are loaded straight into registers one tile ahead (no fp32 staging
//               tile); LN1(x) is exchanged between neighbouring tokens through a 16-bit (IEEE half)
//               shared tile.  a_t / c_t go into 16-bit [token][channel] operand tiles (channel 136 = the
//               ones column), double buffered; vn (the v branch) goes to global memory for tc_apply.
//   warp 12     MMA ISSUER (converged warp, elect-predicated tcgen05.mma): G~ += a~^T c~ over the tile's
//               128 tokens -- MN-major views of the two tiles, two M = 128 row blocks (channels 0..127 and
//               16..143), N = 144, accumulated in TMEM over the CTA's whole token chunk.  Runs under the
//               next tile's LayerNorm / depthwise work.
//   warp 13     HALO: LN1 of the two tokens next to the tile (t0-1, t0+128), one tile ahead.
constexpr int kFrComp = 384;
constexpr int kFrThreads = kFrComp + 64;
constexpr uint32_t kHsRow = kC * 2;           // 272-byte row of 136 halves: 16-byte accesses of
                                              // consecutive lanes fall into distinct bank groups
constexpr int kGramLd = kKP;                  // row stride (floats) of a partial G~ in global memory
constexpr int kGramRows = kC + 1;             // 137: channels + the ones row
struct Front1Vec {
  float ln1w[kC], ln1b[kC];
  float4 dw[3][kC];          // depthwise taps of q, k, v (fp32: the stride-2 blocks, whose centre tap is fp32)
  uint4 dwh[3][kKP / 8][3];  // the same taps as IEEE-half pairs, [branch][8-channel group][tap]: the stride-1
                             // blocks run the depthwise conv in packed half2 on the half LN1 tile
  float part[2][4][3][kTM];  // [parity of the pass][mean | M2 or sum | sumsq (| odd input: mean | M2)][third][token]
};
struct Front1Bars {
  uint64_t abfull;           // compute -> MMA: a~ and c~ tiles staged (12 warp arrivals)
  uint64_t gdone;            // Gram UMMAs of the tile done: a~ / c~ tiles reusable
  uint64_t xfull;            // TMA arrival of the fp32 x tile (stride-1 blocks with a 16-byte row pitch)
  uint64_t halo_full[2];     // halo rows of tile n staged in halo[n & 1]
  uint64_t adone;            // compute warps are done reading the LN1 tile / halo rows (12 warp arrivals)
};
constexpr uint32_t kXTile = kC * kTM * 4;     // 69632: fp32 [136 channels][128 tokens] x tile (TMA box)
constexpr size_t kFront1Smem = (size_t)kXTile + 2 * kTile144 + (size_t)kTM * kHsRow + 4 * kHsRow + sizeof(Front1Vec);
static_assert(kFront1Smem + 1024 <= 227 * 1024, "tc_front1 shared memory");

__device__ __forceinline__ void fr_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFrComp) : "memory"); }
__device__ __forceinline__ void unpack8(const uint4 &p, float (&f)[8]) {
  const __half2 *h = reinterpret_cast<const __half2 *>(&p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// S2 = the stride-2 branch blocks (depthwise convs with stride 2: output token j reads inputs 2j-1,
// 2j, 2j+1).  Thread = OUTPUT token: LN1 of its even input stays in registers (fp32), LN1 of its odd
// input goes to the shared tile (row j = input 2j+1: its own right tap and token j+1's left tap), so
// the tile is the same size as for stride 1; only a left halo row (input 2*t0-1) is needed.  The two
// input tokens are loaded at the start of the tile (no register room to prefetch them a tile ahead).
// 14 warps: the busiest scheduler partition holds 4 of them, i.e. 16384 / 4 / 32 = 128 registers per thread
template <bool F16, bool S2>
__global__ void __launch_bounds__(kFrThreads, 1)
tc_front1_kernel(BlockPack P, const float *__restrict__ x, const __grid_constant__ CUtensorMap xmap, int use_tma,
                 float *__restrict__ gram_part, uint8_t *__restrict__ vn_img, int T, int Tout, int tiles,
                 int tiles_per_chunk, int nchunk, int trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float *xs = reinterpret_cast<float *>(smem); // fp32 x tile [136][128], filled by one tensor-map TMA load per tile
  uint8_t *ab = smem + kXTile;                 // [a~ tile | c~ tile], [token][channel] 16-bit
  uint8_t *hs = ab + 2 * kTile144;             // [128][136] halves: LN1(x) of the tile's tokens
  uint8_t *halo = hs + kTM * kHsRow;           // [2][2][136] halves: LN1(x) of tokens t0-1 / t0+128
  Front1Vec *V = reinterpret_cast<Front1Vec *>(halo + 4 * kHsRow);
  __shared__ Front1Bars bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tile_begin = chunk * tiles_per_chunk;
  const int tile_end = min(tiles, tile_begin + tiles_per_chunk);

  for (int c = threadIdx.x; c < kC; c += kFrThreads) {
    V->ln1w[c] = P.ln1_w[c];
    V->ln1b[c] = P.ln1_b[c];
    V->dw[0][c] = make_float4(P.dwq[3 * c], P.dwq[3 * c + 1], P.dwq[3 * c + 2], 0.f);
    V->dw[1][c] = make_float4(P.dwk[3 * c], P.dwk[3 * c + 1], P.dwk[3 * c + 2], 0.f);
    V->dw[2][c] = make_float4(P.dwv[3 * c], P.dwv[3 * c + 1], P.dwv[3 * c + 2], 0.f);
  }
  for (int e = threadIdx.x; e < 3 * kKP * 3; e += kFrThreads) {   // (branch, channel, tap) -> half
    const int m = e / (kKP * 3), c = (e / 3) % kKP, tap = e % 3;
    const float *src = m == 0 ? P.dwq : (m == 1 ? P.dwk : P.dwv);
    reinterpret_cast<__half *>(&V->dwh[m][c >> 3][tap])[c & 7] = __float2half_rn(c < kC ? src[3 * c + tap] : 0.f);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars.abfull, kFrComp / 32);
    mbar_init(&bars.gdone, 1);
    mbar_init(&bars.xfull, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) mbar_init(&bars.halo_full[i], 1);
    mbar_init(&bars.adone, kFrComp / 32);
    fence_mbar_init();
  }
  if (warp == kFrComp / 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_g0 = tm, t_g1 = tm + kKP;   // rows = a~ channels 0..127 / 16..143, columns = c~ channels
  constexpr uint32_t kFmt = F16 ? 0u : 1u;
  const float *xb = x + (size_t)b * kC * T;

  if (warp == kFrComp / 32) {
    // =============================================================== MMA ISSUER
    const uint32_t idesc_gram = make_idesc_16(kKP, true, true, kFmt);
    const uint32_t ab0 = smem_u32(ab);
    uint32_t n = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      const uint32_t a_s = ab0, c_s = a_s + kTile144;
      mbar_wait(&bars.abfull, n & 1);
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kTM / 16; ++s) {
        const uint32_t ko = s * 2 * kRS144;   // 16 tokens = two 8-row groups
        umma_elect(t_g0, make_desc(a_s + ko, kRS144, kCS), make_desc(c_s + ko, kRS144, kCS), idesc_gram,
                   !(n == 0 && s == 0));
        umma_elect(t_g1, make_desc(a_s + ko + 2 * kCS, kRS144, kCS), make_desc(c_s + ko, kRS144, kCS), idesc_gram,
                   !(n == 0 && s == 0));
      }
      commit_elect(&bars.gdone);
    }
  } else if (warp == kFrComp / 32 + 1) {
    // =============================================================== HALO
    uint32_t n = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      if (n >= 2) mbar_wait(&bars.adone, n & 1);   // tile n-2 no longer reads halo[n & 1]
      const int t0 = tile * kTM;
#pragma unroll 1
      for (int side = 0; side < (S2 ? 1 : 2); ++side) {
        const int t = S2 ? 2 * t0 - 1 : (side ? t0 + kTM : t0 - 1);
        const bool ok = t >= 0 && t < T;
        float v[5];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int c = lane + 32 * k;
          v[k] = (ok && c < kC) ? __ldg(xb + (size_t)c * T + t) : 0.f;
          s += v[k];
        }
        const float mu = warp_sum(s) * (1.0f / kC);
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float d = (lane + 32 * k < kC) ? v[k] - mu : 0.f;
          ss = fmaf(d, d, ss);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(ss) * (1.0f / kC) + 1e-5f);
        __half *row = reinterpret_cast<__half *>(halo + ((n & 1) * 2 + side) * kHsRow);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int c = lane + 32 * k;
          if (c < kC) row[c] = __float2half_rn(ok ? fmaf((v[k] - mu) * rstd, V->ln1w[c], V->ln1b[c]) : 0.f);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.halo_full[n & 1]);
    }
  } else {
    // =============================================================== COMPUTE
    const int q4 = warp & 3, third = warp >> 2;
    const int tok = q4 * 32 + lane;
    const int c_lo = third * 48;
    const int nq = min(48, kC - c_lo);   // valid channels of this third (48, 48, 40)
    float xr[48];              // stride 1: the token's x, one tile ahead; stride 2: the even input 2j
    float xo[S2 ? 48 : 1];     // stride 2: the odd input 2j+1
    auto load_x = [&](int tile) {
      const int t = S2 ? 2 * (tile * kTM + tok) : tile * kTM + tok;
      const float *p = xb + (size_t)c_lo * T + t;
#pragma unroll
      for (int i = 0; i < 48; ++i) {
        xr[i] = (t < T && i < nq) ? __ldg(p) : 0.f;
        if (S2) xo[i] = (t + 1 < T && i < nq) ? __ldg(p + 1) : 0.f;
        p += T;
      }
    };
    // stride-1 blocks whose rows are 16-byte aligned: the tile's x arrives as ONE tensor-map TMA load
    // (cp.async.bulk.tensor.2d, UTMALDG) into shared memory while the previous tile is being processed, and
    // every thread picks its 48 values up from there -- 48 strided global loads per thread kept the compute
    // warps blocked for ~3 k cycles per tile while the SM's load queue drained
    const bool tma_x = !S2 && use_tma != 0;
    auto issue_x_tma = [&](int tile) {   // one thread
      mbar_expect_tx(&bars.xfull, kXTile);
      tma_load_2d(xs, &xmap, tile * kTM, b * kC, &bars.xfull);
    };
    auto take_x = [&](uint32_t nn) {     // xs -> registers
      mbar_wait(&bars.xfull, nn & 1);
      const float *p = xs + c_lo * kTM + tok;
#pragma unroll
      for (int i = 0; i < 48; ++i) xr[i] = i < nq ? p[i * kTM] : 0.f;
    };
    if (tile_begin < tile_end) {
      if (tma_x) {
        if (threadIdx.x == 0) issue_x_tma(tile_begin);
      } else if (!S2) {
        load_x(tile_begin);
      }
    }
    uint32_t n = 0, pp = 0;
    const uint8_t *hrow = hs + tok * kHsRow + c_lo * 2;
    // optional phase trace (build with -DOTP_FRONT_TRACE, then otp_debug_trace / scripts/trace_front.py):
    // compute warp 0 of CTA (0, 0) -> row 3 of the trace buffer.  Compiled out by default: the kernel sits
    // at its 128-register ceiling and the tracer's three registers turn into spills.
#ifdef OTP_FRONT_TRACE
    Tracer tr{(trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ? g_back_trace[3] : nullptr, 0};
#else
    (void)trace;
    auto tr = [](int) {};
#endif
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      const int t0 = tile * kTM;
      const int nvalid = min(kTM, Tout - t0);
      const bool live = tok < nvalid;                             // output token (and its even input) exists
      const bool live_o = S2 && 2 * (t0 + tok) + 1 < T;           // its odd input exists
      uint8_t *vn_tile = vn_img + ((size_t)b * tiles + tile) * kTile144;
      uint8_t *a_tile = ab, *c_tile = a_tile + kTile144;
      tr(0);
      if (S2) load_x(tile);
      if (tma_x) take_x(n);
      // ---- LN1 over the token's 136 channels: per-thread (mean, M2), parallel-variance combine ----
      {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 48; ++i) s += xr[i];
        const float mq = s / (float)nq;
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < 48; ++i) {
          const float d = i < nq ? xr[i] - mq : 0.f;
          m2 = fmaf(d, d, m2);
        }
        V->part[pp][0][third][tok] = mq;
        V->part[pp][1][third][tok] = m2;
        if (S2) {
          float so = 0.f;
#pragma unroll
          for (int i = 0; i < 48; ++i) so += xo[i];
          const float mo = so / (float)nq;
          float m2o = 0.f;
#pragma unroll
          for (int i = 0; i < 48; ++i) {
            const float d = i < nq ? xo[i] - mo : 0.f;
            m2o = fmaf(d, d, m2o);
          }
          V->part[pp][2][third][tok] = mo;
          V->part[pp][3][third][tok] = m2o;
        }
      }
      tr(1);
      fr_bar_sync();
      tr(2);
      if (tma_x && threadIdx.x == 0 && tile + 1 < tile_end) {   // every thread has its x in registers: refill
        fence_async_smem();
        issue_x_tma(tile + 1);
      }
      if (S2) {   // odd input: LN1 -> 16-bit shared row; the even input is normalised in place below (xr := h)
        const float m0 = V->part[pp][2][0][tok], m1 = V->part[pp][2][1][tok], m2 = V->part[pp][2][2][tok];
        const float mu = (48.f * (m0 + m1) + 40.f * m2) * (1.0f / kC);
        const float d0 = m0 - mu, d1 = m1 - mu, d2 = m2 - mu;
        const float var = (V->part[pp][3][0][tok] + V->part[pp][3][1][tok] + V->part[pp][3][2][tok] +
                           48.f * (d0 * d0 + d1 * d1) + 40.f * d2 * d2) * (1.0f / kC);
        const float rstd = rsqrtf(var + 1e-5f);
        uint8_t *dst = hs + tok * kHsRow + c_lo * 2;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < nq) {
            __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c_lo + g * 8 + 2 * e;
              const float a0 = live_o ? fmaf((xo[g * 8 + 2 * e] - mu) * rstd, V->ln1w[c], V->ln1b[c]) : 0.f;
              const float a1 = live_o ? fmaf((xo[g * 8 + 2 * e + 1] - mu) * rstd, V->ln1w[c + 1], V->ln1b[c + 1]) : 0.f;
              h2[e] = __floats2half2_rn(a0, a1);
            }
            *reinterpret_cast<uint4 *>(dst + g * 16) = *reinterpret_cast<const uint4 *>(h2);
          }
        }
      }
      {
        const float m0 = V->part[pp][0][0][tok], m1 = V->part[pp][0][1][tok], m2 = V->part[pp][0][2][tok];
        const float mu = (48.f * (m0 + m1) + 40.f * m2) * (1.0f / kC);
        const float d0 = m0 - mu, d1 = m1 - mu, d2 = m2 - mu;
        const float var = (V->part[pp][1][0][tok] + V->part[pp][1][1][tok] + V->part[pp][1][2][tok] +
                           48.f * (d0 * d0 + d1 * d1) + 40.f * d2 * d2) * (1.0f / kC);
        const float rstd = rsqrtf(var + 1e-5f);
        pp ^= 1;
        uint8_t *dst = hs + tok * kHsRow + c_lo * 2;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < nq) {
            __half2 h2[4];
            // LN1 affine of these 8 channels as four 16-byte broadcasts (not 16 scalar loads)
            const float4 *w4 = reinterpret_cast<const float4 *>(V->ln1w + c_lo + g * 8);
            const float4 *b4 = reinterpret_cast<const float4 *>(V->ln1b + c_lo + g * 8);
            const float4 wa = w4[0], wb = w4[1], ba = b4[0], bb = b4[1];
            const float lw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const float lb[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a0 = live ? fmaf((xr[g * 8 + 2 * e] - mu) * rstd, lw[2 * e], lb[2 * e]) : 0.f;
              const float a1 = live ? fmaf((xr[g * 8 + 2 * e + 1] - mu) * rstd, lw[2 * e + 1], lb[2 * e + 1]) : 0.f;
              if (S2) {   // centre tap stays in registers (fp32)
                xr[g * 8 + 2 * e] = a0;
                xr[g * 8 + 2 * e + 1] = a1;
              } else {
                h2[e] = __floats2half2_rn(a0, a1);   // zero == the conv's zero padding past the sequence end
              }
            }
            if (!S2) *reinterpret_cast<uint4 *>(dst + g * 16) = *reinterpret_cast<const uint4 *>(h2);
          }
        }
      }
      tr(3);
      fr_bar_sync();
      tr(4);
      mbar_wait(&bars.halo_full[n & 1], (n >> 1) & 1);
      tr(5);
      const uint8_t *lrow = tok == 0 ? halo + ((n & 1) * 2 + 0) * kHsRow + c_lo * 2 : hrow - kHsRow;
      const uint8_t *rrow = S2 ? hrow   // stride 2: the right tap is this token's own odd input
                               : (tok == kTM - 1 ? halo + ((n & 1) * 2 + 1) * kHsRow + c_lo * 2 : hrow + kHsRow);
      // ---- q, k, v in turn: depthwise conv (registers) -> statistics -> (d - mean) * rstd -> operand tile ----
      auto pass = [&](auto mc) {
        constexpr int m = decltype(mc)::value;
        float d[48];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < nq) {
            if constexpr (!S2) {
              // stride 1: all three taps are rows of the half LN1 tile -> the conv runs in packed half2 (3
              // HFMA2-class instructions per channel PAIR instead of 6 conversions + 6 fp32 FMAs); the
              // statistics and the normalisation stay fp32
              const uint4 lq = *reinterpret_cast<const uint4 *>(lrow + g * 16);
              const uint4 cq4 = *reinterpret_cast<const uint4 *>(hrow + g * 16);
              const uint4 rq = *reinterpret_cast<const uint4 *>(rrow + g * 16);
              const uint4 *wt = V->dwh[m][(c_lo >> 3) + g];
              const uint4 wl = wt[0], wc = wt[1], wr = wt[2];
              const __half2 *l2 = reinterpret_cast<const __half2 *>(&lq), *c2 = reinterpret_cast<const __half2 *>(&cq4),
                            *r2 = reinterpret_cast<const __half2 *>(&rq), *wl2 = reinterpret_cast<const __half2 *>(&wl),
                            *wc2 = reinterpret_cast<const __half2 *>(&wc), *wr2 = reinterpret_cast<const __half2 *>(&wr);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2 hv = __hfma2(wr2[q], r2[q], __hfma2(wc2[q], c2[q], __hmul2(wl2[q], l2[q])));
                const float2 f = __half22float2(hv);
                d[g * 8 + 2 * q] = f.x;
                d[g * 8 + 2 * q + 1] = f.y;
                s += f.x + f.y;
                ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
              }
            } else {
              float l8[8], c8[8], r8[8];
              unpack8(*reinterpret_cast<const uint4 *>(lrow + g * 16), l8);
#pragma unroll
              for (int e = 0; e < 8; ++e) c8[e] = xr[g * 8 + e];
              unpack8(*reinterpret_cast<const uint4 *>(rrow + g * 16), r8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float4 w = V->dw[m][c_lo + g * 8 + e];
                const float v = fmaf(w.z, r8[e], fmaf(w.y, c8[e], w.x * l8[e]));
                d[g * 8 + e] = v;
                s += v;
                ss = fmaf(v, v, ss);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) d[g * 8 + e] = 0.f;
          }
        }
        V->part[pp][0][third][tok] = s;
        V->part[pp][1][third][tok] = ss;
        tr(10 + m);
        fr_bar_sync();
        tr(20 + m);
        const float mean = (V->part[pp][0][0][tok] + V->part[pp][0][1][tok] + V->part[pp][0][2][tok]) * (1.0f / kC);
        const float var = fmaxf((V->part[pp][1][0][tok] + V->part[pp][1][1][tok] + V->part[pp][1][2][tok]) * (1.0f / kC) -
                                    mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;   // (d - mean) * rstd as one FMA
        pp ^= 1;
        // the Gram UMMAs of the previous tile still read the a~ / c~ tiles (they were issued a whole LN1 phase
        // ago): wait before overwriting them
        if (m == 0 && n >= 1) mbar_wait(&bars.gdone, (n - 1) & 1);
        tr(30 + m);
        uint8_t *dst = (m == 0 ? a_tile : (m == 1 ? c_tile : vn_tile)) + cm_offset(tok, c_lo, kRS144, kCS);
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)   // rows of tokens past the sequence end are zero: they must not reach G~
            o8[e] = (g * 8 + e < nq && (m == 2 || live)) ? fmaf(d[g * 8 + e], rstd, nmr) : 0.f;
          // channel 136 of a~ / c~ = 1 for the tile's real tokens: row / column 136 of G~ carry sum_t c_t,
          // sum_t a_t and the token count, which is how b_q / b_k enter S (gram_project_kernel)
          if (m < 2 && g == 5 && third == 2) o8[0] = live ? 1.f : 0.f;
          *reinterpret_cast<uint4 *>(dst + g * kCS) = pack16x8<F16>(o8);
        }
        tr(40 + m);
      };
      pass(std::integral_constant<int, 0>{});
      pass(std::integral_constant<int, 1>{});
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.abfull);
      // next tile's x: issued behind the tile's only proxy fence (fence.proxy.async waits for the thread's
      // pending loads), in flight during the v pass
      if (!S2 && !tma_x && tile + 1 < tile_end) load_x(tile + 1);
      pass(std::integral_constant<int, 2>{});
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.adone);
    }
    // ---- flush the partial G~: TMEM lane == a~ channel (row), column == c~ channel ----
    if (n > 0) {
      mbar_wait(&bars.gdone, (n - 1) & 1);   // commits complete in order: every tile is in
      tc_fence_after();
      float *gp = gram_part + (size_t)(b * nchunk + chunk) * kKP * kGramLd;
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        if (blk == 1 && q4 != 3) break;                       // row block 1 only contributes channels 128..136
        const int row_ch = blk ? 16 + tok : tok;
        const bool row_ok = blk ? (row_ch >= kTM && row_ch < kGramRows) : true;
#pragma unroll 1
        for (int g = 0; g < 6; ++g) {
          float v[8];
          tmem_ld8(tcol(blk ? t_g1 : t_g0, q4, c_lo + g * 8), v);
          if (row_ok) {
            float4 *dst = reinterpret_cast<float4 *>(gp + (size_t)row_ch * kGramLd + c_lo + g * 8);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFrComp / 32) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------ gram_project (fold, part 1)
// S = Wq~ G~ Wk~^T per head, split over 8 column blocks J of G~ so that 8 * B CTAs share the work and every
// CTA reduces only its own 18 columns of the per-chunk partial Grams (fixed order: deterministic):
//     G_J   = sum_chunks part[chunk][:, J]                         (137 x 18)
//     M1_J  = Wq~ G_J                                              (136 x 18)
//     S^(J)[i][m] = sum_{j in J} M1_J[i][j] Wk~[h(i)*68 + m][j]    (136 x 68 partial of S)
// The 8 partials are summed (again in fixed order) by block_fold_kernel, which also does the softmax and
// the W_eff fold.  wqaT / wkaT: fp32 [144][144] TRANSPOSED augmented weights ([r][i] = Wq~[i][r], LayerNorm
// affine and 1/sqrt(hs) folded in, row 136 = the folded bias), so consecutive threads read consecutive
// addresses.
constexpr int kGpCols = 18, kGpBlocks = kKP / kGpCols, kGpThreads = 2 * kKP;
static_assert(kGpBlocks * kGpCols == kKP && kGpCols % 2 == 0 && kHS % 2 == 0, "column blocks tile the padded Gram");
static_assert(kGpThreads == 2 * kKP, "gram_project: two threads per padded channel (the barrier below is CTA-wide)");
constexpr int kGpElems = (kGramRows * (kGpCols / 2) + kGpThreads - 1) / kGpThreads;   // float2 elements per thread (5)
constexpr int kGpLd = 20;           // padded row of 18 floats: 80 bytes, float4-aligned broadcast reads
struct GpSmem {
  float wq[kGramRows][kKP];        // Wq~^T rows 0..136 (cp.async, lands under the chunk reduction)
  float wk[kGpCols][kKP];          // Wk~^T rows j0 .. j0+17
  float G[kGramRows][kGpLd];
  float M1[kC][kGpLd];             // [q channel][column]
};
// 18 FMAs against one padded row read as five float4 broadcasts
__device__ __forceinline__ void gp_fma_row(float (&acc)[kGpCols], float w, const float *row) {
  const float4 *r4 = reinterpret_cast<const float4 *>(row);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 g = r4[q];
    acc[4 * q] = fmaf(w, g.x, acc[4 * q]);
    acc[4 * q + 1] = fmaf(w, g.y, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(w, g.z, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(w, g.w, acc[4 * q + 3]);
  }
  const float2 g = *reinterpret_cast<const float2 *>(row + 16);
  acc[16] = fmaf(w, g.x, acc[16]);
  acc[17] = fmaf(w, g.y, acc[17]);
}
__global__ void __launch_bounds__(kGpThreads)
gram_project_kernel(const float *__restrict__ gram_part, int nchunk, const float *__restrict__ wqaT,
                    const float *__restrict__ wkaT, float *__restrict__ spart) {
  extern __shared__ __align__(16) uint8_t gp_smem[];
  GpSmem &sm = *reinterpret_cast<GpSmem *>(gp_smem);
  const int b = blockIdx.y, J = blockIdx.x, j0 = J * kGpCols;
  // weights: 16-byte async copies (both arrays are contiguous row ranges of the packed matrices)
  for (int o = threadIdx.x; o < kGramRows * kKP / 4; o += kGpThreads)
    cp_async16(&sm.wq[0][0] + 4 * o, wqaT + 4 * o);
  for (int o = threadIdx.x; o < kGpCols * kKP / 4; o += kGpThreads)
    cp_async16(&sm.wk[0][0] + 4 * o, wkaT + (size_t)j0 * kKP + 4 * o);
  cp_async_commit();
  // G_J = fixed-order sum of the chunk partials; every load of a chunk is issued before the first add
  const float *gp = gram_part + (size_t)b * nchunk * kKP * kGramLd + j0;
  {
    float2 acc[kGpElems];
    size_t off[kGpElems];
#pragma unroll
    for (int k = 0; k < kGpElems; ++k) {
      const int e = min(threadIdx.x + k * kGpThreads, kGramRows * (kGpCols / 2) - 1);
      off[k] = (size_t)(e / (kGpCols / 2)) * kGramLd + 2 * (e % (kGpCols / 2));
      acc[k] = make_float2(0.f, 0.f);
    }
#pragma unroll 5   // 25 independent 8-byte loads in flight per thread: the phase is bound by memory latency
    for (int ch = 0; ch < nchunk; ++ch) {
      float2 v[kGpElems];
#pragma unroll
      for (int k = 0; k < kGpElems; ++k)
        v[k] = __ldg(reinterpret_cast<const float2 *>(gp + (size_t)ch * kKP * kGramLd + off[k]));
#pragma unroll
      for (int k = 0; k < kGpElems; ++k) {
        acc[k].x += v[k].x;
        acc[k].y += v[k].y;
      }
    }
#pragma unroll
    for (int k = 0; k < kGpElems; ++k) {
      const int e = threadIdx.x + k * kGpThreads;
      if (e < kGramRows * (kGpCols / 2))
        *reinterpret_cast<float2 *>(&sm.G[e / (kGpCols / 2)][2 * (e % (kGpCols / 2))]) = acc[k];
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // M1 = Wq~ G_J: thread (half, q channel) sums the rows r = half, half + 2, ... (shared memory is read as
  // float4 broadcasts: one LDS instruction per 4 FMAs, the LSU issue rate is what bounds this kernel)
  const int half = threadIdx.x / kKP, ch = threadIdx.x % kKP;
  const bool active = ch < kC;
  float acc[kGpCols];
#pragma unroll
  for (int jj = 0; jj < kGpCols; ++jj) acc[jj] = 0.f;
  if (active) {
#pragma unroll 4
    for (int r = half; r < kGramRows; r += 2) gp_fma_row(acc, sm.wq[r][ch], sm.G[r]);
  }
  // the two halves of the r range are added in a fixed order: half 0 stores, half 1 accumulates
  // (the barriers are outside every branch: all 288 threads reach them)
  float4 *dst = reinterpret_cast<float4 *>(sm.M1[active ? ch : 0]);
  if (active && half == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    dst[4] = make_float4(acc[16], acc[17], 0.f, 0.f);
  }
  __syncthreads();
  if (active && half == 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 o = dst[q];
      dst[q] = make_float4(o.x + acc[4 * q], o.y + acc[4 * q + 1], o.z + acc[4 * q + 2], o.w + acc[4 * q + 3]);
    }
    const float4 o = dst[4];
    dst[4] = make_float4(o.x + acc[16], o.y + acc[17], 0.f, 0.f);
  }
  __syncthreads();
  // S^(J)[i][m] = sum_j M1[i][j] Wk~[h*68 + m][j]: thread = (k channel m', half of that head's 68 rows)
  if (active) {
    const int h = ch / kHS, m = ch % kHS;
    float wk[kGpCols];
#pragma unroll
    for (int jj = 0; jj < kGpCols; ++jj) wk[jj] = sm.wk[jj][ch];
    float *sp = spart + ((size_t)(b * kGpBlocks + J) * kC + h * kHS) * kHS + m;
#pragma unroll 2
    for (int ii = half * (kHS / 2); ii < (half + 1) * (kHS / 2); ++ii) {
      const float4 *mr = reinterpret_cast<const float4 *>(sm.M1[h * kHS + ii]);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = mr[q];
        s = fmaf(v.x, wk[4 * q], s);
        s = fmaf(v.y, wk[4 * q + 1], s);
        s = fmaf(v.z, wk[4 * q + 2], s);
        s = fmaf(v.w, wk[4 * q + 3], s);
      }
      const float4 v = mr[4];
      s = fmaf(v.x, wk[16], s);
      s = fmaf(v.y, wk[17], s);
      sp[(size_t)ii * kHS] = s;
    }
  }
}

// fp32 [144][144] transposed augmented weight: out[r][i] = w[i][r] * g[r] * scale (r, i < 136),
// out[136][i] = bfold[i] * scale, zero elsewhere
__global__ void pack_aug_T_kernel(const float *__restrict__ w, const float *__restrict__ g,
                                  const float *__restrict__ bfold, float scale, float *__restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kKP * kKP) return;
  const int r = e / kKP, i = e % kKP;
  float v = 0.f;
  if (i < kC) {
    if (r < kC) v = w[(size_t)i * kC + r] * g[r] * scale;
    else if (r == kC) v = bfold[i] * scale;
  }
  out[e] = v;
}
