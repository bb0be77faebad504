// tc_front: LN1 -> {dw_q, dw_k, dw_v} -> {LN_q, LN_k, LN_v} -> q / k projections -> per-CTA partial
// channel Gram, of one TransformerBlock (model/blocks.py:264-268, 400-440), stride-1 stem blocks and
// (template S2) stride-2 branch blocks.  Included by block_tc.cu inside its anonymous namespace.
//
// One warp-specialised CTA per (clip, token chunk), 448 threads:
//   warps 0-11  COMPUTE: thread (q4, lane, third) owns token 32*q4 + lane and 48 channels.  The
//               token's x values are loaded straight into registers one tile ahead (no fp32 staging
//               tile); LN1(x) is exchanged between neighbouring tokens through a 16-bit (IEEE half)
//               shared tile, which is what frees the shared memory to keep BOTH Wq and Wk resident.
//   warp 12     MMA ISSUER (converged warp, elect-predicated tcgen05.mma): q projection as soon as
//               the q operand tile is staged (it runs under the k pass), k projection under the v
//               pass, the channel Gram (MN-major views of the q / k tiles) under the next tile's LN1.
//   warp 13     HALO: LN1 of the two tokens next to the tile (t0-1, t0+128), one tile ahead.
constexpr int kFrComp = 384;
constexpr int kFrThreads = kFrComp + 64;
constexpr uint32_t kHsRow = kC * 2;           // 272-byte row of 136 halves: 16-byte accesses of
                                              // consecutive lanes fall into distinct bank groups
struct Front1Vec {
  float ln1w[kC], ln1b[kC];
  float4 dw[3][kC];          // depthwise taps of q, k, v
  float bq[kKP], bk[kKP];    // folded biases
  float part[2][4][3][kTM];  // [parity of the pass][mean | M2 or sum | sumsq (| odd input: mean | M2)][third][token]
};
struct Front1Bars {
  uint64_t wfull;            // TMA arrival of Wq | Wk
  uint64_t qfull, kfull;     // compute -> MMA: operand tile staged (12 warp arrivals)
  uint64_t gfull;            // compute -> MMA: q / k written back for the Gram (12 warp arrivals)
  uint64_t qdone, kdone;     // projection accumulators ready
  uint64_t gdone;            // Gram UMMAs done: aq / ak reusable
  uint64_t halo_full[2];     // halo rows of tile n staged in halo[n & 1]
  uint64_t adone;            // compute warps are done reading the LN1 tile / halo rows (12 warp arrivals)
};
constexpr size_t kFront1Smem =
    (size_t)kTM * kHsRow + 4 * kHsRow + 2 * kTile144 + 2 * kW144 + sizeof(Front1Vec);
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
template <bool F16, bool S2>
__global__ void __launch_bounds__(kFrThreads, 1)
tc_front1_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const float *__restrict__ bqp,
                 const float *__restrict__ bkp, const float *__restrict__ x, float *__restrict__ gram_part,
                 uint8_t *__restrict__ vn_img, int T, int Tout, int tiles, int tiles_per_chunk, int nchunk,
                 float qscale, int trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *aq = smem;
  uint8_t *ak = aq + kTile144;
  uint8_t *wq = ak + kTile144;
  uint8_t *wk = wq + kW144;
  uint8_t *hs = wk + kW144;                    // [128][136] halves: LN1(x) of the tile's tokens
  uint8_t *halo = hs + kTM * kHsRow;           // [2][2][136] halves: LN1(x) of tokens t0-1 / t0+128
  Front1Vec *V = reinterpret_cast<Front1Vec *>(halo + 4 * kHsRow);
  __shared__ Front1Bars bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  constexpr TcPack L = tc_pack_layout();
  const int tile_begin = chunk * tiles_per_chunk;
  const int tile_end = min(tiles, tile_begin + tiles_per_chunk);

  for (int c = threadIdx.x; c < kC; c += kFrThreads) {
    V->ln1w[c] = P.ln1_w[c];
    V->ln1b[c] = P.ln1_b[c];
    V->dw[0][c] = make_float4(P.dwq[3 * c], P.dwq[3 * c + 1], P.dwq[3 * c + 2], 0.f);
    V->dw[1][c] = make_float4(P.dwk[3 * c], P.dwk[3 * c + 1], P.dwk[3 * c + 2], 0.f);
    V->dw[2][c] = make_float4(P.dwv[3 * c], P.dwv[3 * c + 1], P.dwv[3 * c + 2], 0.f);
  }
  for (int c = threadIdx.x; c < kKP; c += kFrThreads) {
    V->bq[c] = bqp[c];
    V->bk[c] = bkp[c];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars.wfull, 1);
    mbar_init(&bars.qfull, kFrComp / 32);
    mbar_init(&bars.kfull, kFrComp / 32);
    mbar_init(&bars.gfull, kFrComp / 32);
    mbar_init(&bars.adone, kFrComp / 32);
    mbar_init(&bars.qdone, 1);
    mbar_init(&bars.kdone, 1);
    mbar_init(&bars.gdone, 1);
    mbar_init(&bars.halo_full[0], 1);
    mbar_init(&bars.halo_full[1], 1);
    fence_mbar_init();
  }
  if (warp == kFrComp / 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_q = tm, t_k = tm + 144, t_g0 = tm + 288, t_g1 = tm + 368;
  constexpr uint32_t kFmt = F16 ? 0u : 1u;
  const float *xb = x + (size_t)b * kC * T;

  if (warp == kFrComp / 32) {
    // =============================================================== MMA ISSUER
    const uint32_t idesc_qk = make_idesc_16(kKP, false, false, kFmt);
    const uint32_t idesc_gram = make_idesc_16(80, true, true, kFmt);
    const uint32_t a_q = smem_u32(aq), a_k = smem_u32(ak), w_q = smem_u32(wq), w_k = smem_u32(wk);
    static_assert(tc_pack_layout().wk == tc_pack_layout().wq + kW144, "Wq | Wk are one contiguous image");
    if (tile_begin < tile_end) {
      tma_elect(wq, tcw + L.wq, 2 * kW144, &bars.wfull);
      mbar_wait(&bars.wfull, 0);
    }
    uint32_t n = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      mbar_wait(&bars.qfull, n & 1);
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_elect(t_q, make_desc(a_q + s * 2 * kCS, kCS, kRS144), make_desc(w_q + s * 2 * kCS, kCS, kRS144), idesc_qk,
                   s > 0);
      commit_elect(&bars.qdone);
      mbar_wait(&bars.kfull, n & 1);
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_elect(t_k, make_desc(a_k + s * 2 * kCS, kCS, kRS144), make_desc(w_k + s * 2 * kCS, kCS, kRS144), idesc_qk,
                   s > 0);
      commit_elect(&bars.kdone);
      // ---- channel Gram over this tile's tokens: MN-major views of the written-back q / k tiles ----
      mbar_wait(&bars.gfull, n & 1);
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < kTM / 16; ++s) {
        const uint32_t ko = s * 2 * kRS144;
        // head 0: rows = q channels 0..127, cols = k channels 0..79
        umma_elect(t_g0, make_desc(a_q + ko, kRS144, kCS), make_desc(a_k + ko, kRS144, kCS), idesc_gram,
                   !(n == 0 && s == 0));
        // head 1: rows = q channels 8..135, cols = k channels 64..143
        umma_elect(t_g1, make_desc(a_q + ko + kCS, kRS144, kCS), make_desc(a_k + ko + 8 * kCS, kRS144, kCS), idesc_gram,
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
    if (!S2 && tile_begin < tile_end) load_x(tile_begin);
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
      tr(0);
      if (S2) load_x(tile);
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
      if (S2) {   // odd input: LN1 -> 16-bit shared row; the even input is normalised in place below (xr := h)
        const float m0 = V->part[pp][2][0][tok], m1 = V->part[pp][2][1][tok], m2 = V->part[pp][2][2][tok];
        const float mu = (48.f * (m0 + m1) + 40.f * m2) * (1.0f / kC);
        const float d0 = m0 - mu, d1 = m1 - mu, d2 = m2 - mu;
        const float var = (V->part[pp][3][0][tok] + V->part[pp][3][1][tok] + V->part[pp][3][2][tok] +
                           48.f * (d0 * d0 + d1 * d1) + 40.f * d2 * d2) * (1.0f / kC);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
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
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        pp ^= 1;
        uint8_t *dst = hs + tok * kHsRow + c_lo * 2;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < nq) {
            __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c_lo + g * 8 + 2 * e;
              const float a0 = live ? fmaf((xr[g * 8 + 2 * e] - mu) * rstd, V->ln1w[c], V->ln1b[c]) : 0.f;
              const float a1 = live ? fmaf((xr[g * 8 + 2 * e + 1] - mu) * rstd, V->ln1w[c + 1], V->ln1b[c + 1]) : 0.f;
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
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        float d[48];
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < nq) {
            float l8[8], c8[8], r8[8];
            unpack8(*reinterpret_cast<const uint4 *>(lrow + g * 16), l8);
            if (S2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) c8[e] = xr[g * 8 + e];
            } else {
              unpack8(*reinterpret_cast<const uint4 *>(hrow + g * 16), c8);
            }
            unpack8(*reinterpret_cast<const uint4 *>(rrow + g * 16), r8);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 w = V->dw[m][c_lo + g * 8 + e];
              const float v = fmaf(w.z, r8[e], fmaf(w.y, c8[e], w.x * l8[e]));
              d[g * 8 + e] = v;
              s += v;
              ss = fmaf(v, v, ss);
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
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;   // (d - mean) * rstd as one FMA
        pp ^= 1;
        // the previous tile's Gram UMMAs still read aq / ak: wait before overwriting them
        if (m == 0 && n > 0) mbar_wait(&bars.gdone, (n - 1) & 1);
        tr(30 + m);
        uint8_t *dst = (m == 0 ? aq : (m == 1 ? ak : vn_tile)) + cm_offset(tok, c_lo, kRS144, kCS);
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = (g * 8 + e < nq) ? fmaf(d[g * 8 + e], rstd, nmr) : 0.f;
          *reinterpret_cast<uint4 *>(dst + g * kCS) = pack16x8<F16>(o8);
        }
        if (m < 2) {
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(m == 0 ? &bars.qfull : &bars.kfull);
        }
        tr(40 + m);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.adone);
      // next tile's x: issued after the proxy fences of the q and k passes (fence.proxy.async waits for the
      // thread's pending loads), in flight during the epilogue
      if (!S2 && tile + 1 < tile_end) load_x(tile + 1);
      // ---- epilogue: + bias, * 1/sqrt(hs) for q, 16 bit, back into aq / ak as [token][channel] ----
      {
        const uint32_t off = cm_offset(tok, c_lo, kRS144, kCS);
        float v[48];
        tr(50);
        mbar_wait(&bars.qdone, n & 1);
        tc_fence_after();
        tr(51);
        tmem_ld48(tcol(t_q, q4, c_lo), v);
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = live ? (v[g * 8 + e] + V->bq[c_lo + g * 8 + e]) * qscale : 0.f;
          *reinterpret_cast<uint4 *>(aq + off + g * kCS) = pack16x8<F16>(o8);
        }
        tr(52);
        mbar_wait(&bars.kdone, n & 1);
        tc_fence_after();
        tr(53);
        tmem_ld48(tcol(t_k, q4, c_lo), v);
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o8[e] = live ? v[g * 8 + e] + V->bk[c_lo + g * 8 + e] : 0.f;
          *reinterpret_cast<uint4 *>(ak + off + g * kCS) = pack16x8<F16>(o8);
        }
      }
      tr(54);
      fence_async_smem();
      tr(55);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.gfull);
    }
    // ---- flush the partial Gram: TMEM lane == q channel (row), column == k channel ----
    if (n > 0) {
      mbar_wait(&bars.gdone, (n - 1) & 1);
      tc_fence_after();
      if (third < 2) {
        float *gp = gram_part + (size_t)(b * nchunk + chunk) * kC * kHS;
        const int row_ch = third ? 8 + tok : tok;          // q channel of this lane
        const bool row_ok = third ? (row_ch >= kHS && row_ch < kC) : (row_ch < kHS);
        const int col0 = third ? 4 : 0;                     // first useful column
        const uint32_t tg = third ? t_g1 : t_g0;
#pragma unroll 1
        for (int g = 0; g < 10; ++g) {
          float v[8];
          tmem_ld8(tcol(tg, q4, g * 8), v);
          if (row_ok) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = g * 8 + e - col0;
              if (j >= 0 && j < kHS) gp[(size_t)row_ch * kHS + j] = v[e];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFrComp / 32) tmem_dealloc(tm, 512);
}
