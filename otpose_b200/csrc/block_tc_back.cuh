// tc_back: proj -> skip + LN2 -> MLP -> residual of one TransformerBlock (model/blocks.py:264-279,
// 441-452) on tcgen05 tensor cores.  Included by block_tc.cu inside its anonymous namespace
// (uses its tile constants, TcPack and gelu_pair).
//
// Warp-specialised, two CTAs per SM (288 threads, <= 112 registers, ~105 KB shared memory,
// 256 TMEM columns each):
//   warp 8, one lane   CONTROL: streams Wp and the 12 + 12 W1 / W2 chunk images with cp.async.bulk
//                      (TMA) through a 3-slot ring, issues every tcgen05.mma, and signals the
//                      epilogue through tcgen05.commit -> mbarrier.
//   warps 0-7          EPILOGUE: thread (q4, lane, half) owns token 32*q4 + lane and 72 of the 144
//                      accumulator columns; stages the out2 tile, computes u / LN2 / GELU / the
//                      residual, and hands operand tiles to CONTROL through mbarriers.
// There is no __syncthreads in the tile loop.  W1 chunk j+2 is issued as soon as GELU(chunk j)
// has drained its TMEM buffer and the GELU output tile is double-buffered, so the epilogue warps
// never wait for a tensor-core round trip in steady state.
//
//  * out2 tile: the (nh, T', hs) 16-bit attention output re-read as (C, T') is a set of 256-byte
//    runs (128 tokens of one channel), copied with 16-byte cp.async straight into a
//    TOKEN-contiguous ("MN-major") operand tile -- no register transpose; channel 136 of the
//    tile is a row of ones and column 136 of the Wp image holds b_p (bias folded into the MMA).
//  * LN2(u) tile: column 136 = 1, W1 image column 136 = b_1 + W_1 beta_2 (bias folded).
//  * u (the residual stream after attention) is parked in the output buffer in global memory
//    (an L2 round trip) instead of living in registers across the MLP.
// Optional phase trace (otp_debug_trace): CTA 0 records (clock64 << 8 | event) for epilogue warp 0
// (row 0) and the control thread (row 1).
constexpr int kTraceLen = 2048;
__device__ unsigned long long g_back_trace[2][kTraceLen];
struct Tracer {
  unsigned long long *p;
  int n;
  __device__ __forceinline__ void operator()(int ev) {
    if (p && n < kTraceLen) p[n++] = ((unsigned long long)clock64() << 8) | (unsigned)ev;
  }
};

constexpr int kBackEpi = 256;                 // epilogue threads (warps 0..7)
constexpr int kBackThreads = kBackEpi + 32;   // + control warp
constexpr int kBackSlots = 3;                 // weight ring slots of kW1c == kW2c bytes
constexpr int kBackLoads = 2 * kNChunk;       // weight chunk loads per tile
static_assert(kW1c == kW2c && kBackSlots * kW1c == kW144, "Wp fills the whole weight ring");
static_assert(kBackLoads % (2 * kBackSlots) == 0, "ring barrier parities repeat every tile");
static_assert(kNChunk % 4 == 0, "TMEM / H-tile barrier parities repeat every tile");

struct BackVec {
  float sa[kKP], b2[kKP], sm[kKP];
  float part[2][2][kTM];
};
struct BackBars {
  uint64_t afull;        // epilogue -> control: operand tile `a` staged (8 warp arrivals)
  uint64_t pfull;        // proj accumulator ready
  uint64_t thfull[2];    // hidden accumulator t_h[s] ready
  uint64_t hbfull[2];    // epilogue -> control: GELU tile hbuf[s] staged (8 warp arrivals)
  uint64_t hbfree[2];    // W2 chunk MMA done: hbuf[s] reusable; the last one = y accumulator ready
  uint64_t wfull[kBackSlots];   // TMA arrival of a ring slot
  uint64_t wpfull;       // TMA arrival of Wp
  uint64_t cbar;         // control's private "everything issued so far is done"
};
constexpr uint32_t kRSa = (kTM / 8) * 128;   // 2048: channel-group stride of the token-contiguous out2 tile
static_assert(kRSa * (kKP / 8) == kTile144, "out2 tile and LN2(u) tile share one buffer");

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Control-warp primitives: executed by the whole (converged) warp with warp-uniform operands, the
// instruction itself predicated on elect.sync (same leader every time), so that ptxas keeps
// descriptors in uniform registers instead of wrapping every UTCHMMA in a divergence loop.
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_elect(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])::"memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// 24 consecutive accumulator columns of this thread's TMEM lane (x16 + x8, one wait)
__device__ __forceinline__ void tmem_ld24(uint32_t taddr, float (&v)[24]) {
  uint32_t r0[16], r1[8];
  tmem_ld16_nw(taddr, r0);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7])
               : "r"(taddr + 16)
               : "memory");
  tmem_wait_ld();
  reg_fence16(r0);
  asm volatile("" : "+r"(r1[0]), "+r"(r1[1]), "+r"(r1[2]), "+r"(r1[3]), "+r"(r1[4]), "+r"(r1[5]), "+r"(r1[6]),
               "+r"(r1[7])::"memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r0[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[16 + i] = __uint_as_float(r1[i]);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBackEpi) : "memory"); }

// Weight chunk q of the per-tile MMA order W1_0, W1_1, {W2_j, W1_{j+2}}_{j=0..9}, W2_10, W2_11.
__device__ __forceinline__ size_t back_chunk_offset(const TcPack &L, int q) {
  if (q == 0) return L.w1;
  if (q == kBackLoads - 1) return L.w2 + (size_t)(kNChunk - 1) * kW2c;
  return (q & 1) ? L.w1 + (size_t)((q + 1) >> 1) * kW1c : L.w2 + (size_t)((q - 2) >> 1) * kW2c;
}

template <bool F16, bool S2>
__global__ void __launch_bounds__(kBackThreads, 2)   // 96 registers: 5 warps of one SM sub-partition x 96 x 32 <= 16 K
tc_back_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const float *__restrict__ x,
               const unsigned short *__restrict__ obuf, float *__restrict__ y, int B, int T, int Tout, int tiles,
               int trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *a = smem;                    // out2 tile (token-contiguous), then LN2(u) (channel-contiguous)
  uint8_t *hbuf = a + kTile144;         // 2 x GELU(hidden chunk)
  uint8_t *ring = hbuf + 2 * kHTile;    // 3 weight chunk slots | Wp during the proj phase
  BackVec *V = reinterpret_cast<BackVec *>(ring + kBackSlots * kW1c);
  __shared__ BackBars bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr TcPack L = tc_pack_layout();

  for (int n = threadIdx.x; n < kKP; n += kBackThreads) {
    V->sa[n] = P.sa[n];
    V->b2[n] = P.b2[n];
    V->sm[n] = P.sm[n];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars.afull, kBackEpi / 32);
    mbar_init(&bars.pfull, 1);
    mbar_init(&bars.wpfull, 1);
    mbar_init(&bars.cbar, 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.thfull[i], 1);
      mbar_init(&bars.hbfull[i], kBackEpi / 32);
      mbar_init(&bars.hbfree[i], 1);
    }
#pragma unroll
    for (int i = 0; i < kBackSlots; ++i) mbar_init(&bars.wfull[i], 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_y = tm;                           // y accumulator; the proj accumulator before that
  const uint32_t t_h[2] = {tm + 144, tm + 192};      // hidden chunk accumulators (N = 48), double-buffered
  constexpr uint32_t kFmt = F16 ? 0u : 1u;

  if (warp == 8) {
    // =============================================================== CONTROL
    {   // the whole warp runs the control flow (converged); single instructions are elect-predicated
      const uint32_t idesc144 = make_idesc_16(kKP, false, false, kFmt);
      const uint32_t idescP = make_idesc_16(kKP, true, false, kFmt);   // A = out2 tile, token-contiguous
      const uint32_t idescH = make_idesc_16(kNH, false, false, kFmt);
      const uint32_t aa = smem_u32(a), rr = smem_u32(ring);
      uint32_t cph = 0, it = 0;
      Tracer tr{(trace && blockIdx.x == 0 && lane == 0) ? g_back_trace[1] : nullptr, 0};
      for (int g = blockIdx.x; g < B * tiles; g += gridDim.x, ++it) {
        int next_load = 0;
        tr(0);
        auto load_upto = [&](int last) {   // ring slot of chunk q is free once MMA(q - 3) is done
          for (; next_load <= last && next_load < kBackLoads; ++next_load) {
            const int s = next_load % kBackSlots;
            tma_elect(ring + s * kW1c, tcw + back_chunk_offset(L, next_load), kW1c, &bars.wfull[s]);
          }
        };
        auto wait_chunk = [&](int q) { mbar_wait(&bars.wfull[q % kBackSlots], (q / kBackSlots) & 1); };
        auto drain = [&]() {               // every MMA issued so far has completed
          commit_elect(&bars.cbar);
          mbar_wait(&bars.cbar, cph);
          cph ^= 1;
        };
        auto mma1 = [&](int j, int q) {    // D_h[j&1] = [LN2(u) | 1] . [W1_j | b1_j]^T
          wait_chunk(q);
          const uint32_t w1 = rr + (q % kBackSlots) * kW1c;
#pragma unroll
          for (int s = 0; s < kKP / 16; ++s)
            umma_elect(t_h[j & 1], make_desc(aa + s * 2 * kCS, kCS, kRS144), make_desc(w1 + s * 2 * kCS, kCS, kRS144),
                      idescH, s > 0);
          commit_elect(&bars.thfull[j & 1]);
        };
        // ---- proj: D_y = [out2 | 1] . [Wp | b_p]^T (the ring is idle: last tile's MMAs were drained)
        tma_elect(ring, tcw + L.wp, kW144, &bars.wpfull);
        mbar_wait(&bars.afull, 0);
        tr(1);
        tc_fence_after();
        mbar_wait(&bars.wpfull, it & 1);
        tr(2);
#pragma unroll
        for (int s = 0; s < kKP / 16; ++s)
          umma_elect(t_y, make_desc(aa + s * 2 * kRSa, kRSa, kCS), make_desc(rr + s * 2 * kCS, kCS, kRS144), idescP, s > 0);
        commit_elect(&bars.pfull);
        tr(3);
        drain();
        tr(4);
        load_upto(kBackSlots - 1);
        // ---- MLP
        mbar_wait(&bars.afull, 1);
        tr(5);
        tc_fence_after();
        mma1(0, 0);
        mma1(1, 1);
        tr(6);
        drain();
        tr(7);
        load_upto(1 + kBackSlots);
#pragma unroll 1
        for (int j = 0; j < kNChunk; ++j) {
          const int q2 = j + 2 < kNChunk ? 2 * j + 2 : kBackLoads - kNChunk + j;   // W2_j
          mbar_wait(&bars.hbfull[j & 1], (j >> 1) & 1);
          tr(10 + j);
          tc_fence_after();
          wait_chunk(q2);
          tr(30 + j);
          const uint32_t hh = smem_u32(hbuf + (j & 1) * kHTile), w2 = rr + (q2 % kBackSlots) * kW2c;
#pragma unroll
          for (int s = 0; s < kNH / 16; ++s)
            umma_elect(t_y, make_desc(hh + s * 2 * kCS, kCS, kRS96), make_desc(w2 + s * 2 * kCS, kCS, kRS96), idesc144,
                      (j > 0 || s > 0));
          commit_elect(&bars.hbfree[j & 1]);
          int last_q = q2;
          if (j + 2 < kNChunk) {
            mma1(j + 2, 2 * j + 3);   // t_h[j&1] was drained by GELU(chunk j) before hbfull[j&1]
            last_q = 2 * j + 3;
          }
          tr(50 + j);
          drain();
          tr(70 + j);
          load_upto(last_q + kBackSlots);
        }
      }
    }
  } else {
    // =============================================================== EPILOGUE
    const int q4 = warp & 3, half = warp >> 2;
    const int tok = q4 * 32 + lane;
    const int col_lo = half * 72;
    const bool aligned = (Tout & 7) == 0;
    uint32_t it = 0;
    Tracer tr{(trace && blockIdx.x == 0 && threadIdx.x == 0) ? g_back_trace[0] : nullptr, 0};
    for (int g = blockIdx.x; g < B * tiles; g += gridDim.x, ++it) {
      const int b = g / tiles, tile = g % tiles;
      tr(0);
      const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
      const bool live = tok < nvalid;
      const int tt = t0 + tok;
      // ---- out2 tile: channel c of the (C, T') view is the contiguous run obuf_b[c*T' + t0 ...] ----
      {
        const unsigned short *ob = obuf + (size_t)b * kC * Tout + t0;
        if (aligned) {
#pragma unroll 3
          for (int i = threadIdx.x; i < kC * 16; i += kBackEpi) {
            const int c = i >> 4, j = i & 15;
            if (j * 8 < nvalid) cp_async16(a + (c >> 3) * kRSa + j * 128 + (c & 7) * 16, ob + (size_t)c * Tout + j * 8);
          }
        } else {   // T' not a multiple of 8: runs are not 16-byte aligned, element copies
          for (int i = 0; i < 72; ++i) {
            const int c = col_lo + i;
            if (c < kC)
              *reinterpret_cast<unsigned short *>(a + (c >> 3) * kRSa + (tok >> 3) * 128 + (c & 7) * 16 + (tok & 7) * 2) =
                  live ? __ldg(ob + (size_t)c * Tout + tok) : (unsigned short)0;
          }
        }
        if (threadIdx.x < 128) {   // channels 136..143: a row of ones (bias), then zeros
          const uint32_t one2 = F16 ? 0x3C003C00u : 0x3F803F80u;
          const uint32_t v = (threadIdx.x & 7) == 0 ? one2 : 0u;
          *reinterpret_cast<uint4 *>(a + (kC / 8) * kRSa + threadIdx.x * 16) = make_uint4(v, v, v, v);
        }
      }
      cp_async_commit();
      // ---- skip path pool_skip(x): the first 48 columns are loaded now (in flight while the out2 tile
      //      lands and the proj UMMA runs), the last 24 at the start of the u pass (register budget:
      //      a spilled load would serialise on its memory latency) ----
      float u[72];
      const float *xr = x + ((size_t)b * kC + col_lo) * T + (S2 ? 2 * tt : tt);
      const bool has_l = S2 && tt > 0, has_r = S2 && (2 * tt + 1 < T);
      auto skip_at = [&](int i) -> float {   // called with i = 0, 1, 2, ... (xr walks down the channels)
        float sk = 0.f;
        const float *p = xr;
        xr += T;
        if (col_lo + i < kC && live) {
          sk = __ldg(p);
          if (S2) {   // MaxPool1d(3, 2, 1)
            if (has_l) sk = fmaxf(sk, __ldg(p - 1));
            if (has_r) sk = fmaxf(sk, __ldg(p + 1));
          }
        }
        return sk;
      };
#pragma unroll
      for (int i = 0; i < 48; ++i) u[i] = skip_at(i);
      tr(1);
      cp_async_wait<0>();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.afull);
      tr(2);
#pragma unroll
      for (int i = 48; i < 72; ++i) u[i] = skip_at(i);
      mbar_wait(&bars.pfull, it & 1);
      tc_fence_after();
      tr(3);
      // ---- u = skip(x) + s_a * (proj + b_p): parked in y (global), LN2 statistics from registers ----
      {
        float *yp = y + ((size_t)b * kC + col_lo) * Tout + tt;
        float s = 0.f;
#pragma unroll
        for (int g3 = 0; g3 < 9; ++g3) {
          float v[8];
          tmem_ld8(tcol(t_y, q4, col_lo + g3 * 8), v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g3 * 8 + e, n = col_lo + i;
            const float val = (n < kC && live) ? fmaf(V->sa[n], v[e], u[i]) : 0.f;
            u[i] = val;
            s += val;
            if (n < kC && live) *yp = val;
            yp += Tout;
          }
        }
        V->part[0][half][tok] = s;
      }
      tr(4);
      epi_bar_sync();
      const float mu = (V->part[0][0][tok] + V->part[0][1][tok]) * (1.0f / kC);
      {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 72; ++i) {
          const float d = (col_lo + i < kC) ? u[i] - mu : 0.f;
          s = fmaf(d, d, s);
        }
        V->part[1][half][tok] = s;
      }
      epi_bar_sync();
      {
        const float rstd = 1.0f / sqrtf((V->part[1][0][tok] + V->part[1][1][tok]) * (1.0f / kC) + 1e-5f);
        uint8_t *dst = a + cm_offset(tok, col_lo, kRS144, kCS);
#pragma unroll
        for (int gg = 0; gg < 9; ++gg) {
          float h8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int n = col_lo + gg * 8 + e;
            h8[e] = n < kC ? (u[gg * 8 + e] - mu) * rstd : (n == kC ? 1.f : 0.f);   // column 136 = 1 (bias)
          }
          *reinterpret_cast<uint4 *>(dst + gg * kCS) = pack16x8<F16>(h8);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.afull);
      tr(5);
      // ---- MLP epilogue: GELU(hidden chunk j) -> 16-bit H tile, 24 columns per thread ----
#pragma unroll 1
      for (int j = 0; j < kNChunk; ++j) {
        const uint32_t par = (j >> 1) & 1;
        mbar_wait(&bars.thfull[j & 1], par);
        if (j >= 2) mbar_wait(&bars.hbfree[j & 1], par ^ 1);   // W2 chunk j-2 has consumed hbuf[j&1]
        tc_fence_after();
        tr(10 + j);
        const int col = half * 24;
        float hv[24];
        tmem_ld24(tcol(t_h[j & 1], q4, col), hv);
        uint8_t *dst = hbuf + (j & 1) * kHTile + cm_offset(tok, col, kRS96, kCS);
#pragma unroll
        for (int gg = 0; gg < 3; ++gg) {
          const float *v = hv + gg * 8;
          uint4 w4;
          w4.x = gelu_pair<F16>(v[0], v[1]);
          w4.y = gelu_pair<F16>(v[2], v[3]);
          w4.z = gelu_pair<F16>(v[4], v[5]);
          w4.w = gelu_pair<F16>(v[6], v[7]);
          *reinterpret_cast<uint4 *>(dst + gg * kCS) = w4;
        }
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.hbfull[j & 1]);
        tr(30 + j);
      }
      // ---- y = u + s_m * (mlp + b_2): u read back from the output buffer (all 72 loads in flight
      //      across the wait for the last W2 chunk) ----
      {
        float *yp = y + ((size_t)b * kC + col_lo) * Tout + tt;
        mbar_wait(&bars.hbfree[(kNChunk - 1) & 1], ((kNChunk - 1) >> 1) & 1);
        tc_fence_after();
        tr(50);
#pragma unroll 1
        for (int g3 = 0; g3 < 2; ++g3) {   // two batches of 36 loads in flight
          float uu[36];
          {
            const float *up = yp;
#pragma unroll
            for (int e = 0; e < 36; ++e) {
              uu[e] = (col_lo + g3 * 36 + e < kC && live) ? *up : 0.f;
              up += Tout;
            }
          }
#pragma unroll
          for (int g4 = 0; g4 < 9; ++g4) {
            float v[4];
            tmem_ld4(tcol(t_y, q4, col_lo + g3 * 36 + g4 * 4), v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int n = col_lo + g3 * 36 + g4 * 4 + e;
              if (n < kC && live) *yp = fmaf(V->sm[n], v[e] + V->b2[n], uu[g4 * 4 + e]);
              yp += Tout;
            }
          }
        }
      }
      tr(51);
      tc_fence_before();   // ordered before the next tile's afull arrival
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 256);
}
