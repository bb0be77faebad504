// tc_back: proj -> skip + LN2 -> MLP -> residual of one TransformerBlock (model/blocks.py:264-279,
// 441-452) on tcgen05 tensor cores.  Included by block_tc.cu inside its anonymous namespace
// (uses its tile constants, TcPack and gelu_pair).
//
// One persistent, warp-specialised CTA per SM (576 threads, ~218 KB shared memory, 512 TMEM columns):
//   warp 16            MMA ISSUER (converged warp, single instructions elect-predicated): issues every
//                      tcgen05.mma and signals the epilogue through tcgen05.commit -> mbarrier.
//   warp 17            TMA PRODUCER: streams the 9 + 9 W1 / W2 chunk images with cp.async.bulk through a
//                      4-slot ring that runs ahead across tile boundaries.  Wp stays resident.
//   warps 0-15         EPILOGUE: thread (q4, lane, quarter) owns token 32*q4 + lane and 36 of the 144
//                      accumulator columns; in the MLP the warps form two groups (quarter parity) that
//                      take alternate hidden chunks, 32 of the 64 hidden columns per thread, each group
//                      with its own GELU tile in shared memory.
// Software pipeline across tiles: while tile n is in its MLP, the out2 tile of tile n+1 is copied
// into the other operand buffer, its skip rows are prefetched into L2 and its proj UMMA runs into a
// second accumulator, so a tile starts with its proj result already in TMEM.  Within the MLP, W1
// chunk j+3 is issued as soon as GELU(chunk j) has drained its TMEM buffer and there is one GELU tile
// per warp group: the epilogue warps do not wait for tensor-core round trips in steady state.
//
//  * out2 tile: the (nh, T', hs) 16-bit attention output re-read as (C, T') is a set of 256-byte
//    runs (128 tokens of one channel), copied with 16-byte cp.async straight into a
//    TOKEN-contiguous ("MN-major") operand tile -- no register transpose; channel 136 of the
//    tile is a row of ones and column 136 of the Wp image holds b_p (bias folded into the MMA).
//  * LN2(u) tile: column 136 = 1, W1 image column 136 = b_1 + W_1 beta_2 (bias folded).
//  * u (the residual stream after attention) stays in registers across the MLP (36 values per
//    thread), so y is written exactly once.
//
// Optional phase trace (otp_debug_trace): CTA 0 records (clock64 << 8 | event) for epilogue warp 0
// (row 0) and the control warp (row 1).
constexpr int kTraceLen = 2048;
__device__ unsigned long long g_back_trace[4][kTraceLen];   // tc_back: epilogue warp 0, control, epilogue warp 15;
                                                            // row 3: tc_front compute warp 0
struct Tracer {
  unsigned long long *p;
  int n;
  __device__ __forceinline__ void operator()(int ev) {
    if (p && n < kTraceLen) p[n++] = ((unsigned long long)clock64() << 8) | (unsigned)ev;
  }
};

constexpr int kBackEpi = 512;                 // epilogue threads (warps 0..15)
constexpr int kBackThreads = kBackEpi + 64;   // + MMA warp (16) + TMA producer warp (17)
constexpr int kBackSlots = 4;                 // weight ring slots of kW1c == kW2c bytes: every MLP chunk consumes two
                                              // (W2_j, W1_j+3), so four slots keep the TMA stream two chunks ahead
constexpr int kBackHB = 2;                    // GELU tiles in shared memory: one per epilogue warp group (chunk parity)
constexpr int kBackTH = 3;                    // hidden accumulators in TMEM:
                                              // W1 chunk c+3 is issued when GELU(c) is done
constexpr int kBackLoads = 2 * kNChunk;       // weight chunk loads per tile
constexpr int kCQ = kKP / 4;                  // 36 accumulator columns per epilogue thread
constexpr int kHQ = kNH / 4;                  // 16 hidden columns per epilogue thread and chunk
static_assert(kW1c == kW2c, "W1 and W2 chunk images share the ring slots");
static_assert(kNChunk == 9, "weight chunk order and next-tile staging schedule below");
static_assert(kHQ == 16 && kCQ == 36, "epilogue column split");

struct BackVec {
  float sa[kKP], b2[kKP], sm[kKP];
  float part[2][4][kTM];
};
struct BackBars {
  uint64_t aofull[2];    // epilogue -> control: out2 tile staged in abuf[s] (16 warp arrivals)
  uint64_t lnfull;       // epilogue -> control: LN2(u) tile staged (16 warp arrivals)
  uint64_t pfull;        // proj accumulator ready
  uint64_t thfull[kBackTH];   // hidden accumulator t_h[s] ready
  uint64_t hbfull[kBackHB];   // epilogue -> control: GELU tile hbuf[s] staged (8 warp arrivals: the chunk's group)
  uint64_t hbfree[kBackHB];   // W2 chunk MMA done: hbuf[s] reusable
  uint64_t yfull;             // a tile's last W2 chunk MMA done: y accumulator complete (one phase per tile)
  uint64_t wfull[kBackSlots];    // TMA arrival of a ring slot
  uint64_t wempty[kBackSlots];   // the MMAs reading a ring slot are done
  uint64_t wpfull;       // TMA arrival of Wp
};
constexpr uint32_t kRSa = (kTM / 8) * 128;   // 2048: channel-group stride of the token-contiguous out2 tile
constexpr uint32_t kRSH = (kNH / 8) * 128;   // row-group stride of a K = kNH tile
static_assert(kRSa * (kKP / 8) == kTile144, "out2 tile and LN2(u) tile share one buffer");

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Control-warp primitives: executed by the whole (converged) warp with warp-uniform operands, the
// instruction itself predicated on elect.sync (same leader every time), so that ptxas keeps the
// issue sequence straight-line instead of wrapping every UTCHMMA in a divergence loop.
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBackEpi) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Weight chunk q of the per-tile MMA order W1_0, W1_1, W1_2, {W2_j, W1_{j+3}}_{j=0..5}, W2_6, W2_7, W2_8.
__device__ __forceinline__ size_t back_chunk_offset(const TcPack &L, int q) {
  if (q < 3) return L.w1 + (size_t)q * kW1c;
  if (q >= 15) return L.w2 + (size_t)(q - 9) * kW2c;
  return (q & 1) ? L.w2 + (size_t)((q - 3) >> 1) * kW2c : L.w1 + (size_t)((q + 3) >> 1) * kW1c;
}

template <bool F16, bool S2>
__global__ void __launch_bounds__(kBackThreads, 1)
tc_back_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const uint8_t *__restrict__ tcw_lo,
               const float *__restrict__ x,
               const unsigned short *__restrict__ obuf, float *__restrict__ y, int B, int T, int Tout, int tiles,
               int trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *abuf = smem;                      // 2 x { out2 tile (token-contiguous), then LN2(u) (channel-contiguous) }
  uint8_t *wpb = abuf + 2 * kTile144;        // Wp, resident
  uint8_t *hbuf = wpb + kW144;               // kBackHB x GELU(hidden chunk)
  uint8_t *ring = hbuf + kBackHB * kHTile;   // kBackSlots weight chunk slots
  BackVec *V = reinterpret_cast<BackVec *>(ring + kBackSlots * kW1c);
  __shared__ BackBars bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr TcPack L = tc_pack_layout();
  // bfloat16 mode: every W1 / W2 chunk is streamed and multiplied twice (hi term, then the lo remainder
  // into the same accumulator) -- a rounded weight perturbs every token coherently, see block_tc_front.cuh
  constexpr bool kSplit = !F16;
  constexpr int kPasses = kSplit ? 2 : 1;
  const int total = B * tiles;
  const int ntile = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  for (int n = threadIdx.x; n < kKP; n += kBackThreads) {
    V->sa[n] = P.sa[n];
    V->b2[n] = P.b2[n];
    V->sm[n] = P.sm[n];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars.lnfull, kBackEpi / 32);
    mbar_init(&bars.pfull, 1);
    mbar_init(&bars.wpfull, 1);
    mbar_init(&bars.yfull, 1);
#pragma unroll
    for (int i = 0; i < kBackTH; ++i) mbar_init(&bars.thfull[i], 1);
#pragma unroll
    for (int i = 0; i < kBackHB; ++i) {
      mbar_init(&bars.hbfull[i], kBackEpi / 64);
      mbar_init(&bars.hbfree[i], 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) mbar_init(&bars.aofull[i], kBackEpi / 32);
#pragma unroll
    for (int i = 0; i < kBackSlots; ++i) {
      mbar_init(&bars.wfull[i], 1);
      mbar_init(&bars.wempty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kBackEpi / 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t t_y = tm;             // MLP output accumulator (144 columns)
  const uint32_t t_p = tm + kKP;       // proj accumulator of the NEXT tile (144 columns)
  const uint32_t t_h = tm + 2 * kKP;   // kBackTH x hidden chunk accumulator (kNH columns each)
  constexpr uint32_t kFmt = F16 ? 0u : 1u;

  if (warp == kBackEpi / 32 + 1) {
    // =============================================================== TMA PRODUCER
    // Wp once, then the weight chunk stream of every tile of this CTA through the ring; a slot is
    // refilled as soon as the MMAs that read it have completed (wempty), so the stream runs up to
    // three chunks ahead of the tensor core, across tile boundaries.
    if (ntile > 0) tma_elect(wpb, tcw + L.wp, kW144, &bars.wpfull);
    const int nload = kBackLoads * ntile;
    int q = 0, slot = 0;
    uint32_t round = 0;   // loads / kBackSlots
    for (int gq = 0; gq < nload; ++gq) {
      const size_t off = back_chunk_offset(L, q);
#pragma unroll
      for (int pass = 0; pass < kPasses; ++pass) {
        if (round > 0) mbar_wait(&bars.wempty[slot], (round - 1) & 1);
        tma_elect(ring + slot * kW1c, pass ? tcw_lo + (off - L.w1) : tcw + off, kW1c, &bars.wfull[slot]);
        if (++slot == kBackSlots) {
          slot = 0;
          ++round;
        }
      }
      if (++q == kBackLoads) q = 0;
    }
  } else if (warp == kBackEpi / 32) {
    // =============================================================== MMA ISSUER
    const uint32_t idesc144 = make_idesc_16(kKP, false, false, kFmt);
    const uint32_t idescP = make_idesc_16(kKP, true, false, kFmt);   // A = out2 tile, token-contiguous
    const uint32_t idescH = make_idesc_16(kNH, false, false, kFmt);
    const uint32_t ab = smem_u32(abuf), rr = smem_u32(ring), hb = smem_u32(hbuf), wp = smem_u32(wpb);
    Tracer tr{(trace && blockIdx.x == 0 && lane == 0) ? g_back_trace[1] : nullptr, 0};
    int slot = 0;          // ring slot of the next weight chunk (chunks are consumed in load order)
    uint32_t round = 0;
    // wait for the next chunk of the stream; returns its shared-memory address
    auto next_chunk = [&]() -> uint32_t {
      mbar_wait(&bars.wfull[slot], round & 1);
      return rr + slot * kW1c;
    };
    auto release_chunk = [&]() {   // after the MMAs reading it have been issued
      commit_elect(&bars.wempty[slot]);
      if (++slot == kBackSlots) {
        slot = 0;
        ++round;
      }
    };
    auto proj = [&](int it) {          // D_p = [out2 | 1] . [Wp | b_p]^T for tile `it`
      mbar_wait(&bars.aofull[it & 1], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t aa = ab + (it & 1) * kTile144;
#pragma unroll
      for (int s = 0; s < kKP / 16; ++s)
        umma_elect(t_p, make_desc(aa + s * 2 * kRSa, kRSa, kCS), make_desc(wp + s * 2 * kCS, kCS, kRS144), idescP, s > 0);
      commit_elect(&bars.pfull);
    };
    if (ntile > 0) {
      mbar_wait(&bars.wpfull, 0);
      proj(0);
    }
    for (int it = 0; it < ntile; ++it) {
      const int c0 = kNChunk * it;
      const uint32_t aa = ab + (it & 1) * kTile144;
      auto mma1 = [&](int c) {   // D_h[c%3] = [LN2(u) | 1] . [W1_j | b1_j]^T
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const uint32_t w1 = next_chunk();
#pragma unroll
          for (int s = 0; s < kKP / 16; ++s)
            umma_elect(t_h + (c % kBackTH) * kNH, make_desc(aa + s * 2 * kCS, kCS, kRS144),
                       make_desc(w1 + s * 2 * kCS, kCS, kRS144), idescH, (pass > 0 || s > 0));
          release_chunk();
        }
        commit_elect(&bars.thfull[c % kBackTH]);
      };
      tr(0);
      mbar_wait(&bars.lnfull, it & 1);
      tc_fence_after();
      tr(5);
      mma1(c0);
      mma1(c0 + 1);
      mma1(c0 + 2);
      tr(6);
#pragma unroll 1
      for (int j = 0; j < kNChunk; ++j) {
        const int c = c0 + j;
        mbar_wait(&bars.hbfull[c % kBackHB], (c / kBackHB) & 1);
        tr(10 + j);
        tc_fence_after();
        const uint32_t hh = hb + (c % kBackHB) * kHTile;
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass) {
          const uint32_t w2 = next_chunk();
          if (pass == 0) tr(30 + j);
#pragma unroll
          for (int s = 0; s < kNH / 16; ++s)
            umma_elect(t_y, make_desc(hh + s * 2 * kCS, kCS, kRSH), make_desc(w2 + s * 2 * kCS, kCS, kRSH), idesc144,
                       (j > 0 || s > 0 || pass > 0));
          release_chunk();
        }
        commit_elect(&bars.hbfree[c % kBackHB]);
        if (j == kNChunk - 1) commit_elect(&bars.yfull);
        if (j + 3 < kNChunk) mma1(c + 3);   // t_h[c%3] was drained by GELU(chunk c) before hbfull[c%3]
        if (j == kNChunk - 2 && it + 1 < ntile) proj(it + 1);   // its out2 tile is published at chunk 6;
                                                                // E1 of this tile has left t_p (lnfull above)
        tr(50 + j);
      }
    }
  } else {
    // =============================================================== EPILOGUE
    const int q4 = warp & 3, quarter = warp >> 2;
    const int tok = q4 * 32 + lane;
    const int cq = quarter * kCQ;
    const bool aligned = (Tout & 7) == 0;
    Tracer tr{(trace && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15)) ? g_back_trace[warp ? 2 : 0] : nullptr, 0};
    // out2 tile of tile g: channel c of the (C, T') view is the contiguous run obuf_b[c*T' + t0 ...].
    // The copy is cut into kStageParts slices (part < 0: all of them) so that the next tile's slices
    // can be issued one per MLP chunk, in the slack the epilogue warps have there anyway.
    constexpr int kStageParts = (kC * 16 + kBackEpi - 1) / kBackEpi;   // 5
    auto stage_out2 = [&](int g, uint8_t *a, int part) {
      const int b = g / tiles, t0 = (g % tiles) * kTM, nvalid = min(kTM, Tout - t0);
      const unsigned short *ob = obuf + (size_t)b * kC * Tout + t0;
      if (aligned) {
        // lane -> (channel within its group of 8, 8-token run j): 8 consecutive lanes fill one
        // 128-byte row of a core-matrix block (no shared-memory bank conflicts among them)
        for (int k = (part < 0 ? 0 : part); k < (part < 0 ? kStageParts : part + 1); ++k) {
          const int i = threadIdx.x + k * kBackEpi;
          const int c = ((i >> 7) << 3) | (i & 7), j = (i >> 3) & 15;
          if (i < kC * 16 && j * 8 < nvalid)
            cp_async16(a + (c >> 3) * kRSa + j * 128 + (c & 7) * 16, ob + (size_t)c * Tout + j * 8);
        }
      } else if (part <= 0) {   // T' not a multiple of 8: runs are not 16-byte aligned, element copies
        for (int i = 0; i < kCQ; ++i) {
          const int c = cq + i;
          if (c < kC)
            *reinterpret_cast<unsigned short *>(a + (c >> 3) * kRSa + (tok >> 3) * 128 + (c & 7) * 16 + (tok & 7) * 2) =
                tok < nvalid ? __ldg(ob + (size_t)c * Tout + tok) : (unsigned short)0;
        }
      }
      if (part <= 0 && threadIdx.x < 128) {   // channels 136..143: a row of ones (bias), then zeros
        const uint32_t one2 = F16 ? 0x3C003C00u : 0x3F803F80u;
        const uint32_t v = (threadIdx.x & 7) == 0 ? one2 : 0u;
        *reinterpret_cast<uint4 *>(a + (kC / 8) * kRSa + threadIdx.x * 16) = make_uint4(v, v, v, v);
      }
      cp_async_commit();
    };
    auto publish_out2 = [&](int it) {
      cp_async_wait<0>();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.aofull[it & 1]);
    };
    if (ntile > 0) {
      stage_out2(blockIdx.x, abuf, -1);
      publish_out2(0);
    }
    for (int it = 0; it < ntile; ++it) {
      const int g = blockIdx.x + it * gridDim.x;
      const int b = g / tiles, tile = g % tiles;
      const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
      const bool live = tok < nvalid;
      const int tt = t0 + tok;
      const int c0 = kNChunk * it;
      uint8_t *a = abuf + (it & 1) * kTile144;
      tr(0);
      // ---- skip path pool_skip(x) (rows prefetched into L2 during the previous tile's MLP) ----
      float u[kCQ];
      {
        const float *xr = x + ((size_t)b * kC + cq) * T + (S2 ? 2 * tt : tt);
        const bool has_l = S2 && tt > 0, has_r = S2 && (2 * tt + 1 < T);
        if (!S2) {
#pragma unroll
          for (int i = 0; i < kCQ; ++i) {
            u[i] = (cq + i < kC && live) ? __ldg(xr) : 0.f;
            xr += T;
          }
        } else {
          // MaxPool1d(3, 2, 1) over inputs 2t-1, 2t, 2t+1.  Two explicit phases so that every load of
          // a phase is in flight at once within the register budget (a load whose result is consumed
          // at once would serialise on the memory latency): the 36 left taps, then the (2t, 2t+1)
          // pairs in batches of 9.
          const bool pair_ok = (T & 1) == 0;   // (2t, 2t+1) is one aligned 8-byte load
          const float *xl = xr;
#pragma unroll
          for (int i = 0; i < kCQ; ++i) {
            u[i] = (cq + i < kC && live && has_l) ? __ldg(xl - 1) : -3.402823466e38f;
            xl += T;
          }
#pragma unroll
          for (int i0 = 0; i0 < kCQ; i0 += 9) {
            float2 ab[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) {
              const float *p = xr + (size_t)(i0 + e) * T;
              ab[e] = make_float2(0.f, -3.402823466e38f);
              if (cq + i0 + e < kC && live) {
                if (pair_ok) {
                  ab[e] = __ldg(reinterpret_cast<const float2 *>(p));
                } else {
                  ab[e].x = __ldg(p);
                  if (has_r) ab[e].y = __ldg(p + 1);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 9; ++e)
              u[i0 + e] = (cq + i0 + e < kC && live) ? fmaxf(u[i0 + e], fmaxf(ab[e].x, ab[e].y)) : 0.f;
          }
        }
      }
      tr(1);
      mbar_wait(&bars.pfull, it & 1);
      tc_fence_after();
      tr(3);
      // ---- u = skip(x) + s_a * (proj + b_p): stays in registers across the MLP (36 per thread) ----
      {
        float s = 0.f;
        auto piece = [&](const float *v, int i0, int cnt) {
#pragma unroll
          for (int e = 0; e < cnt; ++e) {
            const int i = i0 + e, n = cq + i;
            const float val = (n < kC && live) ? fmaf(V->sa[n], v[e], u[i]) : 0.f;
            u[i] = val;
            s += val;
          }
        };
        {
          float v[16];
          tmem_ld16(tcol(t_p, q4, cq), v);
          piece(v, 0, 16);
          tmem_ld16(tcol(t_p, q4, cq + 16), v);
          piece(v, 16, 16);
          float v4[4];
          tmem_ld4(tcol(t_p, q4, cq + 32), v4);
          piece(v4, 32, 4);
        }
        // per-thread (mean, M2) over its n_q valid columns, combined across the 4 quarters with the
        // parallel-variance formula: one barrier, no E[x^2] - mean^2 cancellation
        const int nq = quarter == 3 ? kC - 3 * kCQ : kCQ;
        const float mq = s / (float)nq;
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < kCQ; ++i) {
          const float d = (cq + i < kC) ? u[i] - mq : 0.f;
          m2 = fmaf(d, d, m2);
        }
        V->part[0][quarter][tok] = mq;
        V->part[1][quarter][tok] = m2;
      }
      tr(4);
      epi_bar_sync();
      float mu, var;
      {
        const float m0 = V->part[0][0][tok], m1 = V->part[0][1][tok], m2 = V->part[0][2][tok], m3 = V->part[0][3][tok];
        constexpr float kN3 = (float)(kC - 3 * kCQ);
        mu = ((float)kCQ * (m0 + m1 + m2) + kN3 * m3) * (1.0f / kC);
        const float d0 = m0 - mu, d1 = m1 - mu, d2 = m2 - mu, d3 = m3 - mu;
        var = (V->part[1][0][tok] + V->part[1][1][tok] + V->part[1][2][tok] + V->part[1][3][tok] +
               (float)kCQ * (d0 * d0 + d1 * d1 + d2 * d2) + kN3 * d3 * d3) * (1.0f / kC);
      }
      {
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        float h[kCQ];
#pragma unroll
        for (int i = 0; i < kCQ; ++i) {
          const int n = cq + i;
          h[i] = n < kC ? (u[i] - mu) * rstd : (n == kC ? 1.f : 0.f);   // column 136 = 1 (bias)
        }
        // 36 columns = 4 full 16-byte operand chunks and one half chunk (leading for odd quarters)
        const int lead = (quarter & 1) ? 4 : 0;
        uint8_t *dst = a + cm_offset(tok, cq + lead, kRS144, kCS);
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          float h8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) h8[e] = (quarter & 1) ? h[4 + gg * 8 + e] : h[gg * 8 + e];
          *reinterpret_cast<uint4 *>(dst + gg * kCS) = pack16x8<F16>(h8);
        }
        uint2 hv2;   // the half chunk: h[0..3] (odd quarters) or h[32..35] (even quarters)
        hv2.x = (quarter & 1) ? pack16x2<F16>(h[0], h[1]) : pack16x2<F16>(h[32], h[33]);
        hv2.y = (quarter & 1) ? pack16x2<F16>(h[2], h[3]) : pack16x2<F16>(h[34], h[35]);
        *reinterpret_cast<uint2 *>(a + cm_offset(tok, cq + ((quarter & 1) ? 0 : 32), kRS144, kCS)) = hv2;
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.lnfull);
      tr(5);
      // ---- MLP epilogue: GELU(hidden chunk) -> 16-bit H tile.  The 16 warps form two groups of 8 that
      //      take ALTERNATE chunks (32 of the 64 hidden columns per thread): the chain of one chunk
      //      (accumulator wait -> tcgen05.ld -> GELU -> shared stores -> proxy fence -> arrive) overlaps
      //      the other group's chain of the next chunk instead of serialising in the same warps.  A
      //      group waits on every other phase of thfull / hbfree; that is safe because it has seen chunk
      //      c-2 (issued after chunk c-3, the previous phase of the same slot) before it waits for c. ----
      const bool has_next = it + 1 < ntile;
#pragma unroll 1
      for (int j = 0; j < kNChunk; ++j) {
        const int c = c0 + j;
        if ((c & 1) == (quarter & 1)) {
          const uint32_t par = (c / kBackTH) & 1;
          mbar_wait(&bars.thfull[c % kBackTH], par);
          if (c >= kBackHB)   // W2 chunk c-2 (this group's previous chunk) has consumed hbuf[c%2]
            mbar_wait(&bars.hbfree[c % kBackHB], ((c / kBackHB) & 1) ^ 1);
          tc_fence_after();
          tr(10 + j);
          const int hcol = (quarter >> 1) * 2 * kHQ;   // 32 hidden columns of this thread
          float hv[2 * kHQ];
          tmem_ld32(tcol(t_h + (c % kBackTH) * kNH, q4, hcol), hv);
          uint8_t *dst = hbuf + (c % kBackHB) * kHTile + cm_offset(tok, hcol, kRSH, kCS);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
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
          if (lane == 0) mbar_arrive(&bars.hbfull[c % kBackHB]);
        }
        tr(30 + j);
        if (has_next) {
          const int gn = g + gridDim.x;
          if (j < kStageParts) {   // next tile: out2 tile into the other operand buffer, a slice per chunk
            stage_out2(gn, abuf + ((it + 1) & 1) * kTile144, j);
          } else if (j == kStageParts) {
            // skip rows of the next tile into L2: this warp's 32 tokens x 36 channels, one lane per
            // (channel, 128-byte line) -- prefetches of the lanes of a warp are not coalesced
            const int bn = gn / tiles, tn0 = (gn % tiles) * kTM + q4 * 32;
            if (tn0 < Tout) {
              const float *xn = x + ((size_t)bn * kC + cq) * T + (S2 ? 2 * tn0 : tn0);
              for (int i = lane; i < (S2 ? 2 : 1) * kCQ; i += 32) {
                const int ch = S2 ? (i >> 1) : i;
                if (cq + ch < kC) prefetch_l2(xn + (size_t)ch * T + (S2 ? (i & 1) * 32 : 0));
              }
            }
          } else if (j == kStageParts + 1) {
            publish_out2(it + 1);
          }
        }
      }
      // ---- y = u + s_m * (mlp + b_2) ----
      {
        float *yp = y + ((size_t)b * kC + cq) * Tout + tt;
        const int cl = c0 + kNChunk - 1;
        (void)cl;
        mbar_wait(&bars.yfull, it & 1);   // last W2 chunk of the tile: y accumulator complete
        tc_fence_after();
        tr(50);
        auto piece = [&](const float *v, int i0, int cnt) {
#pragma unroll
          for (int e = 0; e < cnt; ++e) {
            const int i = i0 + e, n = cq + i;
            if (n < kC && live) *yp = fmaf(V->sm[n], v[e] + V->b2[n], u[i]);
            yp += Tout;
          }
        };
        float v[16];
        tmem_ld16(tcol(t_y, q4, cq), v);
        piece(v, 0, 16);
        tmem_ld16(tcol(t_y, q4, cq + 16), v);
        piece(v, 16, 16);
        float v4[4];
        tmem_ld4(tcol(t_y, q4, cq + 32), v4);
        piece(v4, 32, 4);
      }
      tr(51);
      tc_fence_before();   // ordered before this warp's next hbfull / lnfull arrival
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBackEpi / 32) tmem_dealloc(tm, 512);
}
