// tc_back: proj -> skip + LN2 -> MLP -> residual of one TransformerBlock (model/blocks.py:264-279,
// 441-452) on tcgen05 tensor cores.  Included by block_tc.cu inside its anonymous namespace
// (uses its tile constants, TcPack and gelu_pair).
//
// One persistent, warp-specialised CTA per SM (576 threads, ~217 KB shared memory, 432 TMEM columns):
//   warp 16            MMA ISSUER (converged warp, single instructions elect-predicated): issues every
//                      tcgen05.mma and signals the epilogue through tcgen05.commit -> mbarrier.
//   warp 17            TMA PRODUCER: streams the weight images as 13.8 KB K-slices ([144 rows][48 K]) with
//                      cp.async.bulk through a 5-slot ring that runs ahead across tile boundaries.
//   warps 0-15         EPILOGUE: thread (q4, lane, quarter) owns token 32*q4 + lane and 36 of the 144
//                      accumulator columns; in the MLP the warps form two groups (quarter parity) that
//                      take alternate hidden chunks, 72 of the 144 hidden columns per thread, each group
//                      with its own GELU tile in shared memory.
//
// The hidden dimension (544 -> 576) is cut into FOUR chunks of 144: every UMMA of the kernel is
// M128 x N144 x K16 (73 cycles, the tensor pipe's full rate; the 64-wide chunks of the first version ran
// their N = 64 UMMAs at 52 cycles against an ideal 32 -- shared-memory operand reads bound them) and a
// tile has 4 epilogue <-> tensor-core hand-offs instead of 9.  TMEM: y accumulator (144 columns) + two
// hidden accumulators (2 x 144); the proj accumulator of the NEXT tile aliases hidden accumulator 0, which
// is idle between GELU(chunk 2) and the next tile's first W1 chunk.
//
// Software pipeline across tiles: while tile n is in its MLP, the out2 tile of tile n+1 is copied
// into the other operand buffer, its skip rows are prefetched into L2 and its proj UMMAs run, so a tile
// starts with its proj result already in TMEM.
//
//  * out2 tile: the (nh, T', hs) 16-bit attention output re-read as (C, T') is a set of 256-byte
//    runs (128 tokens of one channel), copied with 16-byte cp.async straight into a
//    TOKEN-contiguous ("MN-major") operand tile -- no register transpose; channel 136 of the
//    tile is a row of ones and column 136 of the Wp image holds b_p (bias folded into the MMA).
//  * LN2(u) tile: column 136 = 1, W1 image column 136 = b_1 + W_1 beta_2 (bias folded).
//  * u (the residual stream after attention) stays in registers across the MLP (36 values per
//    thread), so y is written exactly once.
//  * y leaves through shared memory: the fp32 [channel][token] tile is assembled in the (by then idle)
//    GELU tiles and written with one tensor-map TMA store (cp.async.bulk.tensor.2d, UTMASTG) per tile,
//    asynchronously -- 36 scalar stores per thread kept the epilogue warps blocked on the SM's 32 B/clk
//    store path for ~2.5 k cycles per tile while the tensor pipe idled (and one 512-byte cp.async.bulk per
//    channel cost the issuing warp ~80 cycles per lane).
//
// Optional phase trace (build with -DOTP_BACK_TRACE, then otp_debug_trace): CTA 0 records (clock64 << 8 | event) for epilogue warp 0
// (row 0), the control warp (row 1) and epilogue warp 15 (row 2).
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
constexpr int kBackSlots = 5;                 // ring slots of one weight piece each
constexpr int kBackHB = 2;                    // GELU tiles in shared memory: one per epilogue warp group (chunk parity)
constexpr int kBackTH = 2;                    // hidden accumulators in TMEM
constexpr int kCQ = kKP / 4;                  // 36 accumulator columns per epilogue thread
constexpr int kHC = kNH / 2;                  // 72 hidden columns per epilogue thread and chunk
constexpr int kGroupsPerTile = 9;             // weight images of a tile in MMA order (see back_group_base)
static_assert(kNChunk == 4 && kNH == kKP, "weight group order and TMEM aliasing below assume 4 chunks of 144");
static_assert(kCQ == 36 && kHC == 72, "epilogue column split");

struct BackVec {
  float sa[kKP], b2[kKP] /* s_m * b_2 */, sm[kKP];   // read as float4 broadcasts (16-byte aligned, cq % 4 == 0)
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
};
constexpr uint32_t kRSa = (kTM / 8) * 128;   // 2048: channel-group stride of the token-contiguous out2 tile
static_assert(kRSa * (kKP / 8) == kTile144, "out2 tile and LN2(u) tile share one buffer");
static_assert((size_t)kC * kTM * 4 <= (size_t)kBackHB * kTile144, "the fp32 y tile is staged in the GELU tiles");

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

// Weight image g of the per-tile MMA order  W1_0, W1_1, W2_0, W1_2, W2_1, W1_3, W2_2, Wp(next tile), W2_3
// -> index of its first K-slice in the packed piece array (Wp: 0..2, W1_c: 3 + 3c, W2_c: 15 + 3c).
__device__ __forceinline__ int back_group_base(int g) {
  switch (g) {
    case 0: return 3;
    case 1: return 6;
    case 2: return 15;
    case 3: return 9;
    case 4: return 18;
    case 5: return 12;
    case 6: return 21;
    case 7: return 0;
    default: return 24;
  }
}

template <bool F16, bool S2>
__global__ void __launch_bounds__(kBackThreads, 1)
tc_back_kernel(BlockPack P, const uint8_t *__restrict__ tcw, const uint8_t *__restrict__ tcw_lo,
               const float *__restrict__ x, const unsigned short *__restrict__ obuf, float *__restrict__ y,
               const __grid_constant__ CUtensorMap ymap, int use_tma, int B, int T, int Tout, int tiles, int trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *abuf = smem;                      // 2 x { out2 tile (token-contiguous), then LN2(u) (channel-contiguous) }
  uint8_t *hbuf = abuf + 2 * kTile144;       // kBackHB x GELU(hidden chunk) [128][144]; after the MLP: the y tile
  uint8_t *ring = hbuf + kBackHB * kTile144; // kBackSlots weight pieces
  BackVec *V = reinterpret_cast<BackVec *>(ring + kBackSlots * kPiece);
  __shared__ BackBars bars;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr TcPack L = tc_pack_layout();
  // bfloat16 mode: every W1 / W2 piece is streamed and multiplied twice (hi term, then the lo remainder
  // into the same accumulator) -- a rounded weight perturbs every token coherently, see block_tc_front.cuh
  constexpr bool kSplit = !F16;
  const int total = B * tiles;
  const int ntile = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  for (int n = threadIdx.x; n < kKP; n += kBackThreads) {
    V->sa[n] = P.sa[n];
    V->b2[n] = P.sm[n] * P.b2[n];   // s_m * b_2: added to u once the LN2 tile is written, so that E3 is one FMA
    V->sm[n] = P.sm[n];
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars.lnfull, kBackEpi / 32);
    mbar_init(&bars.pfull, 1);
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
  const uint32_t t_h = tm + kKP;       // kBackTH x hidden chunk accumulator (144 columns each)
  const uint32_t t_p = t_h;            // proj accumulator of the NEXT tile == hidden accumulator 0 (see header)
  constexpr uint32_t kFmt = F16 ? 0u : 1u;

  if (warp == kBackEpi / 32 + 1) {
    // =============================================================== TMA PRODUCER
    // The piece stream in exactly the order the MMA issuer consumes it; a slot is refilled as soon as the
    // UMMAs that read it have completed (wempty), so the stream runs up to five pieces ahead of the tensor
    // core, across tile boundaries.
    int slot = 0;
    uint32_t round = 0;   // loads / kBackSlots
    auto load = [&](const uint8_t *src) {
      if (round > 0) mbar_wait(&bars.wempty[slot], (round - 1) & 1);
      tma_elect(ring + slot * kPiece, src, kPiece, &bars.wfull[slot]);
      if (++slot == kBackSlots) {
        slot = 0;
        ++round;
      }
    };
    if (ntile > 0) {
#pragma unroll 1
      for (int s = 0; s < 3; ++s) load(tcw + L.wp + (size_t)s * kPiece);
    }
#pragma unroll 1
    for (int it = 0; it < ntile; ++it) {
#pragma unroll 1
      for (int g = 0; g < kGroupsPerTile; ++g) {
        if (g == 7 && it + 1 >= ntile) continue;   // no next tile: no Wp
        const int base = back_group_base(g);
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
          load(tcw + (size_t)(base + s) * kPiece);
          if (kSplit && g != 7) load(tcw_lo + (size_t)(base + s - 3) * kPiece);
        }
      }
    }
  } else if (warp == kBackEpi / 32) {
    // =============================================================== MMA ISSUER
    const uint32_t idesc144 = make_idesc_16(kKP, false, false, kFmt);
    const uint32_t idescP = make_idesc_16(kKP, true, false, kFmt);   // A = out2 tile, token-contiguous
    const uint32_t ab = smem_u32(abuf), rr = smem_u32(ring), hb = smem_u32(hbuf);
#ifdef OTP_BACK_TRACE
    Tracer tr{(trace && blockIdx.x == 0 && lane == 0) ? g_back_trace[1] : nullptr, 0};
#else
    auto tr = [](int) {};   // compiled out by default (build with -DOTP_BACK_TRACE for scripts/trace_back.py): the 25
                            // trace points of a tile were 4.6 % of the kernel's instructions even when switched off
#endif
    int slot = 0;          // ring slot of the next weight piece (pieces are consumed in load order)
    uint32_t round = 0;
    // One weight image = 3 K-slices of 48 (x 2 in bfloat16 mode: hi, lo): D (+)= A[:, K] . W[:, K]^T.
    // a_step: byte step of the A operand per K-step of 16; (a_lbo, a_sbo): its descriptor strides.
    auto run_group = [&](uint32_t d_tmem, uint32_t a_base, uint32_t a_step, uint32_t a_lbo, uint32_t a_sbo,
                         uint32_t idesc, bool fresh, bool split) {
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
#pragma unroll 1
        for (int pass = 0; pass < (split ? 2 : 1); ++pass) {
          mbar_wait(&bars.wfull[slot], round & 1);
          const uint32_t w = rr + slot * kPiece;
#pragma unroll
          for (int k = 0; k < 3; ++k)
            umma_elect(d_tmem, make_desc(a_base + (3 * s + k) * a_step, a_lbo, a_sbo),
                       make_desc(w + k * 2 * kCS, kCS, kRS48), idesc, !(fresh && s == 0 && k == 0 && pass == 0));
          commit_elect(&bars.wempty[slot]);
          if (++slot == kBackSlots) {
            slot = 0;
            ++round;
          }
        }
      }
    };
    auto proj = [&](int it) {          // D_p = [out2 | 1] . [Wp | b_p]^T for tile `it`
      mbar_wait(&bars.aofull[it & 1], (it >> 1) & 1);
      tc_fence_after();
      run_group(t_p, ab + (it & 1) * kTile144, 2 * kRSa, kRSa, kCS, idescP, true, false);
      commit_elect(&bars.pfull);
    };
    if (ntile > 0) proj(0);
    for (int it = 0; it < ntile; ++it) {
      const uint32_t aa = ab + (it & 1) * kTile144;
      auto mma1 = [&](int c) {   // D_h[c&1] = [LN2(u) | 1] . [W1_c | b1_c]^T
        run_group(t_h + (c & 1) * kNH, aa, 2 * kCS, kCS, kRS144, idesc144, true, kSplit);
        commit_elect(&bars.thfull[c & 1]);
      };
      tr(0);
      mbar_wait(&bars.lnfull, it & 1);
      tc_fence_after();
      tr(5);
      mma1(0);
      mma1(1);
      tr(6);
#pragma unroll 1
      for (int c = 0; c < kNChunk; ++c) {
        mbar_wait(&bars.hbfull[c & 1], (c >> 1) & 1);   // two uses of each GELU tile per tile: parity = c >> 1
        tr(10 + c);
        tc_fence_after();
        run_group(t_y, hb + (c & 1) * kTile144, 2 * kCS, kCS, kRS144, idesc144, c == 0, kSplit);
        commit_elect(&bars.hbfree[c & 1]);
        if (c == kNChunk - 1) commit_elect(&bars.yfull);
        if (c + 2 < kNChunk) mma1(c + 2);   // t_h[c&1] was drained by GELU(chunk c) before hbfull[c&1]
        if (c == 2 && it + 1 < ntile) proj(it + 1);   // t_p == t_h[0]: idle from GELU(chunk 2) to the next W1_0
        tr(50 + c);
      }
    }
  } else {
    // =============================================================== EPILOGUE
    const int q4 = warp & 3, quarter = warp >> 2;
    const int tok = q4 * 32 + lane;
    const int cq = quarter * kCQ;
    const bool aligned = (Tout & 7) == 0;
#ifdef OTP_BACK_TRACE
    Tracer tr{(trace && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15)) ? g_back_trace[warp ? 2 : 0] : nullptr, 0};
#else
    auto tr = [](int) {};
#endif
    // out2 tile of tile g: channel c of the (C, T') view is the contiguous run obuf_b[c*T' + t0 ...].
    // The copy is cut into kStageParts slices so that the next tile's slices can be issued between the
    // MLP chunks, in the slack the epilogue warps have there anyway.
    constexpr int kStageParts = (kC * 16 + kBackEpi - 1) / kBackEpi;   // 5
    auto stage_out2 = [&](int g, uint8_t *a, int part_lo, int part_hi) {
      const int b = g / tiles, t0 = (g % tiles) * kTM, nvalid = min(kTM, Tout - t0);
      const unsigned short *ob = obuf + (size_t)b * kC * Tout + t0;
      if (aligned) {
        // lane -> (channel within its group of 8, 8-token run j): 8 consecutive lanes fill one
        // 128-byte row of a core-matrix block (no shared-memory bank conflicts among them)
        for (int k = part_lo; k < part_hi; ++k) {
          const int i = threadIdx.x + k * kBackEpi;
          const int c = ((i >> 7) << 3) | (i & 7), j = (i >> 3) & 15;
          if (i < kC * 16 && j * 8 < nvalid)
            cp_async16(a + (c >> 3) * kRSa + j * 128 + (c & 7) * 16, ob + (size_t)c * Tout + j * 8);
        }
      } else if (part_lo == 0) {   // T' not a multiple of 8: runs are not 16-byte aligned, element copies
        for (int i = 0; i < kCQ; ++i) {
          const int c = cq + i;
          if (c < kC)
            *reinterpret_cast<unsigned short *>(a + (c >> 3) * kRSa + (tok >> 3) * 128 + (c & 7) * 16 + (tok & 7) * 2) =
                tok < nvalid ? __ldg(ob + (size_t)c * Tout + tok) : (unsigned short)0;
        }
      }
      if (part_lo == 0 && threadIdx.x < 128) {   // channels 136..143: a row of ones (bias), then zeros
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
      stage_out2(blockIdx.x, abuf, 0, kStageParts);
      publish_out2(0);
    }
    bool y_store_pending = false;   // this thread has a bulk store of the previous tile's y in flight
    for (int it = 0; it < ntile; ++it) {
      const int g = blockIdx.x + it * gridDim.x;
      const int b = g / tiles, tile = g % tiles;
      const int t0 = tile * kTM, nvalid = min(kTM, Tout - t0);
      const bool live = tok < nvalid;
      const int tt = t0 + tok;
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
            float2 ab2[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) {
              const float *p = xr + (size_t)(i0 + e) * T;
              ab2[e] = make_float2(0.f, -3.402823466e38f);
              if (cq + i0 + e < kC && live) {
                if (pair_ok) {
                  ab2[e] = __ldg(reinterpret_cast<const float2 *>(p));
                } else {
                  ab2[e].x = __ldg(p);
                  if (has_r) ab2[e].y = __ldg(p + 1);
                }
              }
            }
#pragma unroll
            for (int e = 0; e < 9; ++e)
              u[i0 + e] = (cq + i0 + e < kC && live) ? fmaxf(u[i0 + e], fmaxf(ab2[e].x, ab2[e].y)) : 0.f;
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
        {
          float v[kCQ];
          tmem_ld36(tcol(t_p, q4, cq), v);   // the three loads issued back to back, one wait
          const float4 *sa4 = reinterpret_cast<const float4 *>(V->sa + cq);
#pragma unroll
          for (int q = 0; q < kCQ / 4; ++q) {
            const float4 w = sa4[q];
            const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q + e, n = cq + i;
              const float val = (n < kC && live) ? fmaf(ww[e], v[i], u[i]) : 0.f;
              u[i] = val;
              s += val;
            }
          }
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
      tr(8);
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
        const float rstd = rsqrtf(var + 1e-5f);
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
        const float4 *b4 = reinterpret_cast<const float4 *>(V->b2 + cq);   // u += s_m * b_2 (LN2 is done with u)
#pragma unroll
        for (int q = 0; q < kCQ / 4; ++q) {
          const float4 w = b4[q];
          u[4 * q] += w.x;
          u[4 * q + 1] += w.y;
          u[4 * q + 2] += w.z;
          u[4 * q + 3] += w.w;
        }
      }
      // the previous tile's y tile (staged in the GELU tiles) has been read by its bulk stores: the GELU
      // tiles are only rewritten behind the MMAs that this arrival releases
      if (y_store_pending) {
        tma_store_wait_read();
        y_store_pending = false;
      }
      tr(6);
      fence_async_smem();
      tr(7);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.lnfull);
      tr(5);
      // ---- MLP epilogue: GELU(hidden chunk) -> 16-bit H tile.  The 16 warps form two groups of 8 that
      //      take ALTERNATE chunks (72 of the 144 hidden columns per thread): the chain of one chunk
      //      (accumulator wait -> tcgen05.ld -> GELU -> shared stores -> proxy fence -> arrive) overlaps
      //      the other group's chain of the next chunk instead of serialising in the same warps. ----
      const bool has_next = it + 1 < ntile;
#pragma unroll 1
      for (int c = 0; c < kNChunk; ++c) {
        if ((c & 1) == (quarter & 1)) {
          mbar_wait(&bars.thfull[c & 1], (c >> 1) & 1);
          if (it > 0 || c >= kBackHB)   // the W2 UMMAs of this GELU tile's previous use have consumed it
            mbar_wait(&bars.hbfree[c & 1], ((c >> 1) + 1) & 1);
          tc_fence_after();
          tr(10 + c);
          const int hcol = (quarter >> 1) * kHC;   // 72 hidden columns of this thread
          const uint32_t th = tcol(t_h + (c & 1) * kNH, q4, hcol);
          uint8_t *dst = hbuf + (c & 1) * kTile144 + cm_offset(tok, hcol, kRS144, kCS);
          // 72 columns as 4 x 16 + 8: the tcgen05.ld of piece i+1 is in flight while piece i is processed
          uint32_t r[2][16];
          tmem_ld16_nw(th, r[0]);
#pragma unroll
          for (int pc = 0; pc < 5; ++pc) {
            tmem_wait_ld();
            reg_fence16(r[pc & 1]);
            if (pc + 1 < 4) {
              tmem_ld16_nw(th + 16 * (pc + 1), r[(pc + 1) & 1]);
            } else if (pc + 1 == 4) {
              uint32_t(&q8)[16] = r[(pc + 1) & 1];
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(q8[0]), "=r"(q8[1]), "=r"(q8[2]), "=r"(q8[3]), "=r"(q8[4]), "=r"(q8[5]), "=r"(q8[6]),
                             "=r"(q8[7])
                           : "r"(th + 64)
                           : "memory");
            }
            const uint32_t(&rv)[16] = r[pc & 1];
#pragma unroll
            for (int gg = 0; gg < (pc < 4 ? 2 : 1); ++gg) {
              uint4 w4;
              w4.x = gelu_pair<F16>(__uint_as_float(rv[8 * gg]), __uint_as_float(rv[8 * gg + 1]));
              w4.y = gelu_pair<F16>(__uint_as_float(rv[8 * gg + 2]), __uint_as_float(rv[8 * gg + 3]));
              w4.z = gelu_pair<F16>(__uint_as_float(rv[8 * gg + 4]), __uint_as_float(rv[8 * gg + 5]));
              w4.w = gelu_pair<F16>(__uint_as_float(rv[8 * gg + 6]), __uint_as_float(rv[8 * gg + 7]));
              *reinterpret_cast<uint4 *>(dst + (2 * pc + gg) * kCS) = w4;
            }
          }
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.hbfull[c & 1]);
        }
        tr(30 + c);
        if (has_next) {
          const int gn = g + gridDim.x;
          if (c == 0) {          // next tile: out2 tile into the other operand buffer
            stage_out2(gn, abuf + ((it + 1) & 1) * kTile144, 0, 3);
          } else if (c == 1) {
            stage_out2(gn, abuf + ((it + 1) & 1) * kTile144, 3, kStageParts);
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
          } else if (c == 2) {
            publish_out2(it + 1);
          }
        }
      }
      // ---- y = u + s_m * (mlp + b_2) ----
      {
        mbar_wait(&bars.yfull, it & 1);   // last W2 chunk of the tile: y accumulator complete, GELU tiles idle
        tc_fence_after();
        tr(50);
        // the [channel][token] fp32 tile is assembled in the GELU tiles and leaves with ONE tensor-map TMA
        // store (rows = (clip, channel), columns = tokens; the part of a ragged last tile beyond T' is
        // clipped by the hardware).  Row pitches that are not a multiple of 16 bytes store directly.
        const bool bulk = use_tma != 0;
        float *stage = reinterpret_cast<float *>(hbuf);
        float *yp = y + ((size_t)b * kC + cq) * Tout + tt;
        {
          float v[kCQ];
          tmem_ld36(tcol(t_y, q4, cq), v);
          const float4 *sm4 = reinterpret_cast<const float4 *>(V->sm + cq);
#pragma unroll
          for (int q = 0; q < kCQ / 4; ++q) {
            const float4 w = sm4[q];
            const float ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * q + e, n = cq + i;
              if (n < kC) {
                const float val = fmaf(ww[e], v[i], u[i]);   // u already holds + s_m * b_2
                if (bulk) stage[n * kTM + tok] = val;
                else if (live) yp[(size_t)i * Tout] = val;
              }
            }
          }
        }
        tc_fence_before();   // ordered before this warp's next hbfull / lnfull arrival
        tr(52);
        if (bulk) {
          fence_async_smem();
          tr(53);
          epi_bar_sync();
          tr(54);
          if (threadIdx.x == 0) {
            tma_store_2d(&ymap, stage, t0, b * kC);
            tma_store_commit();
            y_store_pending = true;
          }
        }
      }
      tr(51);
    }
    if (y_store_pending) tma_store_wait_all();   // writes done before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBackEpi / 32) tmem_dealloc(tm, 512);
}
