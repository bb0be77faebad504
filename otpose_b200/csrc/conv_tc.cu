// a7 (16-bit tensor-core path): the small-channel Conv2d (+ folded BatchNorm + ReLU, fused input
// add / residual) of model/RSB.py:106-139 as an implicit GEMM on tcgen05 -- WITHOUT materialising
// an im2col tile.
//
// Tile = 128 consecutive pixels of one image (W % 8 == 0).  The input rows the tile's taps touch
// (pixels p0-W .. p0+127+W, 8-pixel chunks) are converted to 16 bit once and stored three times,
// pre-shifted by dx = -1 / 0 / +1 with the row-edge zeros applied, in the layout
//
//     addr(dx, chunk j, channel c, e) = S[dx] + j*(Cg*128) + (c/8)*128 + (c%8)*16 + e*2
//
// i.e. 8 channels x 8 pixels "core matrices" whose 16-byte rows are runs of 8 pixels.  Read as a
// pixel-contiguous (MN-major) A operand with LBO = 128 (next 8 channels) and SBO = Cg*128 (next
// 8 pixels), the operand of tap (dy, dx) is simply the shifted copy dx starting W/8 * dy chunks
// further: 9 taps x Cpad/16 UMMAs (M128 x Npad x K16) read the staged rows in place.  K index of
// the packed weight image = tap * Cpad + channel.  Epilogue: TMEM -> + bias (+ residual) -> ReLU
// -> coalesced fp32 stores.  ~70 KB of shared memory and 32..96 TMEM columns per 128-thread CTA,
// so three CTAs per SM overlap each other's staging / UMMA / epilogue phases.
#include "common.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;
namespace {

constexpr int kCtMaxThreads = 256;   // CTAs of 128 threads (small convs: more CTAs per SM) or 256 threads
                                     // (two warps per TMEM lane quarter split the output channels)
constexpr int kCtTM = 128;   // pixels per tile

struct ConvTcShape {
  int cpad, npad, taps, kdim;          // channels padded to 16, outputs padded to 16, 1 or 9, taps*cpad
  uint32_t w_bytes, s_chunks, s_bytes;  // weight image, staged chunks per shifted copy, bytes per copy
  size_t smem;
};
__host__ __device__ inline ConvTcShape conv_tc_shape(int cin, int cout, int k, int w) {
  ConvTcShape s;
  s.cpad = (cin + 15) / 16 * 16;
  s.npad = (cout + 15) / 16 * 16;
  s.taps = k * k;
  s.kdim = s.taps * s.cpad;
  s.w_bytes = (uint32_t)s.npad * s.kdim * 2;
  s.s_chunks = kCtTM / 8 + (k == 3 ? 2 * (w / 8) : 0);
  s.s_bytes = s.s_chunks * (uint32_t)(s.cpad / 8) * 128;
  s.smem = (size_t)s.w_bytes + (size_t)(k == 3 ? 3 : 1) * s.s_bytes;
  return s;
}

// kU = staged items per thread kept in flight (loads issued before the first use); kMinCta sizes the
// register budget so that as many CTAs as the shared memory allows are resident per SM.
// the CTA-wide sync that publishes the TMEM base address, the mbarrier init and the cp.async'ed weights
__device__ __forceinline__ uint32_t tmem_slot_after_sync(const uint32_t *slot) {
  cp_async_wait<0>();
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *slot;
}

template <bool F16, int kU, int kMinCta, int kCtThreads>
__global__ void __launch_bounds__(kCtThreads, kMinCta)
conv_tc_kernel(const float *__restrict__ x, long long x_bs, const float *__restrict__ x_add, long long xa_bs,
               const uint8_t *__restrict__ wimg, const float *__restrict__ bias, const float *__restrict__ residual,
               long long r_bs, float *__restrict__ y, long long y_bs, int B, int cin, int H, int W, int cout, int k,
               int relu, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[96];
  const ConvTcShape S = conv_tc_shape(cin, cout, k, W);
  uint8_t *wsm = smem;
  uint8_t *stg = smem + S.w_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = H * W;
  const int cg = S.cpad / 8, wch = W / 8;
  const int halo = k == 3 ? wch : 0;   // chunks staged before the tile's first pixel
  // two accumulators (UMMAs of the next tile issued before this tile's epilogue) up to 64 output
  // channels; wider convs keep one accumulator so that 3-4 CTAs per SM still fit the 512 TMEM columns
  const bool dbl = S.npad <= 64;
  const uint32_t ncols = S.npad <= 16 ? 32 : (S.npad <= 32 ? 64 : 128);

  // weights: contiguous 16-bit operand image (K-major rows of taps*cpad)
  for (uint32_t o = threadIdx.x * 16; o < S.w_bytes; o += kCtThreads * 16) cp_async16(wsm + o, wimg + o);
  cp_async_commit();
  for (int n = threadIdx.x; n < S.npad; n += kCtThreads) sbias[n] = n < cout ? bias[n] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, ncols);

  // ---- tile-invariant decomposition of this thread's staged items (channel, 8-pixel chunk) ----
  // lane -> channel within its group of 8 (16 bytes apart in shared memory: 8 lanes fill one
  // 128-byte core matrix, no bank conflicts), then the chunk, then the channel group
  const int nitem = S.cpad * (int)S.s_chunks;
  int i_src[kU];        // c * P + (j - halo) * 8: element offset from the image base at tile pixel 0
  int i_dj[kU];         // (j - halo) * 8
  int i_jw[kU];         // i_dj mod W, in [0, W)
  uint32_t i_dst[kU];   // byte offset inside a shifted copy; ~0u = no item, bit 31..: see i_ld
  bool i_ld[kU];        // real channel (c < cin): has global data
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int it = threadIdx.x + u * kCtThreads;
    const int j = (it >> 3) % (int)S.s_chunks, c = ((it >> 3) / (int)S.s_chunks) * 8 + (it & 7);
    i_dj[u] = (j - halo) * 8;
    i_jw[u] = ((i_dj[u] % W) + W) % W;
    i_src[u] = c * P + i_dj[u];
    i_ld[u] = it < nitem && c < cin;
    i_dst[u] = it < nitem ? (uint32_t)(j * cg * 128 + (c >> 3) * 128 + (c & 7) * 16) : 0xffffffffu;
  }
  float v[kU][10];   // pixels q-1 .. q+8 of each item
  // global -> registers for tile g (all loads of the tile issued back to back)
  auto load_tile = [&](int g) {
    const int b = g / tiles, p0 = (g % tiles) * kCtTM, pw = p0 % W;
    const float *xb = x + (size_t)b * x_bs + p0;
    const float *ab = x_add ? x_add + (size_t)b * xa_bs + p0 : nullptr;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
#pragma unroll
      for (int e = 0; e < 10; ++e) v[u][e] = 0.f;
      const int q = p0 + i_dj[u];
      if (i_ld[u] && q >= 0 && q < P) {   // P % 8 == 0: a chunk is entirely inside or outside the image
        int col0 = pw + i_jw[u];
        if (col0 >= W) col0 -= W;
        const float *src = xb + i_src[u];
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(src));
        const float4 a1 = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        v[u][1] = a0.x, v[u][2] = a0.y, v[u][3] = a0.z, v[u][4] = a0.w;
        v[u][5] = a1.x, v[u][6] = a1.y, v[u][7] = a1.z, v[u][8] = a1.w;
        if (k == 3) {
          if (col0 > 0) v[u][0] = __ldg(src - 1);          // same image row (W % 8 == 0)
          if (col0 + 8 < W) v[u][9] = __ldg(src + 8);
        }
        if (ab) {
          const float *as = ab + i_src[u];
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(as));
          const float4 b1 = __ldg(reinterpret_cast<const float4 *>(as) + 1);
          v[u][1] += b0.x, v[u][2] += b0.y, v[u][3] += b0.z, v[u][4] += b0.w;
          v[u][5] += b1.x, v[u][6] += b1.y, v[u][7] += b1.z, v[u][8] += b1.w;
          if (k == 3) {
            if (col0 > 0) v[u][0] += __ldg(as - 1);
            if (col0 + 8 < W) v[u][9] += __ldg(as + 8);
          }
        }
      }
    }
  };
  // registers -> the three pre-shifted 16-bit copies
  auto store_tile = [&]() {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (i_dst[u] == 0xffffffffu) continue;
      uint8_t *dst = stg + i_dst[u];
      if (k == 3) {
        *reinterpret_cast<uint4 *>(dst) = pack16x8<F16>(v[u]);                       // dx = -1: pixel e <- x[q+e-1]
        *reinterpret_cast<uint4 *>(dst + S.s_bytes) = pack16x8<F16>(v[u] + 1);       // dx =  0
        *reinterpret_cast<uint4 *>(dst + 2 * S.s_bytes) = pack16x8<F16>(v[u] + 2);   // dx = +1
      } else {
        *reinterpret_cast<uint4 *>(dst) = pack16x8<F16>(v[u] + 1);
      }
    }
  };

  // Persistent over (image, tile): TMEM, barrier and weights are set up once per CTA.  Software
  // pipeline, two accumulators in TMEM: the UMMAs of tile n+1 are issued BEFORE the epilogue of
  // tile n, and the global loads of tile n+2 are in flight across that epilogue.  The proxy fence
  // sits right after the accumulator wait, where no global store of this thread is outstanding
  // any more (fence.proxy.async waits for the thread's pending memory operations).
  const uint32_t tm = tmem_slot_after_sync(&tmem_slot);
  const int total = tiles * B, stepg = gridDim.x;
  auto issue = [&](uint32_t acc) {   // warp 0, converged: the UMMA chain of the staged tile
    const uint32_t idesc = make_idesc_16(S.npad, true, false, F16 ? 0u : 1u);
    const uint32_t sa = smem_u32(stg), sw = smem_u32(wsm);
    const uint32_t rsw = (uint32_t)(S.kdim / 8) * 128;   // weight image row-group stride
    const int ks = S.cpad / 16;
    uint32_t first = 1;
    for (int t = 0; t < S.taps; ++t) {
      const int dy = k == 3 ? t / 3 - 1 : 0, dxi = k == 3 ? t % 3 : 0;
      const uint32_t a0 = sa + dxi * S.s_bytes + (uint32_t)((halo + dy * wch) * cg) * 128;
      const uint32_t b0 = sw + (uint32_t)(t * S.cpad / 8) * 128;
      for (int s = 0; s < ks; ++s) {
        const uint64_t ad = make_desc(a0 + s * 256, 128, (uint32_t)cg * 128);   // LBO: next 8 channels, SBO: next 8 pixels
        const uint64_t bd = make_desc(b0 + s * 256, 128, rsw);
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.eq.b32 p, %4, 0;\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(acc), "l"(ad), "l"(bd), "r"(idesc), "r"(first)
            : "memory");
        first = 0;
      }
    }
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(&bar))
        : "memory");
  };
  auto stage_and_issue = [&](uint32_t acc) {   // registers -> staged copies -> UMMAs into accumulator acc
    store_tile();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) issue(acc);
  };
  uint32_t phase = 0, n = 0;
  int g = blockIdx.x;
  if (g < total) {
    load_tile(g);
    stage_and_issue(tm);
    if (g + stepg < total) load_tile(g + stepg);
  }
  for (; g < total; g += stepg, ++n) {
    const int b = g / tiles, p0 = (g % tiles) * kCtTM;
    const uint32_t acc = dbl ? tm + (n & 1) * S.npad : tm;
    mbar_wait(&bar, phase);   // tile n: accumulator ready, staged copies free
    phase ^= 1;
    tc_fence_after();
    if (dbl && g + stepg < total) {
      stage_and_issue(tm + ((n + 1) & 1) * S.npad);   // tile n+1 (v[] holds its data)
      if (g + 2 * stepg < total) load_tile(g + 2 * stepg);
    }
    // ---- epilogue of tile n: lane == pixel, column == output channel ----
    {
      const int q4 = warp & 3, hw = warp >> 2;   // lane quarter; which 16-column groups (even / odd)
      const int p = p0 + q4 * 32 + lane;
      const bool live = p < P;
      for (int n0 = hw * 16; n0 < S.npad; n0 += 16 * (kCtThreads / 128)) {
        float a16[16];
        tmem_ld16(acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)n0, a16);
        float *yp = y + (size_t)b * y_bs + (size_t)n0 * P + p;
        const float *rp = residual ? residual + (size_t)b * r_bs + (size_t)n0 * P + p : nullptr;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (n0 + e < cout && live) {
            float o = a16[e] + sbias[n0 + e];
            if (rp) o += __ldg(rp);
            if (relu) o = fmaxf(o, 0.f);
            *yp = o;
          }
          yp += P;
          if (rp) rp += P;
        }
      }
    }
    tc_fence_before();   // this accumulator is rewritten by the UMMAs issued after the next __syncthreads
    if (!dbl && g + stepg < total) {
      stage_and_issue(tm);
      if (g + 2 * stepg < total) load_tile(g + 2 * stepg);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, ncols);
}

// weight (cout, cin, k, k) fp32 -> image[n][tap*cpad + c] (K-major core-matrix layout), zero padded
template <bool F16>
__global__ void conv_tc_pack_kernel(const float *__restrict__ w, int cin, int cout, int k, uint8_t *__restrict__ img) {
  const ConvTcShape S = conv_tc_shape(cin, cout, k, 8);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S.npad * S.kdim) return;
  const int n = e / S.kdim, kk = e % S.kdim, t = kk / S.cpad, c = kk % S.cpad;
  const float v = (n < cout && c < cin) ? w[((size_t)n * cin + c) * S.taps + t] : 0.f;
  *reinterpret_cast<unsigned short *>(img + cm_offset(n, kk, (S.kdim / 8) * 128, 128)) = to16<F16>(v);
}

bool conv_tc_supported(int cin, int cout, int h, int w, int k) {
  if ((k != 1 && k != 3) || w % 8 != 0 || cin > 96 || cout > 96) return false;
  const ConvTcShape s = conv_tc_shape(cin, cout, k, w);
  // every staged item of a tile lives in the registers of one thread slot (<= 6 per thread)
  return s.smem <= 100 * 1024 && ceil_div(s.cpad * (int)s.s_chunks, kCtMaxThreads) <= 6;
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_conv2d_tc_supported(int cin, int cout, int h, int w, int k) {
  return (cin > 0 && cout > 0 && h > 0 && w > 0 && conv_tc_supported(cin, cout, h, w, k)) ? 1 : 0;
}

extern "C" size_t otp_conv2d_tc_pack_bytes(int cin, int cout, int k) {
  if (cin <= 0 || cout <= 0 || (k != 1 && k != 3)) return 0;
  return conv_tc_shape(cin, cout, k, 8).w_bytes;
}

extern "C" int otp_conv2d_tc_pack(const float *weight, int cin, int cout, int k, int precision, void *packed,
                                  size_t packed_bytes, otp_stream_t stream) {
  OTP_REQUIRE(weight && packed && cin > 0 && cout > 0 && (k == 1 || k == 3));
  OTP_REQUIRE(precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  const ConvTcShape S = conv_tc_shape(cin, cout, k, 8);
  if (packed_bytes < S.w_bytes) {
    set_error("otp_conv2d_tc_pack: buffer of %zu B, need %u B", packed_bytes, S.w_bytes);
    return OTP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PACK, st);
  const int n = S.npad * S.kdim;
  if (precision == OTP_PREC_FP16)
    conv_tc_pack_kernel<true><<<ceil_div(n, 256), 256, 0, st>>>(weight, cin, cout, k, static_cast<uint8_t *>(packed));
  else
    conv_tc_pack_kernel<false><<<ceil_div(n, 256), 256, 0, st>>>(weight, cin, cout, k, static_cast<uint8_t *>(packed));
  return check_launch("conv_tc_pack_kernel");
}

extern "C" int otp_conv2d_tc(const float *x, long long x_bstride, const float *x_add, long long x_add_bstride,
                             const void *packed, const float *bias, const float *residual,
                             long long residual_bstride, float *y, long long y_bstride, int b, int cin, int h, int w,
                             int cout, int k, int relu, int precision, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && cin > 0 && h > 0 && w > 0 && cout > 0 && b <= 65535);
  OTP_REQUIRE(precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  if (!conv_tc_supported(cin, cout, h, w, k)) {
    set_error("otp_conv2d_tc: shape cin=%d cout=%d k=%d w=%d not built (k in {1,3}, w %% 8 == 0, channels <= 96)", cin,
              cout, k, w);
    return OTP_ERR_UNSUPPORTED;
  }
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(x && packed && bias && y);
  auto al16 = [](const void *q, long long bs) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && bs % 4 == 0; };
  OTP_REQUIRE(al16(x, x_bstride) && (!x_add || al16(x_add, x_add_bstride)));
  const ConvTcShape S = conv_tc_shape(cin, cout, k, w);
  const int tiles = ceil_div(h * w, kCtTM);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr;
  if (attr.first()) {
    const size_t lim = 100 * 1024;
    bool ok = true;
#define OTP_CONV_ATTR(U, M, NT)                                                        \
  ok = ok && set_max_smem(conv_tc_kernel<true, U, M, NT>, lim, "conv_tc_kernel") &&    \
       set_max_smem(conv_tc_kernel<false, U, M, NT>, lim, "conv_tc_kernel")
    OTP_CONV_ATTR(3, 6, 128);
    OTP_CONV_ATTR(5, 4, 128);
    OTP_CONV_ATTR(5, 2, 256);
    OTP_CONV_ATTR(6, 2, 256);
#undef OTP_CONV_ATTR
    if (!ok) return OTP_ERR_CUDA;
  }
  LaunchScope ls(K_CONV2D, st);
  const uint8_t *pk = static_cast<const uint8_t *>(packed);
  const int total = tiles * b;
#define OTP_CONV_TC(F16, U, MINCTA, NT)                                                                          \
  conv_tc_kernel<F16, U, MINCTA, NT><<<min(total, num_sms() * min(MINCTA, max_cta)), NT, S.smem, st>>>(               \
      x, x_bstride, x_add, x_add_bstride, pk, bias, residual, residual_bstride, y, y_bstride, b, cin, h, w, cout, k, \
      relu, tiles)
  // resident CTAs that shared memory and the 512 TMEM columns (two accumulators per CTA) allow
  const int tcols = S.npad <= 16 ? 32 : (S.npad <= 32 ? 64 : 128);
  const int max_cta = max(1, min((int)((227 * 1024) / (S.smem + 2048)), 512 / tcols));
  // small convs (<= 5 staged items per thread at 128 threads): 128-thread CTAs, 4-6 per SM;
  // larger ones: 256-thread CTAs (half the per-tile latency chain), 2 per SM
  const int items = S.cpad * (int)S.s_chunks;
  const bool f16 = precision == OTP_PREC_FP16;
  if (items <= 3 * 128) {
    if (f16) OTP_CONV_TC(true, 3, 6, 128); else OTP_CONV_TC(false, 3, 6, 128);
  } else if (items <= 5 * 128) {
    if (f16) OTP_CONV_TC(true, 5, 4, 128); else OTP_CONV_TC(false, 5, 4, 128);
  } else if (items <= 5 * 256) {
    if (f16) OTP_CONV_TC(true, 5, 2, 256); else OTP_CONV_TC(false, 5, 2, 256);
  } else {
    if (f16) OTP_CONV_TC(true, 6, 2, 256); else OTP_CONV_TC(false, 6, 2, 256);
  }
#undef OTP_CONV_TC
  return check_launch("conv_tc_kernel");
}
