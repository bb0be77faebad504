// a8 + a9 + a10 fused (16-bit tensor-core path): dilated 3x3 offset conv (32 -> 306) and
// mask conv (32 -> 153) as one implicit GEMM on tcgen05, consumed straight out of TMEM
// by the modulated deformable convolution and the 0.2-weighted accumulation.
//
// Reference: model/OTPose.py:381-392 runs, per dilation, two cuDNN convs that write
// offsets (B,306,H,W) and masks (B,153,H,W) to HBM -- 63 MB fp32 per clip over the five
// dilations -- and the DCN op (thirdparty/deform_conv/src/deform_conv_cuda.cpp:474-549)
// reads them back per sample.  Here those tensors never exist: a tile of 128 pixels
// forms its im2col operand in shared memory (9 taps x [128 px][32 ch]), one UMMA chain
// (M128 x N256 x K288) leaves the tile's offsets and masks in TMEM with pixel == TMEM
// lane, and each thread reads its own pixel's 27 values per joint with tcgen05.ld,
// samples def_heatmaps bilinearly (deform_conv_cuda_kernel.cu:402-432, 505-571 semantics)
// and contracts with the 17x153 DCN weight.
//
// The 17 joints (deformable groups) are split 9 + 8 over two launches so that one
// launch's conv weights (9 taps x [256][32] x 2 B = 144 KB) stay resident in shared
// memory for a persistent CTA; the second launch adds its partial sums to the first
// (stream-ordered, deterministic).
#include "common.cuh"
#include "tc_common.cuh"

namespace otp {
using namespace tc;

namespace {
constexpr int kJ = 17, kCin = 32, kTaps = 9;
constexpr int kTM = 128, kThreads = 384;   // thread = (pixel, third): 3 taps / 3 joints each
constexpr int kNPad = 256;                        // UMMA N (>= 28 * 9 = 252)
constexpr int kJRows = 28;                        // accumulator columns per joint
constexpr uint32_t kCS = 128, kRS32 = (kCin / 8) * 128;   // 512
constexpr uint32_t kATap = (kTM / 8) * kRS32;     // 8192   [128 px][32 ch]
constexpr uint32_t kWTap = (kNPad / 8) * kRS32;   // 16384  [256 rows][32 ch]
constexpr uint32_t kWSlice = kTaps * kWTap;       // 147456 one joint-half
constexpr uint32_t kASlice = kTaps * kATap;       // 73728
constexpr int kJ0[2] = {0, 9}, kNJ[2] = {9, 8};
constexpr size_t kPackBytes = (size_t)2 * 2 * kWSlice;   // [format][joint half]
constexpr int kWdLd = 20;                         // DCN weight row: 17 outputs padded to 5 float4
constexpr size_t kSmem = (size_t)kWSlice + kASlice + (size_t)81 * kWdLd * 4 + 64;
static_assert(kSmem + 1024 <= 227 * 1024, "fused DCN shared memory");

__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float &a, float &b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, float &a) {
  uint32_t r0;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = __uint_as_float(r0);
}

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
tc_dcn_fused_kernel(const uint8_t *__restrict__ wimg, const float *__restrict__ trans,
                    const float *__restrict__ x, const float *__restrict__ dcn_w,
                    const float *__restrict__ dcn_b, float *__restrict__ out, int B, int H, int W, int dil,
                    int j0, int nj, float alpha, int accumulate, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *ws = smem;                         // 9 x [256][32] conv weights, resident
  uint8_t *as = smem + kWSlice;               // 9 x [128][32] im2col taps; later the partial-sum exchange
  float *wd = reinterpret_cast<float *>(as + kASlice);   // [nj*9][20] DCN weight slice, output channel fastest
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q4 = warp & 3, third = warp >> 2, tok = q4 * 32 + lane;
  const int P = H * W;

  for (uint32_t o = threadIdx.x * 16; o < kWSlice; o += kThreads * 16) cp_async16(ws + o, wimg + o);
  cp_async_commit();
  for (int e = threadIdx.x; e < nj * 9 * kWdLd; e += kThreads) {
    const int r = e / kWdLd, o = e % kWdLd;
    wd[e] = o < kJ ? dcn_w[(size_t)o * kJ * 9 + j0 * 9 + r] : 0.f;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t idesc = make_idesc_16(kNPad, false, false, F16 ? 0u : 1u);
  uint32_t ph = 0;
  // joints of this thread in the DCN phase: the CTA's nj joints split over the three thread thirds
  const int jl_lo = 3 * third, jl_hi = min(nj, 3 * third + 3);

  for (int g = blockIdx.x; g < B * tiles; g += gridDim.x) {
    const int b = g / tiles, tile = g % tiles;
    const int p = tile * kTM + tok;
    const bool live = p < P;
    const int h = live ? p / W : 0, w = live ? p % W : 0;
    // ---- im2col: this thread's 3 of the 9 dilated taps, all 32 channels (96 loads in flight) ----
    {
      const float *tb = trans + (size_t)b * kCin * P;
      float v[3][kCin];
#pragma unroll
      for (int tt = 0; tt < 3; ++tt) {
        const int t = 3 * third + tt;
        const int hh = h + (t / 3 - 1) * dil, ww = w + (t % 3 - 1) * dil;
        const bool ok = live && hh >= 0 && hh < H && ww >= 0 && ww < W;
        const float *src = tb + (ok ? hh * W + ww : 0);
        if (ok) {      // one branch per tap instead of a select per channel; channel c is c * P elements further
#pragma unroll
          for (int c = 0; c < kCin; ++c) v[tt][c] = __ldg(src + (size_t)c * P);
        } else {
#pragma unroll
          for (int c = 0; c < kCin; ++c) v[tt][c] = 0.f;
        }
      }
#pragma unroll
      for (int tt = 0; tt < 3; ++tt) {
        uint8_t *dst = as + (3 * third + tt) * kATap + cm_offset(tok, 0, kRS32, kCS);
#pragma unroll
        for (int g = 0; g < kCin / 8; ++g) *reinterpret_cast<uint4 *>(dst + g * kCS) = pack16x8<F16>(v[tt] + 8 * g);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(as), w0 = smem_u32(ws);
#pragma unroll
      for (int t = 0; t < kTaps; ++t)
#pragma unroll
        for (int s = 0; s < kCin / 16; ++s)
          umma_bf16(tm, make_desc(a0 + t * kATap + s * 2 * kCS, kCS, kRS32),
                    make_desc(w0 + t * kWTap + s * 2 * kCS, kCS, kRS32), idesc, (t > 0 || s > 0));
      umma_commit(&bar);
    }
    mbar_wait(&bar, ph);
    ph ^= 1;
    tc_fence_after();
    // ---- modulated deformable sampling + 17 x (nj*9) contraction, offsets/masks from TMEM ----
    float acc[kJ];
#pragma unroll
    for (int o = 0; o < kJ; ++o) acc[o] = 0.f;
    const uint32_t trow = tm + ((uint32_t)(q4 * 32) << 16);
    const float *xb = x + (size_t)b * kJ * P;
#pragma unroll 1
    for (int jl = jl_lo; jl < jl_hi; ++jl) {
      float off[18], msk[9];
      {
        // 28 consecutive columns of this joint: x16 + x8 + x4 loads, one wait
        uint32_t r0[16], r1[8], r2[4];
        const uint32_t ta = trow + kJRows * jl;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]),
              "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
            : "r"(ta)
            : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7])
                     : "r"(ta + 16)
                     : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3])
                     : "r"(ta + 24)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("" : "+r"(r0[i])::"memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("" : "+r"(r1[i])::"memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) asm volatile("" : "+r"(r2[i])::"memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) off[i] = __uint_as_float(r0[i]);
        off[16] = __uint_as_float(r1[0]);
        off[17] = __uint_as_float(r1[1]);
#pragma unroll
        for (int i = 0; i < 6; ++i) msk[i] = __uint_as_float(r1[2 + i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) msk[6 + i] = __uint_as_float(r2[i]);
      }
      if (live) {
        const float *img = xb + (size_t)(j0 + jl) * P;
        // Branch-free bilinear taps (same arithmetic as dmcn_im2col_bilinear): out-of-range corners
        // get weight 0 and a clamped address, so the 12 loads of a tap triple are independent
        // and in flight together instead of sitting behind data-dependent branches.  Two rounds
        // (taps 0-4, 5-8): 20 + 16 gathers in flight.
#pragma unroll
        for (int tg = 0; tg < 2; ++tg) {
          float wgt[5][4];
          int adr[5][4];
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int t = 5 * tg + u;
            if (t >= kTaps) continue;
            const float h_im = (float)(h + (t / 3 - 1) * dil) + off[2 * t];
            const float w_im = (float)(w + (t % 3 - 1) * dil) + off[2 * t + 1];
            const bool in = h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W;
            const float hf = floorf(fminf(fmaxf(h_im, -2.f), (float)H + 1.f));
            const float wf = floorf(fminf(fmaxf(w_im, -2.f), (float)W + 1.f));
            const int h_low = (int)hf, w_low = (int)wf;
            const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
            const bool h0 = in && h_low >= 0, h1 = in && h_low + 1 <= H - 1;
            const bool w0 = w_low >= 0, w1 = w_low + 1 <= W - 1;
            wgt[u][0] = (h0 && w0) ? hh * hw : 0.f;
            wgt[u][1] = (h0 && w1) ? hh * lw : 0.f;
            wgt[u][2] = (h1 && w0) ? lh * hw : 0.f;
            wgt[u][3] = (h1 && w1) ? lh * lw : 0.f;
            const int hc0 = min(max(h_low, 0), H - 1), hc1 = min(max(h_low + 1, 0), H - 1);
            const int wc0 = min(max(w_low, 0), W - 1), wc1 = min(max(w_low + 1, 0), W - 1);
            adr[u][0] = hc0 * W + wc0;
            adr[u][1] = hc0 * W + wc1;
            adr[u][2] = hc1 * W + wc0;
            adr[u][3] = hc1 * W + wc1;
          }
          float val[5][4];
#pragma unroll
          for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (5 * tg + u < kTaps) val[u][q] = __ldg(img + adr[u][q]);
#pragma unroll
          for (int u = 0; u < 5; ++u) {
            const int t = 5 * tg + u;
            if (t >= kTaps) continue;
            const float v = wgt[u][0] * val[u][0] + wgt[u][1] * val[u][1] + wgt[u][2] * val[u][2] +
                            wgt[u][3] * val[u][3];
            const float col = v * msk[t];
            const float4 *wr = reinterpret_cast<const float4 *>(wd + (jl * 9 + t) * kWdLd);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w4 = wr[q];
              acc[4 * q] = fmaf(w4.x, col, acc[4 * q]);
              acc[4 * q + 1] = fmaf(w4.y, col, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(w4.z, col, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(w4.w, col, acc[4 * q + 3]);
            }
            acc[16] = fmaf(wr[4].x, col, acc[16]);
          }
        }
      }
    }
    // ---- combine the three thread thirds (exchange buffer aliases the dead im2col taps): every
    //      thread publishes its 17 partial sums, then third q reduces and stores outputs q, q+3, ... ----
    float *ex = reinterpret_cast<float *>(as);
    // the running sums this thread adds to (launches 2..10 of a forward): all six loads in flight across the
    // exchange barrier -- read one by one next to their stores they were six serialised round trips per tile
    // (17 % of the kernel's stall samples)
    float prev[(kJ + 2) / 3];
#pragma unroll
    for (int i = 0; i < (kJ + 2) / 3; ++i) {
      const int o = third + 3 * i;
      prev[i] = (live && accumulate && o < kJ) ? out[((size_t)b * kJ + o) * P + p] : 0.f;
    }
#pragma unroll
    for (int o = 0; o < kJ; ++o) ex[(third * kJ + o) * kTM + tok] = acc[o];
    tc_fence_before();
    __syncthreads();
    if (live) {
#pragma unroll
      for (int i = 0; i < (kJ + 2) / 3; ++i) {
        const int o = third + 3 * i;
        if (o < kJ) {
          const float v = ex[o * kTM + tok] + ex[(kJ + o) * kTM + tok] + ex[(2 * kJ + o) * kTM + tok] +
                          (dcn_b ? __ldg(dcn_b + o) : 0.f);
          out[((size_t)b * kJ + o) * P + p] = fmaf(alpha, v, prev[i]);
        }
      }
    }
    __syncthreads();   // exchange buffer / TMEM are rewritten by the next tile
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// image row r of joint-half hf: 28 rows per joint (18 offsets, 9 masks, one zero row), so that a
// thread reads a joint's 27 values with one batch of tcgen05.ld and a single wait
template <bool F16>
__global__ void pack_offset_mask_kernel(const float *__restrict__ w_off, const float *__restrict__ w_msk,
                                        uint8_t *__restrict__ dst, int j0, int nj) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kTaps * kNPad * kCin) return;
  const int c = e % kCin, r = (e / kCin) % kNPad, t = e / (kCin * kNPad);
  float v = 0.f;
  const int jl = r / kJRows, i = r % kJRows;   // per joint: 18 offset rows, 9 mask rows, 1 zero row
  if (jl < nj && i < 18) v = w_off[((size_t)((j0 + jl) * 18 + i) * kCin + c) * 9 + t];
  else if (jl < nj && i < 27) v = w_msk[((size_t)((j0 + jl) * 9 + i - 18) * kCin + c) * 9 + t];
  *reinterpret_cast<unsigned short *>(dst + (size_t)t * kWTap + cm_offset(r, c, kRS32, kCS)) = to16<F16>(v);
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" size_t otp_offset_mask_pack_bytes(void) { return kPackBytes; }

extern "C" int otp_offset_mask_pack(const float *w_off, const float *w_msk, int joints, int cin, void *packed,
                                    size_t packed_bytes, otp_stream_t stream) {
  if (joints != kJ || cin != kCin) {
    set_error("fused offset/mask/DCN kernel is built for 17 joints x 32 channels (got %d x %d)", joints, cin);
    return OTP_ERR_UNSUPPORTED;
  }
  OTP_REQUIRE(w_off && w_msk && packed);
  if (packed_bytes < kPackBytes) {
    set_error("otp_offset_mask_pack: buffer of %zu B, need %zu B", packed_bytes, kPackBytes);
    return OTP_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PACK, st, 4);
  const int n = kTaps * kNPad * kCin;
  for (int hf = 0; hf < 2; ++hf) {
    uint8_t *d = static_cast<uint8_t *>(packed);
    pack_offset_mask_kernel<false><<<ceil_div(n, 256), 256, 0, st>>>(w_off, w_msk, d + (0 * 2 + hf) * (size_t)kWSlice,
                                                                      kJ0[hf], kNJ[hf]);
    pack_offset_mask_kernel<true><<<ceil_div(n, 256), 256, 0, st>>>(w_off, w_msk, d + (1 * 2 + hf) * (size_t)kWSlice,
                                                                     kJ0[hf], kNJ[hf]);
  }
  return check_launch("pack_offset_mask_kernel");
}

extern "C" int otp_offset_mask_dcn_forward(const void *packed, const float *trans, const float *x,
                                           const float *dcn_w, const float *dcn_b, float *out, int b, int h,
                                           int w, int dilation, float alpha, int accumulate, int precision,
                                           otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && h > 0 && w > 0 && dilation > 0);
  OTP_REQUIRE(precision == OTP_PREC_BF16 || precision == OTP_PREC_FP16);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(packed && trans && x && dcn_w && out);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f16 = precision == OTP_PREC_FP16;
  static PerDeviceOnce attr;
  if (attr.first()) {
    if (!set_max_smem(tc_dcn_fused_kernel<true>, kSmem, "tc_dcn_fused_kernel") ||
        !set_max_smem(tc_dcn_fused_kernel<false>, kSmem, "tc_dcn_fused_kernel"))
      return OTP_ERR_CUDA;
  }
  const int tiles = ceil_div(h * w, kTM);
  const int grid = b * tiles < num_sms() ? b * tiles : num_sms();
  for (int hf = 0; hf < 2; ++hf) {
    const uint8_t *wimg = static_cast<const uint8_t *>(packed) + ((f16 ? 2 : 0) + hf) * (size_t)kWSlice;
    LaunchScope ls(K_TC_CONV, st);
    if (f16)
      tc_dcn_fused_kernel<true><<<grid, kThreads, kSmem, st>>>(wimg, trans, x, dcn_w, hf == 0 ? dcn_b : nullptr, out,
                                                              b, h, w, dilation, kJ0[hf], kNJ[hf], alpha,
                                                              hf == 0 ? accumulate : 1, tiles);
    else
      tc_dcn_fused_kernel<false><<<grid, kThreads, kSmem, st>>>(wimg, trans, x, dcn_w, hf == 0 ? dcn_b : nullptr,
                                                               out, b, h, w, dilation, kJ0[hf], kNJ[hf], alpha,
                                                               hf == 0 ? accumulate : 1, tiles);
  }
  return check_launch("tc_dcn_fused_kernel");
}
