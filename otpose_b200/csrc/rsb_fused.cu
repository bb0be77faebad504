// a7 (16-bit modes): an RSB_BLOCK (reference model/RSB.py:26-103) as NINE lean launches -- one per depth level of
// its conv DAG -- on 16-bit pixel-major intermediates, instead of 13 launches of the general conv kernel that
// round-trip fp32 NCHW maps (and re-stage / re-convert them on every read).
//
//   spx = relu(bn(conv1x1(x)))  -> 4 slices of bc channels                                   level 0
//   ten dense-connected 3x3 convs on bc channels:  {1_1} {2_1} {2_2, 3_1} {3_2, 4_1} {3_3, 4_2} {4_3} {4_4}   levels 1..7
//   y = relu(bn(conv1x1(cat(out_1_1, out_2_2, out_3_3, out_4_4))) + (x | bn(conv1x1(x))))    level 8
//
// Every intermediate map is a "plane": [B][H][W][cp] IEEE halves, cp = bc padded to 8 (L2 resident between the
// launches; 48 bytes per pixel at bc = 20 against 80 in fp32 NCHW).  A pixel's channels are then one ldmatrix
// row, so a launch's input tile -- TH image rows (+ one halo row either side for the 3x3 levels), full width --
// is a straight cp.async copy into shared memory and the A operand of tap (dy, dx) is read IN PLACE from the
// shifted tile rows: implicit GEMM per (op, row, 48-pixel group) on mma.sync m16n8k16 / m16n8k8 (half operands,
// fp32 accumulate), the B operand pre-packed in fragment order.  Summed inputs (spx[i] + out) are added as half
// pairs in the A fragments; ops of one level that share an input (2_2 and 3_1 read out_2_1) share its tile.
// BatchNorm is folded (otp_conv_bn_fold arithmetic) when the block is packed; the downsample conv is folded into
// level 8 along K.
//
// Why warp-level MMA and not tcgen05: the GEMMs are 6..20 channels wide (N = 8..24, K = 8..24 per tap).  The tensor
// pipe is < 10 % busy either way; what the per-conv tcgen05 kernel (conv_tc.cu, ~25 us per launch) pays for is
// the fp32 NCHW <-> operand-layout conversion, three pre-shifted copies, TMEM allocation and an mbarrier round
// trip per 128 pixels.  Here a launch is: copy tile, one barrier, MMAs, 16-bit stores.
//
// (A single-launch streaming variant -- row rings in shared memory, all 13 convs per step -- was built first and
// measured 2-3x SLOWER than the per-conv path: one row per step leaves 12 warps ~10 % issue-busy between barriers.
// profiles/r02_experiments.md.)
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace otp {
namespace {
using tc::pack16x2;

constexpr int RTH = 256, RNW = RTH / 32;   // threads / warps per CTA
constexpr int RMAXT = 6, RMAXOP = 4, RMAXIN = 5;
constexpr int RMG = 3;                     // m-tiles (16 pixels) per warp task
constexpr size_t RSMEM = 100 * 1024;       // two CTAs per SM

struct RsbTensor {
  int off;      // byte offset of the tile in dynamic shared memory
  int ch;       // channels (multiple of 8, zero padded)
  int sp;       // halves per pixel in shared memory (>= ch; sp / 8 odd where it is free: conflict-free ldmatrix)
  int plane;    // source: 16-bit plane index, -1 = the block input x (fp32 NCHW), -2 = its 16-bit copy (xplane)
};
struct RsbOp {
  int nin, in[RMAXIN];   // input tensors: 3x3: summed (<= 2);  1x1: concatenated along K
  int k;                 // 1 | 3
  int nt, kbw;           // output n8 tiles; k-blocks of 8 per tap
  int wg, wn, bg, bn;    // weight words / biases of this op in the packed unit image (global offsets, counts)
  int wofs, bofs;        // ... and where they sit in shared memory (word / float offsets)
  int relu;
  int plane;             // output plane (-1: the block output y, fp32 NCHW, + identity residual, ReLU)
  int gch;               // real output channels of y
};
struct RsbArgs {
  RsbTensor t[RMAXT];
  RsbOp op[RMAXOP];
  int ntens, nop, halo;    // halo = 1: 3x3 level (tiles carry one row above and below)
  int H, W, TH, tiles_y, MT, BWP, rows;   // rows = TH + 2 * halo tile rows
  int tile_off, tile_bytes;
  const uint32_t *wunits;  // packed unit image
  const float *bunits;
  const float *x;          // (B, cin, H, W) fp32
  long long x_bs;
  int cin;
  float *y;                // (B, planes, H, W) fp32
  long long y_bs;
  const float *res;        // identity skip (NULL: the downsample conv is part of level 8)
  long long res_bs;
  __half *planes;          // [plane][B][H][W][cp]
  __half *xplane;          // [B][H][W][cinp]: 16-bit copy of x written by level 0, read by level 8 (downsample)
  int cp, B, cinp, write_x;
};

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_k16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ uint32_t hadd2u(uint32_t a, uint32_t b) {
  const __half2 r = __hadd2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
  return *reinterpret_cast<const uint32_t *>(&r);
}

// Epilogue of one accumulator tile set (image row r of clip b): + bias, ReLU -> 16-bit plane, or
// + identity residual, ReLU -> the block's fp32 NCHW output.
template <int NT, int MG>
__device__ __forceinline__ void conv_epilogue(const RsbArgs &A, const RsbOp &op, const uint8_t *sm,
                                              const float (&acc)[MG][NT][4], int b, int r, int m0) {
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const float *bias = reinterpret_cast<const float *>(sm) + op.bofs + 2 * q;
  float2 bs[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) bs[n] = *reinterpret_cast<const float2 *>(bias + 8 * n);
  if (op.plane >= 0) {
    // ---- 16-bit plane: one base pointer per (m-tile, row half), the n-tiles are 16-byte steps from it ----
    __half *prow = A.planes + ((((size_t)op.plane * A.B + b) * A.H + r) * A.W) * A.cp + 2 * q;
    const bool relu = op.relu != 0;
#pragma unroll
    for (int m = 0; m < MG; ++m)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int p = 16 * (m0 + m) + g + 8 * hh;
        if (p >= A.W) continue;
        uint32_t *dst = reinterpret_cast<uint32_t *>(prow + (size_t)p * A.cp);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          float v0 = acc[m][n][2 * hh] + bs[n].x, v1 = acc[m][n][2 * hh + 1] + bs[n].y;
          if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
          dst[4 * n] = pack16x2<true>(v0, v1);
        }
      }
    return;
  }
  // ---- the block's fp32 NCHW output: + identity skip, ReLU ----
  const size_t P = (size_t)A.H * A.W;
  float *yb = A.y + (size_t)b * A.y_bs + (size_t)r * A.W + (size_t)(2 * q) * P;
  const float *rb = A.res ? A.res + (size_t)b * A.res_bs + (size_t)r * A.W + (size_t)(2 * q) * P : nullptr;
#pragma unroll
  for (int m = 0; m < MG; ++m) {
    float res[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)   // every skip load of the tile in flight before the first store
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = 16 * (m0 + m) + g + 8 * (i >> 1), ch = 8 * n + 2 * q + (i & 1);
        res[n][i] = (rb && p < A.W && ch < op.gch) ? __ldg(rb + (size_t)(8 * n + (i & 1)) * P + p) : 0.f;
      }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int ch = 8 * n + 2 * q;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int p = 16 * (m0 + m) + g + 8 * hh;
        if (p >= A.W) continue;
        float *yp = yb + (size_t)(8 * n) * P + p;
        if (ch < op.gch) yp[0] = fmaxf(acc[m][n][2 * hh] + bs[n].x + res[n][2 * hh], 0.f);
        if (ch + 1 < op.gch) yp[P] = fmaxf(acc[m][n][2 * hh + 1] + bs[n].y + res[n][2 * hh + 1], 0.f);
      }
    }
  }
}

// One (3x3 conv, group of <= RMG m-tiles) of tile row j (image row r).  CPB = branch channels / 8 (compile time:
// the k16 / k8 split, the n-tiles and every shared-memory offset are immediates); TWO: the input is a sum.
// MG = m-tiles of the group (compile time: an mma.sync under a data-dependent predicate costs a WARPSYNC + NOP each).
template <int CPB, bool TWO, int MG>
__device__ __forceinline__ void conv3_task(const RsbArgs &A, const RsbOp &op, uint8_t *sm, int b, int r, int j, int m0) {
  constexpr int SP = 8 * CPB, NT = CPB;
  const int lane = threadIdx.x & 31;
  const uint32_t *wf = reinterpret_cast<const uint32_t *>(sm) + op.wofs + lane;
  const uint32_t rowb = (uint32_t)(A.BWP * SP * 2);
  // tile row j - 1 of each input + this lane's ldmatrix row (pixel lane & 15, channel block lane >> 4)
  const uint32_t lofs = (uint32_t)(j - 1) * rowb + (uint32_t)(((16 * m0 + 1 + (lane & 15)) * SP + 8 * (lane >> 4)) * 2);
  const uint32_t base0 = tc::smem_u32(sm) + A.t[op.in[0]].off + lofs;
  const uint32_t base1 = TWO ? tc::smem_u32(sm) + A.t[op.in[1]].off + lofs : 0u;
  float acc[MG][NT][4];
#pragma unroll
  for (int m = 0; m < MG; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  // A fragment of (tap, 8-channel block kb) for m-tile m: k16 -> 4 registers, one 8-block -> 2 registers at a[half]
  auto tap_ofs = [&](int tap) { return (uint32_t)(tap / 3) * rowb + (uint32_t)((tap % 3 - 1) * SP * 2); };
  auto load16 = [&](uint32_t (&a)[4], int tap, int kb, int m) {
    const uint32_t o = tap_ofs(tap) + (uint32_t)((16 * m * SP + 8 * kb) * 2);
    ldsm4(a, base0 + o);
    if (TWO) {
      uint32_t a2[4];
      ldsm4(a2, base1 + o);
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = hadd2u(a[i], a2[i]);
    }
  };
  auto load8 = [&](uint32_t (&a)[4], int half, int tap, int kb, int m) {
    const uint32_t o = tap_ofs(tap) + (uint32_t)((16 * m * SP + 8 * kb) * 2);
    uint32_t t[4];
    ldsm2(t, base0 + o);
    if (TWO) {
      uint32_t a2[4];
      ldsm2(a2, base1 + o);
      t[0] = hadd2u(t[0], a2[0]);
      t[1] = hadd2u(t[1], a2[1]);
    }
    a[2 * half] = t[0], a[2 * half + 1] = t[1];
  };
  // full 16-channel blocks of every tap
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int kb = 0; kb + 1 < CPB; kb += 2) {
      uint32_t a[MG][4];
#pragma unroll
      for (int m = 0; m < MG; ++m) load16(a[m], tap, kb, m);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t *wp = wf + ((tap * NT + n) * CPB + kb) * 32;
        const uint32_t b0 = wp[0], b1 = wp[32];
#pragma unroll
        for (int m = 0; m < MG; ++m) mma_k16(acc[m][n], a[m], b0, b1);
      }
    }
  }
  // odd 8-channel tail (cp = 8, 24): the tails of two taps make ONE k16 MMA (the tensor pipe, not the issue slots,
  // bounds these kernels: an m16n8k8 costs the pipe as much as an m16n8k16)
  if constexpr (CPB % 2 == 1) {
    constexpr int kb = CPB - 1;
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
      uint32_t a[MG][4];
#pragma unroll
      for (int m = 0; m < MG; ++m) {
        load8(a[m], 0, 2 * tp, kb, m);
        load8(a[m], 1, 2 * tp + 1, kb, m);
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t b0 = wf[(((2 * tp) * NT + n) * CPB + kb) * 32], b1 = wf[(((2 * tp + 1) * NT + n) * CPB + kb) * 32];
#pragma unroll
        for (int m = 0; m < MG; ++m) mma_k16(acc[m][n], a[m], b0, b1);
      }
    }
    uint32_t a[MG][4];
#pragma unroll
    for (int m = 0; m < MG; ++m) load8(a[m], 0, 8, kb, m);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const uint32_t b0 = wf[((8 * NT + n) * CPB + kb) * 32];
#pragma unroll
      for (int m = 0; m < MG; ++m) mma_k8(acc[m][n], a[m], b0);
    }
  }
  conv_epilogue<NT, MG>(A, op, sm, acc, b, r, m0);
}

// One (1x1 conv over the K-concatenation of its input tensors, group of <= RMG m-tiles) of tile row j.
template <int NT, int MG>
__device__ __forceinline__ void conv1_task(const RsbArgs &A, const RsbOp &op, uint8_t *sm, int b, int r, int j, int m0) {
  const int lane = threadIdx.x & 31;
  const uint32_t *wf = reinterpret_cast<const uint32_t *>(sm) + op.wofs + lane;
  float acc[MG][NT][4];
#pragma unroll
  for (int m = 0; m < MG; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  int kbg = 0;   // k-block index into the weight fragments
  for (int ii = 0; ii < op.nin; ++ii) {
    const RsbTensor &T0 = A.t[op.in[ii]];
    const int sp = T0.sp, KB = T0.ch >> 3;
    const uint32_t row0 = tc::smem_u32(sm) + T0.off + (uint32_t)(j * A.BWP * sp * 2) +
                          (uint32_t)(((16 * m0 + 1 + (lane & 15)) * sp + 8 * (lane >> 4)) * 2);
    for (int kb = 0; kb < KB; kb += 2) {
      const bool k16 = kb + 1 < KB;
      uint32_t a[MG][4];
#pragma unroll
      for (int m = 0; m < MG; ++m) {
        {
          const uint32_t o = (uint32_t)((16 * m * sp + 8 * kb) * 2);
          if (k16) ldsm4(a[m], row0 + o);
          else ldsm2(a[m], row0 + o);
        }
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t *wp = wf + (n * op.kbw + kbg + kb) * 32;
        const uint32_t b0 = wp[0];
        if (k16) {
          const uint32_t b1 = wp[32];
#pragma unroll
          for (int m = 0; m < MG; ++m) mma_k16(acc[m][n], a[m], b0, b1);
        } else {
#pragma unroll
          for (int m = 0; m < MG; ++m) mma_k8(acc[m][n], a[m], b0);
        }
      }
    }
    kbg += KB;
  }
  conv_epilogue<NT, MG>(A, op, sm, acc, b, r, m0);
}

// One depth level of the block: CTA = (clip, TH image rows).  3x3 levels: CPB = cp / 8 of the block, NT1 = 0;
// 1x1 levels: CPB = 0, NT1 = n8 tiles of their ops (compile time: one accumulator shape per kernel).
template <int CPB, int NT1>
__global__ void __launch_bounds__(RTH, 2) rsb_level_kernel(const __grid_constant__ RsbArgs A) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / A.tiles_y, tr0 = (blockIdx.x % A.tiles_y) * A.TH, tr1 = min(A.H, tr0 + A.TH);
  const size_t P = (size_t)A.H * A.W;
  // ---- weights / biases of this level's ops and the 16-bit input tiles by cp.async; pad pixels and rows outside
  //      the image are zero (the 3x3 zero padding) ----
  for (int o = 0; o < A.nop; ++o) {
    const RsbOp &op = A.op[o];
    uint32_t *w = reinterpret_cast<uint32_t *>(sm) + op.wofs;
    for (int i = threadIdx.x * 4; i < op.wn; i += RTH * 4) tc::cp_async16(w + i, A.wunits + op.wg + i);
    float *bs = reinterpret_cast<float *>(sm) + op.bofs;
    for (int i = threadIdx.x; i < op.bn; i += RTH) bs[i] = __ldg(A.bunits + op.bg + i);
  }
  for (int t = 0; t < A.ntens; ++t) {
    const RsbTensor &T = A.t[t];
    uint8_t *tile = sm + T.off;
    const int rowb = A.BWP * T.sp * 2;
    if (T.plane != -1) {
      const int per_px = T.ch >> 3, per_row = A.W * per_px;   // 16-byte chunks (a plane pixel holds T.ch channels)
      const __half *src = T.plane >= 0 ? A.planes + (((size_t)T.plane * A.B + b) * A.H) * A.W * A.cp
                                       : A.xplane + ((size_t)b * A.H) * A.W * A.cinp;
      const uint32_t magic = (65536u + per_px - 1) / per_px;   // c / per_px for c < 672 (W <= 96 x 7 chunks)
      for (int j = warp; j < A.rows; j += RNW) {
        const int r = tr0 - A.halo + j;
        uint8_t *drow = tile + j * rowb + T.sp * 2;
        const bool in = r >= 0 && r < A.H;
        const uint4 *srow = reinterpret_cast<const uint4 *>(src + (size_t)(in ? r : 0) * A.W * T.ch);
        for (int c = lane; c < per_row; c += 32) {
          const int px = (int)(((uint32_t)c * magic) >> 16), cb = c - px * per_px;
          uint8_t *dst = drow + (px * T.sp + 8 * cb) * 2;
          if (in) tc::cp_async16(dst, srow + c);
          else *reinterpret_cast<uint4 *>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        // pad pixels: column 0 and columns W + 1 .. BWP - 1
        const int spc = T.sp >> 3, npad = (A.BWP - A.W) * spc;
        for (int c = lane; c < npad; c += 32) {
          const int cb = c % spc, k = c / spc, col = k == 0 ? 0 : A.W + k;
          *reinterpret_cast<uint4 *>(tile + j * rowb + (col * T.sp + 8 * cb) * 2) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    } else {
      // the block input x: fp32 NCHW -> 16-bit pixel-major.  Work item = (tile row, 16-pixel group, channel block);
      // 8 channels of a pixel become one 16-byte store; a half-warp reads 64 contiguous bytes per channel.
      // half-warp = (tile row, 16-pixel group); its channel blocks four at a time (32 loads in flight per thread)
      const int ncb = T.ch >> 3;
      for (int it = warp * 2 + (lane >> 4); it < A.rows * A.MT; it += RNW * 2) {
        const int j = it / A.MT, m = it - j * A.MT, r = tr0 - A.halo + j, p = 16 * m + (lane & 15);
        const bool ok = r >= 0 && r < A.H && p < A.W;
        const float *src = A.x + (size_t)b * A.x_bs + (size_t)(ok ? r : 0) * A.W + (ok ? p : 0);
        uint8_t *dst = tile + j * rowb + (p + 1) * T.sp * 2;
        __half *xdst = A.xplane + (((size_t)b * A.H + r) * A.W + p) * A.cinp;
        for (int cb0 = 0; cb0 < ncb; cb0 += 4) {
          float v[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c0 = 8 * (cb0 + u);
            const float *s8 = src + (size_t)c0 * P;      // one 64-bit product per channel block, then steps of P
            if (ok && c0 + 8 <= A.cin) {                 // a whole block of real channels: no per-channel test
#pragma unroll
              for (int e = 0; e < 8; ++e) v[u][e] = __ldg(s8 + (size_t)e * P);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[u][e] = (ok && c0 + e < A.cin) ? __ldg(s8 + (size_t)e * P) : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (cb0 + u >= ncb) continue;
            const uint4 h = tc::pack16x8<true>(v[u]);
            *reinterpret_cast<uint4 *>(dst + 16 * (cb0 + u)) = h;
            if (A.write_x && ok) *reinterpret_cast<uint4 *>(xdst + 8 * (cb0 + u)) = h;   // 16-bit copy of x for level 8
          }
        }
      }
      for (int i = threadIdx.x; i < A.rows * (T.sp >> 3) * 2; i += RTH) {   // pixels 0 and BWP - 1 (16 MT + 1)
        const int j = i / ((T.sp >> 3) * 2), c = i % ((T.sp >> 3) * 2), col = (c & 1) ? A.BWP - 1 : 0;
        *reinterpret_cast<uint4 *>(tile + j * rowb + (col * T.sp + 8 * (c >> 1)) * 2) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  tc::cp_async_commit();
  tc::cp_async_wait<0>();
  __syncthreads();
  // ---- tasks: (op, image row, group of <= RMG m-tiles), round-robin over the warps ----
  constexpr int GM = (CPB > 0 || NT1 < 3) ? RMG : 2;   // m-tiles per task (the wide 1x1 accumulators: 2)
  const int ngrp = (A.MT + GM - 1) / GM, nrow = tr1 - tr0;
  // task -> (op, row, group) with two float-reciprocal divisions (exact for these small integers)
  const float inv_grp = 1.0f / (float)ngrp, inv_row = 1.0f / (float)nrow;
  for (int task = warp; task < A.nop * nrow * ngrp; task += RNW) {
    const int q1 = (int)(((float)task + 0.5f) * inv_grp), oi = (int)(((float)q1 + 0.5f) * inv_row);
    const RsbOp &op = A.op[oi];
    const int jr = q1 - oi * nrow, m0 = (task - q1 * ngrp) * GM, mg = min(GM, A.MT - m0);
    const int r = tr0 + jr, j = jr + A.halo;
    if constexpr (CPB > 0) {
      if (op.nin == 2) {
        if (mg == 3) conv3_task<CPB, true, 3>(A, op, sm, b, r, j, m0);
        else if (mg == 2) conv3_task<CPB, true, 2>(A, op, sm, b, r, j, m0);
        else conv3_task<CPB, true, 1>(A, op, sm, b, r, j, m0);
      } else {
        if (mg == 3) conv3_task<CPB, false, 3>(A, op, sm, b, r, j, m0);
        else if (mg == 2) conv3_task<CPB, false, 2>(A, op, sm, b, r, j, m0);
        else conv3_task<CPB, false, 1>(A, op, sm, b, r, j, m0);
      }
    } else {
      if (GM == 3 && mg == 3) conv1_task<NT1, GM>(A, op, sm, b, r, j, m0);
      else if (mg == 2) conv1_task<NT1, 2>(A, op, sm, b, r, j, m0);
      else conv1_task<NT1, 1>(A, op, sm, b, r, j, m0);
    }
  }
}

// ------------------------------------------------------------------------------------------ packed weights
// Conv order inside an RSB_BLOCK: 0 conv_bn_relu1, 1..10 conv_bn_relu2_{1_1,2_1,2_2,3_1,3_2,3_3,4_1,4_2,4_3,4_4},
// 11 conv_bn_relu3, 12 downsample.  Weight "units" of the packed image (fragment order, see rsb_pack_kernel):
// 0..3 conv1 slice i, 4..13 the 3x3 convs, 14 conv3 (+ downsample along K).
constexpr int kUnits = 15;
struct RsbShape {
  int cin, planes, bc, has_ds;
  int cinp, cp, np;            // x channels / branch channels / output channels padded to 8
  int w_off[kUnits + 1];       // word offsets
  int b_off[kUnits + 1];       // float offsets
  int nt[kUnits], kbw[kUnits], taps[kUnits];
};
RsbShape rsb_shape(int cin, int planes, int has_ds) {
  RsbShape s{};
  s.cin = cin, s.planes = planes, s.has_ds = has_ds, s.bc = cin * 26 / 64;
  s.cinp = (cin + 7) / 8 * 8, s.cp = (s.bc + 7) / 8 * 8, s.np = (planes + 7) / 8 * 8;
  int w = 0, bo = 0;
  for (int u = 0; u < kUnits; ++u) {
    if (u < 4) s.nt[u] = s.cp / 8, s.kbw[u] = s.cinp / 8, s.taps[u] = 1;
    else if (u < 14) s.nt[u] = s.cp / 8, s.kbw[u] = s.cp / 8, s.taps[u] = 9;
    else s.nt[u] = s.np / 8, s.kbw[u] = 4 * s.cp / 8 + (has_ds ? s.cinp / 8 : 0), s.taps[u] = 1;
    s.w_off[u] = w, s.b_off[u] = bo;
    w += s.taps[u] * s.nt[u] * s.kbw[u] * 32;
    w = (w + 3) / 4 * 4;
    bo += s.nt[u] * 8;
  }
  s.w_off[kUnits] = w, s.b_off[kUnits] = bo;
  return s;
}
struct PackArgs {
  const float *w[13], *b[13];
  RsbShape s;
  uint32_t *wout;
  float *bout;
};
// W(u, n, kk, tap): the folded fp32 weight behind entry (output channel n, K index kk) of unit u
__device__ __forceinline__ float unit_weight(const PackArgs &A, int u, int n, int kk, int tap) {
  const RsbShape &s = A.s;
  if (u < 4) {          // conv1 slice u: outputs u*bc + n, K = x channels
    if (n >= s.bc || kk >= s.cin) return 0.f;
    return A.w[0][(size_t)(u * s.bc + n) * s.cin + kk];
  }
  if (u < 14) {         // 3x3 conv u - 3
    if (n >= s.bc || kk >= s.bc) return 0.f;
    return A.w[u - 3][((size_t)n * s.bc + kk) * 9 + tap];
  }
  if (n >= s.planes) return 0.f;
  if (kk < 4 * s.cp) {  // conv3 over cat(out_1_1, out_2_2, out_3_3, out_4_4), each padded to cp channels
    const int i = kk / s.cp, c = kk % s.cp;
    if (c >= s.bc) return 0.f;
    return A.w[11][(size_t)n * (4 * s.bc) + i * s.bc + c];
  }
  const int c = kk - 4 * s.cp;   // downsample conv on x, concatenated along K
  if (!s.has_ds || c >= s.cin) return 0.f;
  return A.w[12][(size_t)n * s.cin + c];
}
__global__ void rsb_pack_kernel(const PackArgs A) {
  const RsbShape &s = A.s;
  const int u = blockIdx.y;
  const int words = s.taps[u] * s.nt[u] * s.kbw[u] * 32;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < words; e += gridDim.x * blockDim.x) {
    // word ((tap * NT + nt) * KBW + kb) * 32 + lane = { W[8nt + g][8kb + 2q], W[8nt + g][8kb + 2q + 1] }
    const int lane = e & 31, kb = (e >> 5) % s.kbw[u], nt = ((e >> 5) / s.kbw[u]) % s.nt[u], tap = (e >> 5) / (s.kbw[u] * s.nt[u]);
    const int n = 8 * nt + (lane >> 2), kk = 8 * kb + 2 * (lane & 3);
    A.wout[s.w_off[u] + e] = pack16x2<true>(unit_weight(A, u, n, kk, tap), unit_weight(A, u, n, kk + 1, tap));
  }
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < s.nt[u] * 8; n += blockDim.x) {
      float v = 0.f;
      if (u < 4) v = n < s.bc ? A.b[0][u * s.bc + n] : 0.f;
      else if (u < 14) v = n < s.bc ? A.b[u - 3][n] : 0.f;
      else if (n < s.planes) v = A.b[11][n] + (s.has_ds ? A.b[12][n] : 0.f);
      A.bout[s.b_off[u] + n] = v;
    }
}

// ------------------------------------------------------------------------------------------ levels
// planes of the 16-bit scratch
// ... 14 maps in 6 slots: a map takes the slot of one whose last reader is an EARLIER level (a level's CTAs read halo
// rows that other CTAs of the same launch must not be overwriting).  Live ranges (written at level -> last read):
// S0 0-1, S1 0-2, S2 0-3, S3 0-4, 11 1-8, 21 2-3, 22 3-8, 31 3-4, 32 4-5, 41 4-5, 33 5-8, 42 5-6, 43 6-7, 44 7-8.
enum {
  PL_S0 = 0, PL_S1 = 1, PL_S2 = 2, PL_S3 = 3, PL_11 = 4,
  PL_21 = 0, PL_22 = 1, PL_31 = 5, PL_32 = 0, PL_41 = 2, PL_33 = 3, PL_42 = 5, PL_43 = 0, PL_44 = 2,
  PL_COUNT = 6
};
struct LevelOp {
  int unit, k, plane;
  std::vector<int> in;   // input planes (-1 = x)
};
static std::vector<std::vector<LevelOp>> make_levels(bool has_ds) {
  return {
      {{0, 1, PL_S0, {-1}}, {1, 1, PL_S1, {-1}}, {2, 1, PL_S2, {-1}}, {3, 1, PL_S3, {-1}}},
      {{4, 3, PL_11, {PL_S0}}},
      {{5, 3, PL_21, {PL_S1, PL_11}}},
      {{6, 3, PL_22, {PL_21}}, {7, 3, PL_31, {PL_S2, PL_21}}},
      {{8, 3, PL_32, {PL_31, PL_22}}, {10, 3, PL_41, {PL_S3, PL_31}}},
      {{9, 3, PL_33, {PL_32}}, {11, 3, PL_42, {PL_41, PL_32}}},
      {{12, 3, PL_43, {PL_42, PL_33}}},
      {{13, 3, PL_44, {PL_43}}},
      {{14, 1, -1, has_ds ? std::vector<int>{PL_11, PL_22, PL_33, PL_44, -2} : std::vector<int>{PL_11, PL_22, PL_33, PL_44}}},
  };
}
static const std::vector<std::vector<LevelOp>> &rsb_levels(bool has_ds) {
  // function-local statics: initialised once, thread-safe (a process may drive several GPUs from several threads)
  static const std::vector<std::vector<LevelOp>> with_ds = make_levels(true), without = make_levels(false);
  return has_ds ? with_ds : without;
}

// shared-memory plan of one level; false if even a one-row tile does not fit
bool build_level(const RsbShape &s, const std::vector<LevelOp> &ops, int b, int H, int W, RsbArgs &a, size_t &smem) {
  a = RsbArgs{};
  a.halo = ops[0].k == 3 ? 1 : 0;
  a.MT = ceil_div(W, 16), a.BWP = 16 * a.MT + 2;
  a.H = H, a.W = W, a.B = b, a.cp = s.cp, a.cin = s.cin, a.cinp = s.cinp;
  // tensors: the distinct inputs of the level
  std::vector<int> planes;
  for (const LevelOp &o : ops)
    for (int p : o.in)
      if (std::find(planes.begin(), planes.end(), p) == planes.end()) planes.push_back(p);
  if ((int)planes.size() > RMAXT || (int)ops.size() > RMAXOP) return false;
  a.ntens = (int)planes.size(), a.nop = (int)ops.size();
  size_t wbytes = 0;
  for (int i = 0; i < a.nop; ++i) {
    RsbOp &o = a.op[i];
    const int u = ops[i].unit;
    o.nin = (int)ops[i].in.size();
    for (int e = 0; e < o.nin; ++e) o.in[e] = (int)(std::find(planes.begin(), planes.end(), ops[i].in[e]) - planes.begin());
    o.k = ops[i].k, o.nt = s.nt[u], o.kbw = s.kbw[u];
    o.wg = s.w_off[u], o.wn = s.w_off[u + 1] - s.w_off[u], o.bg = s.b_off[u], o.bn = s.b_off[u + 1] - s.b_off[u];
    o.relu = ops[i].plane >= 0, o.plane = ops[i].plane, o.gch = s.planes;
    o.wofs = (int)(wbytes / 4);
    wbytes += (size_t)o.wn * 4;
  }
  for (int i = 0; i < a.nop; ++i) {
    a.op[i].bofs = (int)(wbytes / 4);
    wbytes += (size_t)a.op[i].bn * 4;
  }
  wbytes = align_up(wbytes, 128);
  size_t row_bytes = 0;
  for (int t = 0; t < a.ntens; ++t) {
    RsbTensor &T = a.t[t];
    T.plane = planes[t];
    T.ch = planes[t] < 0 ? s.cinp : s.cp;
    T.sp = T.ch;
    if (planes[t] < 0 && (T.ch / 8) % 2 == 0) T.sp = T.ch + 8;   // odd number of 16-byte chunks per pixel
    if (planes[t] == -1) a.write_x = s.has_ds;
    row_bytes += (size_t)a.BWP * T.sp * 2;
  }
  // rows per tile: the fewest waves of CTAs (two per SM) weighted by the rows a CTA walks (+ ~3 rows of fixed cost)
  int th = 0;
  long long best = -1;
  for (int c = 1; c <= H; ++c) {
    if (wbytes + (size_t)(c + 2 * a.halo) * row_bytes > RSMEM) break;
    const long long ctas = (long long)b * ceil_div(H, c), slots = 2LL * num_sms();
    const long long cost = ((ctas + slots - 1) / slots) * (c + 2 * a.halo + 3);
    if (best < 0 || cost <= best) best = cost, th = c;
  }
  if (th < 1) return false;
  a.TH = th, a.tiles_y = ceil_div(H, th), a.rows = th + 2 * a.halo;
  size_t off = wbytes;
  a.tile_off = (int)off;
  for (int t = 0; t < a.ntens; ++t) {
    a.t[t].off = (int)off;
    off += align_up((size_t)a.rows * a.BWP * a.t[t].sp * 2, 128);
  }
  a.tile_bytes = (int)(off - wbytes);
  smem = off;
  return smem <= RSMEM;
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_rsb_block_supported(int cin, int planes, int h, int w) {
  if (cin < 8 || cin > 56 || planes < 1 || planes > 32 || h < 1 || w < 1 || w > 96) return 0;
  const RsbShape s = rsb_shape(cin, planes, 1);
  if (s.bc < 1 || s.cp > 24) return 0;
  for (const auto &lv : rsb_levels(true)) {
    RsbArgs a;
    size_t smem;
    if (!build_level(s, lv, 32, h, w, a, smem)) return 0;
  }
  return 1;
}

extern "C" size_t otp_rsb_block_pack_bytes(int cin, int planes, int has_downsample) {
  if (cin <= 0 || planes <= 0) return 0;
  const RsbShape s = rsb_shape(cin, planes, has_downsample);
  return align_up((size_t)s.w_off[kUnits] * 4 + (size_t)s.b_off[kUnits] * 4 + 256, 256);
}

extern "C" int otp_rsb_block_pack(const float *const *weights, const float *const *biases, int cin, int planes,
                                  int has_downsample, void *packed, size_t packed_bytes, otp_stream_t stream) {
  OTP_REQUIRE(weights && biases && packed && cin > 0 && planes > 0);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0);
  if (packed_bytes < otp_rsb_block_pack_bytes(cin, planes, has_downsample)) {
    set_error("otp_rsb_block_pack: buffer of %zu B, need %zu B", packed_bytes,
              otp_rsb_block_pack_bytes(cin, planes, has_downsample));
    return OTP_ERR_WORKSPACE;
  }
  PackArgs A{};
  A.s = rsb_shape(cin, planes, has_downsample);
  for (int i = 0; i < 13; ++i) {
    if (i == 12 && !has_downsample) continue;
    OTP_REQUIRE(weights[i] != nullptr && biases[i] != nullptr);
    A.w[i] = weights[i], A.b[i] = biases[i];
  }
  A.wout = static_cast<uint32_t *>(packed);
  A.bout = reinterpret_cast<float *>(static_cast<uint8_t *>(packed) + (size_t)A.s.w_off[kUnits] * 4);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_PACK, st);
  rsb_pack_kernel<<<dim3(8, kUnits), 256, 0, st>>>(A);
  return check_launch("rsb_pack_kernel");
}

extern "C" size_t otp_rsb_block_workspace_bytes(int b, int cin, int planes, int h, int w) {
  if (b <= 0 || cin <= 0 || planes <= 0 || h <= 0 || w <= 0) return 0;
  const RsbShape s = rsb_shape(cin, planes, 1);
  return align_up((size_t)PL_COUNT * b * h * w * s.cp * 2 + (size_t)b * h * w * s.cinp * 2 + 256, 256);
}

extern "C" int otp_rsb_block_forward(const void *packed, const float *x, long long x_bstride, float *y,
                                     long long y_bstride, int b, int cin, int planes, int has_downsample, int h, int w,
                                     void *workspace, size_t workspace_bytes, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && cin > 0 && planes > 0 && h > 0 && w > 0);
  if (!otp_rsb_block_supported(cin, planes, h, w)) {
    set_error("otp_rsb_block_forward: cin=%d planes=%d %dx%d not built", cin, planes, h, w);
    return OTP_ERR_UNSUPPORTED;
  }
  OTP_REQUIRE(has_downsample || cin == planes);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(packed && x && y && workspace && x != y);
  OTP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (workspace_bytes < otp_rsb_block_workspace_bytes(b, cin, planes, h, w)) {
    set_error("otp_rsb_block_forward: workspace of %zu B, need %zu B", workspace_bytes,
              otp_rsb_block_workspace_bytes(b, cin, planes, h, w));
    return OTP_ERR_WORKSPACE;
  }
  const RsbShape s = rsb_shape(cin, planes, has_downsample);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceOnce attr;
  if (attr.first()) {
    bool ok = true;
#define OTP_RSB_ATTR(C, N) ok = ok && set_max_smem(rsb_level_kernel<C, N>, RSMEM, "rsb_level_kernel")
    OTP_RSB_ATTR(1, 0); OTP_RSB_ATTR(2, 0); OTP_RSB_ATTR(3, 0);
    OTP_RSB_ATTR(0, 1); OTP_RSB_ATTR(0, 2); OTP_RSB_ATTR(0, 3); OTP_RSB_ATTR(0, 4);
#undef OTP_RSB_ATTR
    if (!ok) return OTP_ERR_CUDA;
  }
  const auto &levels = rsb_levels(has_downsample != 0);
  LaunchScope ls(K_RSB, st, (int)levels.size());
  for (const auto &lv : levels) {
    RsbArgs a;
    size_t smem;
    if (!build_level(s, lv, b, h, w, a, smem)) {
      set_error("otp_rsb_block_forward: no tile plan for cin=%d planes=%d %dx%d", cin, planes, h, w);
      return OTP_ERR_UNSUPPORTED;
    }
    a.wunits = static_cast<const uint32_t *>(packed);
    a.bunits = reinterpret_cast<const float *>(static_cast<const uint8_t *>(packed) + (size_t)s.w_off[kUnits] * 4);
    a.x = x, a.x_bs = x_bstride, a.y = y, a.y_bs = y_bstride;
    a.res = has_downsample ? nullptr : x, a.res_bs = x_bstride;
    a.planes = static_cast<__half *>(workspace);
    a.xplane = a.planes + (size_t)PL_COUNT * b * h * w * s.cp;
    const int grid = b * a.tiles_y;
    if (lv[0].k == 1) {
      switch (a.op[0].nt) {
        case 1: rsb_level_kernel<0, 1><<<grid, RTH, smem, st>>>(a); break;
        case 2: rsb_level_kernel<0, 2><<<grid, RTH, smem, st>>>(a); break;
        case 3: rsb_level_kernel<0, 3><<<grid, RTH, smem, st>>>(a); break;
        default: rsb_level_kernel<0, 4><<<grid, RTH, smem, st>>>(a); break;
      }
    } else if (s.cp == 8) {
      rsb_level_kernel<1, 0><<<grid, RTH, smem, st>>>(a);
    } else if (s.cp == 16) {
      rsb_level_kernel<2, 0><<<grid, RTH, smem, st>>>(a);
    } else {
      rsb_level_kernel<3, 0><<<grid, RTH, smem, st>>>(a);
    }
    if (int e = check_launch("rsb_level_kernel")) return e;
  }
  return OTP_OK;
}
