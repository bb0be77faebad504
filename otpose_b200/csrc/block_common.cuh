// Packed-weight layout and shared device helpers of the TransformerBlock kernels.
#pragma once
#include "common.cuh"

namespace otp {

// Compile-time shape of one ConvTransformer width.  OTPose instantiates exactly
// two: C = 8*J = 136 (temporal encoders, 2 heads) and C = J = 17 (flow encoder,
// 1 head) -- reference model/OTPose.py:206-216.
template <int C_>
struct BlockCfg;
template <>
struct BlockCfg<136> {
  static constexpr int C = 136, NH = 2, HS = 68, NPT = 16, NWARP = 9, NPAD = 144;
};
template <>
struct BlockCfg<17> {
  // 2 warps x 9 channels: the 17-wide blocks are latency / barrier bound, so small CTAs
  // (many resident per SM) beat wide ones
  static constexpr int C = 17, NH = 1, HS = 17, NPT = 9, NWARP = 2, NPAD = 18;
};

constexpr int kTT = 64;           // output tokens per tile
constexpr int kTPL = 2;           // tokens per lane
constexpr int kLD = kTT + 1;      // smem row stride of a (C, kTT) tile
constexpr int kMaxIn = 2 * kTT + 2;  // input tokens of a stride-2 tile incl. halo
constexpr int kLDX = kMaxIn + 1;  // smem row stride of the input tile

inline int npad_of(int c) { return c == 136 ? 144 : 18; }

// Offsets (in floats) of every section of the packed fp32 block weights.
struct BlockPackLayout {
  size_t wqT, wkT, wpT;        // [C][NPAD]  transposed pointwise weights
  size_t wv;                   // [C][C]     value weight, reference layout
  size_t w1T, w2T;             // [4][C][NPAD]  MLP weights per hidden chunk of C
  size_t ln1_w, ln1_b, ln2_w, ln2_b, qn_w, qn_b, kn_w, kn_b, vn_w, vn_b;  // [C]
  size_t dwq, dwk, dwv;        // [C][3]
  size_t bq, bk, bv, bp, b2, sa, sm;  // [NPAD] (zero / one padded)
  size_t b1;                   // [4][NPAD]
  size_t total;                // floats
};

inline BlockPackLayout block_pack_layout(int c) {
  const size_t np = npad_of(c), C = c;
  BlockPackLayout L{};
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t r = o;
    o += (n + 63) / 64 * 64;  // 256-byte aligned sections
    return r;
  };
  L.wqT = take(C * np);
  L.wkT = take(C * np);
  L.wpT = take(C * np);
  L.wv = take(C * C);
  L.w1T = take(4 * C * np);
  L.w2T = take(4 * C * np);
  L.ln1_w = take(C); L.ln1_b = take(C); L.ln2_w = take(C); L.ln2_b = take(C);
  L.qn_w = take(C); L.qn_b = take(C); L.kn_w = take(C); L.kn_b = take(C);
  L.vn_w = take(C); L.vn_b = take(C);
  L.dwq = take(3 * C); L.dwk = take(3 * C); L.dwv = take(3 * C);
  L.bq = take(np); L.bk = take(np); L.bv = take(np); L.bp = take(np); L.b2 = take(np);
  L.sa = take(np); L.sm = take(np);
  L.b1 = take(4 * np);
  L.total = o;
  return L;
}

// Device view of a packed block (pointer per section).
struct BlockPack {
  const float *wqT, *wkT, *wpT, *wv, *w1T, *w2T;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qn_w, *qn_b, *kn_w, *kn_b, *vn_w, *vn_b;
  const float *dwq, *dwk, *dwv;
  const float *bq, *bk, *bv, *bp, *b2, *sa, *sm, *b1;
};

inline BlockPack block_pack_view(const void *packed, int c) {
  const float *f = static_cast<const float *>(packed);
  BlockPackLayout L = block_pack_layout(c);
  BlockPack v;
  v.wqT = f + L.wqT; v.wkT = f + L.wkT; v.wpT = f + L.wpT; v.wv = f + L.wv;
  v.w1T = f + L.w1T; v.w2T = f + L.w2T;
  v.ln1_w = f + L.ln1_w; v.ln1_b = f + L.ln1_b; v.ln2_w = f + L.ln2_w; v.ln2_b = f + L.ln2_b;
  v.qn_w = f + L.qn_w; v.qn_b = f + L.qn_b; v.kn_w = f + L.kn_w; v.kn_b = f + L.kn_b;
  v.vn_w = f + L.vn_w; v.vn_b = f + L.vn_b;
  v.dwq = f + L.dwq; v.dwk = f + L.dwk; v.dwv = f + L.dwv;
  v.bq = f + L.bq; v.bk = f + L.bk; v.bv = f + L.bv; v.bp = f + L.bp; v.b2 = f + L.b2;
  v.sa = f + L.sa; v.sm = f + L.sm; v.b1 = f + L.b1;
  return v;
}

// Workspace of one block forward (fp32 sections, 256-byte aligned).
struct BlockWorkspace {
  size_t gram_part;  // [B][nchunk][C][HS]
  size_t weffT;      // [B][C][NPAD]
  size_t beff;       // [B][NPAD]
  size_t obuf;       // [B][C*Tout]   the (nh, T', hs) "scramble" buffer
  size_t total;      // bytes
  int nchunk, tiles_per_chunk, tout;
};

inline BlockWorkspace block_workspace(int b, int c, int t, int n_head, int stride) {
  BlockWorkspace w{};
  w.tout = stride == 1 ? t : (t - 1) / 2 + 1;
  const int tiles = ceil_div(w.tout, kTT);
  // front-pass CTAs: ~2 per SM for the wide (1 CTA/SM) C=136 kernel, many small ones for C=17
  const long long target = c == 136 ? 296 : 4096;
  w.tiles_per_chunk = (int)((((long long)tiles * b) + target - 1) / target);
  if (w.tiles_per_chunk < 1) w.tiles_per_chunk = 1;
  w.nchunk = ceil_div(tiles, w.tiles_per_chunk);
  if (c == 17) {   // thread-per-token kernels (block_small.cu): one partial Gram per 128-token CTA
    w.tiles_per_chunk = 1;
    w.nchunk = ceil_div(w.tout, 128);
  }
  const size_t hs = c / n_head, np = npad_of(c);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += align_up(bytes, 256);
    return r;
  };
  w.gram_part = take((size_t)b * w.nchunk * c * hs * 4);
  w.weffT = take((size_t)b * c * np * 4);
  w.beff = take((size_t)b * np * 4);
  w.obuf = take((size_t)b * c * w.tout * 4);
  w.total = o;
  return w;
}

}  // namespace otp
