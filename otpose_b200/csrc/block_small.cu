// a2-a5 for the narrow flow encoder (C = 17, one head; reference model/OTPose.py:214-216):
// 0.6 % of the head's FLOPs but 22 % of its launches, so it gets kernels shaped for it
// instead of the tiled CUDA-core path: ONE THREAD PER TOKEN with the token's 17 channels
// in registers.  LayerNorms, the depthwise taps (the two neighbour tokens are re-read --
// they are L1/L2 hits -- and re-normalised in registers), the 17x17 projections and the
// 17 -> 68 -> 17 MLP need no shared-memory exchange and no barrier; only the channel
// Gram (a reduction over tokens) goes through shared memory once per CTA.  Weights are
// shared-memory broadcasts; every global access is a coalesced row segment.
#include "block_common.cuh"
#include "block_fold.cuh"

namespace otp {
namespace {
constexpr int SC = 17, SNP = 18, STH = 128;   // channels, padded row of the fp32 pack, threads

__device__ __forceinline__ void ln17(float (&v)[SC], const float *w, const float *b) {
  float mu = 0.f;
#pragma unroll
  for (int c = 0; c < SC; ++c) mu += v[c];
  mu *= (1.0f / SC);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < SC; ++c) {
    const float d = v[c] - mu;
    var = fmaf(d, d, var);
  }
  const float rstd = 1.0f / sqrtf(var * (1.0f / SC) + 1e-5f);
#pragma unroll
  for (int c = 0; c < SC; ++c) v[c] = fmaf((v[c] - mu) * rstd, w[c], b[c]);
}

// d[c] = sum_k dw[c][k] * LN1(x)[c][stride*t - 1 + k]   (zero padding outside [0, T))
__device__ __forceinline__ void ln1_dwconv(const float *__restrict__ xb, int T, int t, int stride,
                                           const float *ln_w, const float *ln_b, const float *dw,
                                           float (&d)[SC]) {
#pragma unroll
  for (int c = 0; c < SC; ++c) d[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int ti = stride * t - 1 + k;
    if (ti >= 0 && ti < T) {
      float y[SC];
#pragma unroll
      for (int c = 0; c < SC; ++c) y[c] = __ldg(xb + (size_t)c * T + ti);
      ln17(y, ln_w, ln_b);
#pragma unroll
      for (int c = 0; c < SC; ++c) d[c] = fmaf(dw[3 * c + k], y[c], d[c]);
    }
  }
}

// out[n] = bias[n] + sum_c wT[c*SNP + n] * in[c]
__device__ __forceinline__ void matvec17(const float *wT, const float *bias, const float (&in)[SC], float (&out)[SC]) {
#pragma unroll
  for (int n = 0; n < SC; ++n) out[n] = bias[n];
#pragma unroll
  for (int c = 0; c < SC; ++c)
#pragma unroll
    for (int n = 0; n < SC; ++n) out[n] = fmaf(wT[c * SNP + n], in[c], out[n]);
}

__device__ __forceinline__ void stage(float *dst, const float *__restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += STH) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(STH)
small_front_kernel(BlockPack P, const float *__restrict__ x, float *__restrict__ gram_part, int T, int Tout,
                   int stride, int tiles, float qscale) {
  __shared__ float wq[SC * SNP], wk[SC * SNP], bq[SNP], bk[SNP];
  __shared__ float l1w[SC], l1b[SC], qnw[SC], qnb[SC], knw[SC], knb[SC], dwq[3 * SC], dwk[3 * SC];
  __shared__ float qs[SC][STH + 1], ks[SC][STH + 1];
  stage(wq, P.wqT, SC * SNP); stage(wk, P.wkT, SC * SNP); stage(bq, P.bq, SNP); stage(bk, P.bk, SNP);
  stage(l1w, P.ln1_w, SC); stage(l1b, P.ln1_b, SC); stage(qnw, P.qn_w, SC); stage(qnb, P.qn_b, SC);
  stage(knw, P.kn_w, SC); stage(knb, P.kn_b, SC); stage(dwq, P.dwq, 3 * SC); stage(dwk, P.dwk, 3 * SC);
  __syncthreads();
  const int b = blockIdx.y, tile = blockIdx.x, t = tile * STH + threadIdx.x;
  const float *xb = x + (size_t)b * SC * T;
  float q[SC], k[SC];
  if (t < Tout) {
    float d[SC];
    ln1_dwconv(xb, T, t, stride, l1w, l1b, dwq, d);
    ln17(d, qnw, qnb);
    matvec17(wq, bq, d, q);
    ln1_dwconv(xb, T, t, stride, l1w, l1b, dwk, d);
    ln17(d, knw, knb);
    matvec17(wk, bk, d, k);
  }
#pragma unroll
  for (int c = 0; c < SC; ++c) {
    qs[c][threadIdx.x] = t < Tout ? q[c] * qscale : 0.f;
    ks[c][threadIdx.x] = t < Tout ? k[c] : 0.f;
  }
  __syncthreads();
  float *gp = gram_part + (size_t)(b * tiles + tile) * SC * SC;
  for (int e = threadIdx.x; e < SC * SC; e += STH) {
    const int i = e / SC, j = e % SC;
    float acc = 0.f;
#pragma unroll 8
    for (int tok = 0; tok < STH; ++tok) acc = fmaf(qs[i][tok], ks[j][tok], acc);
    gp[e] = acc;
  }
}

__global__ void __launch_bounds__(STH)
small_apply_kernel(BlockPack P, const float *__restrict__ x, const float *__restrict__ weffT,
                   const float *__restrict__ beff, float *__restrict__ obuf, int T, int Tout, int stride) {
  __shared__ float we[SC * SNP], be[SNP], l1w[SC], l1b[SC], vnw[SC], vnb[SC], dwv[3 * SC];
  const int b = blockIdx.y, t = blockIdx.x * STH + threadIdx.x;
  stage(we, weffT + (size_t)b * SC * SNP, SC * SNP); stage(be, beff + (size_t)b * SNP, SNP);
  stage(l1w, P.ln1_w, SC); stage(l1b, P.ln1_b, SC); stage(vnw, P.vn_w, SC); stage(vnb, P.vn_b, SC);
  stage(dwv, P.dwv, 3 * SC);
  __syncthreads();
  if (t >= Tout) return;
  float d[SC], o[SC];
  ln1_dwconv(x + (size_t)b * SC * T, T, t, stride, l1w, l1b, dwv, d);
  ln17(d, vnw, vnb);
  matvec17(we, be, d, o);
  // one head: obuf[b][t][c] -- the buffer the reference re-reads as (C, T') (blocks.py:447)
  float *dst = obuf + (size_t)b * SC * Tout + (size_t)t * SC;
#pragma unroll
  for (int c = 0; c < SC; ++c) dst[c] = o[c];
}

__global__ void __launch_bounds__(STH)
small_back_kernel(BlockPack P, const float *__restrict__ x, const float *__restrict__ obuf,
                  float *__restrict__ y, int T, int Tout, int stride) {
  // W1n[h][c] (row = hidden unit, c padded to 20) and W2n[h][m] (m padded to 20): float4 broadcasts
  __shared__ __align__(16) float w1n[4 * SC][20], w2n[4 * SC][20];
  __shared__ float wp[SC * SNP], bp[SNP], sa[SNP], b2[SNP], sm[SNP], b1[4 * SC], l2w[SC], l2b[SC];
  for (int e = threadIdx.x; e < 4 * SC * 20; e += STH) {
    const int h = e / 20, c = e % 20, q = h / SC, n = h % SC;
    w1n[h][c] = c < SC ? __ldg(P.w1T + (size_t)q * SC * SNP + c * SNP + n) : 0.f;   // W1[q*17+n][c]
    w2n[h][c] = c < SC ? __ldg(P.w2T + (size_t)q * SC * SNP + n * SNP + c) : 0.f;   // W2[m=c][q*17+n]
  }
  for (int e = threadIdx.x; e < 4 * SC; e += STH) b1[e] = __ldg(P.b1 + (e / SC) * SNP + e % SC);
  stage(wp, P.wpT, SC * SNP); stage(bp, P.bp, SNP); stage(sa, P.sa, SNP); stage(b2, P.b2, SNP);
  stage(sm, P.sm, SNP); stage(l2w, P.ln2_w, SC); stage(l2b, P.ln2_b, SC);
  __syncthreads();
  const int b = blockIdx.y, t = blockIdx.x * STH + threadIdx.x;
  if (t >= Tout) return;
  float o2[SC], u[SC], l[SC];
  const float *ob = obuf + (size_t)b * SC * Tout + t;
#pragma unroll
  for (int c = 0; c < SC; ++c) o2[c] = __ldg(ob + (size_t)c * Tout);   // (nh,T',hs) buffer viewed (C,T')
  matvec17(wp, bp, o2, u);
  const float *xb = x + (size_t)b * SC * T;
#pragma unroll
  for (int c = 0; c < SC; ++c) {
    const float *xr = xb + (size_t)c * T;
    float skip;
    if (stride == 1) {
      skip = __ldg(xr + t);
    } else {   // MaxPool1d(3, 2, 1)
      const int c0 = 2 * t;
      skip = __ldg(xr + c0);
      if (c0 - 1 >= 0) skip = fmaxf(skip, __ldg(xr + c0 - 1));
      if (c0 + 1 < T) skip = fmaxf(skip, __ldg(xr + c0 + 1));
    }
    u[c] = fmaf(sa[c], u[c], skip);
    l[c] = u[c];
  }
  ln17(l, l2w, l2b);
  float acc[20];
#pragma unroll
  for (int m = 0; m < 20; ++m) acc[m] = 0.f;
#pragma unroll 2
  for (int h = 0; h < 4 * SC; ++h) {
    const float4 *w1 = reinterpret_cast<const float4 *>(w1n[h]);
    float s = b1[h];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      const float4 w = w1[qd];
      s = fmaf(w.x, l[4 * qd], fmaf(w.y, l[4 * qd + 1], fmaf(w.z, l[4 * qd + 2], fmaf(w.w, l[4 * qd + 3], s))));
    }
    s = fmaf(w1n[h][16], l[16], s);
    const float g = gelu_erf(s);
    const float4 *w2 = reinterpret_cast<const float4 *>(w2n[h]);
#pragma unroll
    for (int qd = 0; qd < 5; ++qd) {
      const float4 w = w2[qd];
      acc[4 * qd] = fmaf(w.x, g, acc[4 * qd]);
      acc[4 * qd + 1] = fmaf(w.y, g, acc[4 * qd + 1]);
      acc[4 * qd + 2] = fmaf(w.z, g, acc[4 * qd + 2]);
      acc[4 * qd + 3] = fmaf(w.w, g, acc[4 * qd + 3]);
    }
  }
  float *yb = y + (size_t)b * SC * Tout + t;
#pragma unroll
  for (int c = 0; c < SC; ++c) yb[(size_t)c * Tout] = fmaf(sm[c], acc[c] + b2[c], u[c]);
}
}  // namespace

int block_forward_small(const void *packed, const float *x, float *y, int b, int t, int stride, void *ws,
                        cudaStream_t st) {
  const BlockWorkspace W = block_workspace(b, SC, t, 1, stride);
  const BlockPack P = block_pack_view(packed, SC);
  char *wsb = static_cast<char *>(ws);
  float *gram = reinterpret_cast<float *>(wsb + W.gram_part);
  float *weffT = reinterpret_cast<float *>(wsb + W.weffT);
  float *beff = reinterpret_cast<float *>(wsb + W.beff);
  float *obuf = reinterpret_cast<float *>(wsb + W.obuf);
  const int tiles = ceil_div(W.tout, STH);   // <= W.nchunk partial Grams (sized for 64-token tiles)
  {
    LaunchScope ls(K_BLOCK_FRONT, st);
    small_front_kernel<<<dim3(tiles, b), STH, 0, st>>>(P, x, gram, t, W.tout, stride, tiles, 1.0f / sqrtf((float)SC));
  }
  {
    LaunchScope ls(K_BLOCK_FOLD, st);
    block_fold_kernel<SC, 0><<<dim3(b, FoldCfg<SC>::NBLK), kFoldThreads, 0, st>>>(P.wv, P.bv, gram, tiles, weffT, beff,
                                                                                SC, SNP);
  }
  {
    LaunchScope ls(K_BLOCK_APPLY, st);
    small_apply_kernel<<<dim3(tiles, b), STH, 0, st>>>(P, x, weffT, beff, obuf, t, W.tout, stride);
  }
  {
    LaunchScope ls(K_BLOCK_BACK, st);
    small_back_kernel<<<dim3(tiles, b), STH, 0, st>>>(P, x, obuf, y, t, W.tout, stride);
  }
  return check_launch("block_forward_small");
}

}  // namespace otp
