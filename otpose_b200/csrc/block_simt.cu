// a2-a5 (fp32 path): one ConvTransformer TransformerBlock as three fused passes
// over the token axis plus a tiny per-clip softmax/fold kernel.
//
// Reference (model/blocks.py:264-279, 400-452) launches ~85 ATen kernels per block
// and round-trips q, k, v, the 4C-wide MLP hidden tensor and every LayerNorm
// intermediate through HBM.  Here, per block and per clip:
//
//   front   x -> LN1 -> {dw_q, dw_k} -> {LN_q, LN_k} -> {Wq, Wk} -> partial Gram
//           S_h += (q_h/sqrt(hs)) k_h^T, accumulated in registers over a token
//           chunk; only the (C, hs) partial sums reach HBM.
//   fold    fixed-order sum of the partial Grams, row softmax (68x68 per head),
//           and W_eff = softmax(S) . W_v , b_eff = softmax(S) . b_v per clip:
//           att @ (W_v vn + b_v) == W_eff vn + b_eff, so v is never formed.
//   apply   x -> LN1 -> dw_v -> LN_v -> W_eff -> o, stored token-major
//           [head][token][channel]: exactly the buffer that the reference's
//           `transpose(2,3).contiguous().view(B,C,-1)` (blocks.py:447) re-reads
//           channel-major, so the "scramble" costs nothing.
//   back    scramble buffer -> W_proj -> u = skip(x) + s_a*(.) -> LN2 -> W_1 ->
//           GELU -> W_2 -> y = u + s_m*(.); the 4C hidden activations live in
//           shared memory one C-wide chunk at a time.
//
// This file is the CUDA-core fp32 implementation (exact-precision mode and the
// C=17 flow encoder); block_tc.cu holds the tcgen05 version of the same passes.
#include "block_common.cuh"
#include "block_fold.cuh"

namespace otp {

constexpr int kRedLd = 132;  // >= kMaxIn

template <int NPT>
__device__ __forceinline__ void load_w(const float *__restrict__ p, float (&w)[NPT]) {
  if constexpr (NPT % 4 == 0) {
#pragma unroll
    for (int q = 0; q < NPT / 4; ++q) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(p) + q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
  } else if constexpr (NPT % 2 == 0) {
#pragma unroll
    for (int q = 0; q < NPT / 2; ++q) {
      float2 v = __ldg(reinterpret_cast<const float2 *>(p) + q);
      w[2 * q] = v.x; w[2 * q + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < NPT; ++q) w[q] = __ldg(p + q);
  }
}

// acc[i][j] += sum_k wt[k*NPAD + i] * in_s[k*ld + 32*j]
// (wt already offset to this warp's first output channel, in_s to this lane).
template <class Cfg>
__device__ __forceinline__ void tile_gemm(float (&acc)[Cfg::NPT][kTPL], const float *__restrict__ wt,
                                          int K, const float *__restrict__ in_s, int ld) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[kTPL];
#pragma unroll
    for (int j = 0; j < kTPL; ++j) a[j] = in_s[k * ld + 32 * j];
    float w[Cfg::NPT];
    load_w<Cfg::NPT>(wt + (size_t)k * Cfg::NPAD, w);
#pragma unroll
    for (int i = 0; i < Cfg::NPT; ++i)
#pragma unroll
      for (int j = 0; j < kTPL; ++j) acc[i][j] = fmaf(w[i], a[j], acc[i][j]);
  }
}

// Channel LayerNorm (model/blocks.py:95-110) of the columns [0, ntok) of the smem
// tile S[c*ld + i], in place; two-pass mean / biased variance like the reference.
// Columns outside [vlo, vhi) are set to 0 (they stand for conv zero padding).
template <class Cfg>
__device__ void ln_tile(float *S, int ld, int ntok, const float *__restrict__ w,
                        const float *__restrict__ b, float *red, float *stat, int vlo, int vhi) {
  constexpr int C = Cfg::C, NW = Cfg::NWARP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < ntok; i += 32) {
    float s = 0.f;
    for (int c = warp; c < C; c += NW) s += S[c * ld + i];
    red[warp * kRedLd + i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntok; i += NW * 32) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NW; ++q) s += red[q * kRedLd + i];
    stat[i] = s / (float)C;
  }
  __syncthreads();
  for (int i = lane; i < ntok; i += 32) {
    const float mu = stat[i];
    float s = 0.f;
    for (int c = warp; c < C; c += NW) {
      float d = S[c * ld + i] - mu;
      s = fmaf(d, d, s);
    }
    red[warp * kRedLd + i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntok; i += NW * 32) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NW; ++q) s += red[q * kRedLd + i];
    stat[kRedLd + i] = 1.0f / sqrtf(s / (float)C + 1e-5f);
  }
  __syncthreads();
  for (int c = warp; c < C; c += NW) {
    const float wc = __ldg(w + c), bc = __ldg(b + c);
    for (int i = lane; i < ntok; i += 32) {
      float v = (S[c * ld + i] - stat[i]) * stat[kRedLd + i] * wc + bc;
      S[c * ld + i] = (i >= vlo && i < vhi) ? v : 0.f;
    }
  }
  __syncthreads();
}

// Xs[c][i] = LN1(x)[c][stride*t0 - 1 + i], zero outside [0, T)  (conv zero padding)
template <class Cfg>
__device__ void load_ln1_tile(float *Xs, const float *__restrict__ xb, int T, int t0, int stride,
                              const float *ln_w, const float *ln_b, float *red, float *stat) {
  constexpr int C = Cfg::C, NW = Cfg::NWARP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ni = stride * kTT + 2;
  const int base = stride * t0 - 1;
  for (int c = warp; c < C; c += NW) {
    const float *xr = xb + (size_t)c * T;
    for (int i = lane; i < ni; i += 32) {
      int ti = base + i;
      Xs[c * kLDX + i] = (ti >= 0 && ti < T) ? __ldg(xr + ti) : 0.f;
    }
  }
  __syncthreads();
  int vlo = -base, vhi = T - base;
  ln_tile<Cfg>(Xs, kLDX, ni, ln_w, ln_b, red, stat, vlo < 0 ? 0 : vlo, vhi > ni ? ni : vhi);
}

// Ps[c][t] = sum_k dw[c][k] * Xs[c][stride*t + k]   (depthwise Conv1d k=3, pad=1)
template <class Cfg>
__device__ void dwconv_tile(float *Ps, const float *Xs, const float *__restrict__ dw, int stride) {
  constexpr int C = Cfg::C, NW = Cfg::NWARP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += NW) {
    const float w0 = __ldg(dw + 3 * c), w1 = __ldg(dw + 3 * c + 1), w2 = __ldg(dw + 3 * c + 2);
    const float *xr = Xs + c * kLDX;
    for (int t = lane; t < kTT; t += 32) {
      int i = stride * t;
      Ps[c * kLD + t] = w0 * xr[i] + w1 * xr[i + 1] + w2 * xr[i + 2];
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------ front pass
template <int C_>
__global__ void __launch_bounds__(BlockCfg<C_>::NWARP * 32)
block_front_kernel(BlockPack P, const float *__restrict__ x, float *__restrict__ gram_part, int T,
                   int Tout, int stride, int tiles_per_chunk, int nchunk, float qscale) {
  using Cfg = BlockCfg<C_>;
  constexpr int C = Cfg::C, HS = Cfg::HS, NPT = Cfg::NPT;
  extern __shared__ float smem[];
  float *Xs = smem;
  float *Ps = Xs + C * kLDX;
  float *Qs = Ps + C * kLD;
  float *Ks = Qs + C * kLD;
  float *red = Ks + C * kLD;
  float *stat = red + Cfg::NWARP * kRedLd;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float *xb = x + (size_t)b * C * T;
  const int n0 = warp * NPT;

  // Gram ownership: rows (2*ip, 2*ip+1), columns [jq*17, jq*17+17) of the rows' head
  constexpr int JB = 17, NJQ = HS / JB, NPAIR = (C + 1) / 2;
  const bool gact = threadIdx.x < NPAIR * NJQ;
  const int r0 = 2 * (threadIdx.x / NJQ), jq = threadIdx.x % NJQ;
  const int head = gact ? r0 / HS : 0;
  const bool has_r1 = r0 + 1 < C;
  float g[2][JB];
#pragma unroll
  for (int jj = 0; jj < JB; ++jj) g[0][jj] = g[1][jj] = 0.f;

  const int tiles = ceil_div(Tout, kTT);
  const int tile_end = min(tiles, (chunk + 1) * tiles_per_chunk);
  for (int tile = chunk * tiles_per_chunk; tile < tile_end; ++tile) {
    const int t0 = tile * kTT;
    const int nvalid = min(kTT, Tout - t0);
    load_ln1_tile<Cfg>(Xs, xb, T, t0, stride, P.ln1_w, P.ln1_b, red, stat);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      dwconv_tile<Cfg>(Ps, Xs, which == 0 ? P.dwq : P.dwk, stride);
      ln_tile<Cfg>(Ps, kLD, kTT, which == 0 ? P.qn_w : P.kn_w, which == 0 ? P.qn_b : P.kn_b, red, stat, 0, kTT);
      float acc[NPT][kTPL];
#pragma unroll
      for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int j = 0; j < kTPL; ++j) acc[i][j] = 0.f;
      tile_gemm<Cfg>(acc, (which == 0 ? P.wqT : P.wkT) + n0, C, Ps + lane, kLD);
      float *dst = which == 0 ? Qs : Ks;
      const float *bias = which == 0 ? P.bq : P.bk;
      const float sc = which == 0 ? qscale : 1.0f;
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const int n = n0 + i;
        if (n < C) {
          const float bn = __ldg(bias + n);
#pragma unroll
          for (int j = 0; j < kTPL; ++j) dst[n * kLD + lane + 32 * j] = (acc[i][j] + bn) * sc;
        }
      }
      __syncthreads();
    }
    if (gact) {
      const float *q0 = Qs + r0 * kLD;
      const float *q1 = Qs + (has_r1 ? r0 + 1 : r0) * kLD;
      const float *kb = Ks + (head * HS + jq * JB) * kLD;
      for (int t = 0; t < nvalid; ++t) {
        const float a0 = q0[t], a1 = q1[t];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const float kv = kb[jj * kLD + t];
          g[0][jj] = fmaf(a0, kv, g[0][jj]);
          g[1][jj] = fmaf(a1, kv, g[1][jj]);
        }
      }
    }
    // the next tile's first smem writes to Qs/Ks happen after several barriers
  }
  if (gact) {
    float *gp = gram_part + ((size_t)(b * nchunk + chunk) * C + r0) * HS + jq * JB;
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      gp[jj] = g[0][jj];
      if (has_r1) gp[HS + jj] = g[1][jj];
    }
  }
}

// ------------------------------------------------------------------ apply pass
template <int C_>
__global__ void __launch_bounds__(BlockCfg<C_>::NWARP * 32)
block_apply_kernel(BlockPack P, const float *__restrict__ x, const float *__restrict__ weffT,
                   const float *__restrict__ beff, float *__restrict__ obuf, int T, int Tout,
                   int stride) {
  using Cfg = BlockCfg<C_>;
  constexpr int C = Cfg::C, HS = Cfg::HS, NPT = Cfg::NPT, NPAD = Cfg::NPAD;
  extern __shared__ float smem[];
  float *Xs = smem;
  float *Ps = Xs + C * kLDX;
  float *red = Ps + C * kLD;
  float *stat = red + Cfg::NWARP * kRedLd;
  float *Os = Xs;  // Xs is dead once dw_v has been taken

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * kTT;
  const int nvalid = min(kTT, Tout - t0);
  const int n0 = warp * NPT;
  load_ln1_tile<Cfg>(Xs, x + (size_t)b * C * T, T, t0, stride, P.ln1_w, P.ln1_b, red, stat);
  dwconv_tile<Cfg>(Ps, Xs, P.dwv, stride);
  ln_tile<Cfg>(Ps, kLD, kTT, P.vn_w, P.vn_b, red, stat, 0, kTT);
  float acc[NPT][kTPL];
#pragma unroll
  for (int i = 0; i < NPT; ++i)
#pragma unroll
    for (int j = 0; j < kTPL; ++j) acc[i][j] = 0.f;
  tile_gemm<Cfg>(acc, weffT + (size_t)b * C * NPAD + n0, C, Ps + lane, kLD);
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int n = n0 + i;
    if (n < C) {
      const float bn = __ldg(beff + (size_t)b * NPAD + n);
#pragma unroll
      for (int j = 0; j < kTPL; ++j) Os[n * kLD + lane + 32 * j] = acc[i][j] + bn;
    }
  }
  __syncthreads();
  // token-major store: obuf[b][h][t][c'] -- contiguous nvalid*HS floats per head
  for (int h = 0; h < Cfg::NH; ++h) {
    float *dst = obuf + (size_t)b * C * Tout + (size_t)h * Tout * HS + (size_t)t0 * HS;
    for (int e = threadIdx.x; e < nvalid * HS; e += Cfg::NWARP * 32)
      dst[e] = Os[(h * HS + e % HS) * kLD + e / HS];
  }
}

// ------------------------------------------------------------------ back pass
template <int C_>
__global__ void __launch_bounds__(BlockCfg<C_>::NWARP * 32)
block_back_kernel(BlockPack P, const float *__restrict__ x, const float *__restrict__ obuf,
                  float *__restrict__ y, int T, int Tout, int stride) {
  using Cfg = BlockCfg<C_>;
  constexpr int C = Cfg::C, NPT = Cfg::NPT, NPAD = Cfg::NPAD;
  extern __shared__ float smem[];
  float *Us = smem;            // u = skip + s_a * proj(...)
  float *Ps = Us + C * kLD;    // scramble tile, then LN2(u)
  float *Hs = Ps + C * kLD;    // one C-wide chunk of the hidden activations
  float *red = Hs + C * kLD;
  float *stat = red + Cfg::NWARP * kRedLd;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, t0 = blockIdx.x * kTT;
  const int nvalid = min(kTT, Tout - t0);
  const int n0 = warp * NPT;

  // the (nh, T', hs) buffer re-read as (C, T') -- the reference's scramble view
  const float *ob = obuf + (size_t)b * C * Tout;
  for (int c = warp; c < C; c += Cfg::NWARP)
    for (int t = lane; t < kTT; t += 32)
      Ps[c * kLD + t] = t < nvalid ? __ldg(ob + (size_t)c * Tout + t0 + t) : 0.f;
  __syncthreads();

  float acc[NPT][kTPL];
#pragma unroll
  for (int i = 0; i < NPT; ++i)
#pragma unroll
    for (int j = 0; j < kTPL; ++j) acc[i][j] = 0.f;
  tile_gemm<Cfg>(acc, P.wpT + n0, C, Ps + lane, kLD);
  const float *xb = x + (size_t)b * C * T;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int n = n0 + i;
    if (n < C) {
      const float bn = __ldg(P.bp + n), sa = __ldg(P.sa + n);
      const float *xr = xb + (size_t)n * T;
#pragma unroll
      for (int j = 0; j < kTPL; ++j) {
        const int t = lane + 32 * j, tt = t0 + t;
        float skip = 0.f;
        if (tt < Tout) {
          if (stride == 1) {
            skip = __ldg(xr + tt);
          } else {  // MaxPool1d(3, 2, 1), -inf padding
            const int c0 = 2 * tt;
            skip = __ldg(xr + c0);
            if (c0 - 1 >= 0) skip = fmaxf(skip, __ldg(xr + c0 - 1));
            if (c0 + 1 < T) skip = fmaxf(skip, __ldg(xr + c0 + 1));
          }
        }
        Us[n * kLD + t] = skip + sa * (acc[i][j] + bn);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * kTT; e += Cfg::NWARP * 32) {
    const int c = e / kTT, t = e % kTT;
    Ps[c * kLD + t] = Us[c * kLD + t];
  }
  __syncthreads();
  ln_tile<Cfg>(Ps, kLD, kTT, P.ln2_w, P.ln2_b, red, stat, 0, kTT);

  float yacc[NPT][kTPL];
#pragma unroll
  for (int i = 0; i < NPT; ++i)
#pragma unroll
    for (int j = 0; j < kTPL; ++j) yacc[i][j] = 0.f;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
      for (int j = 0; j < kTPL; ++j) acc[i][j] = 0.f;
    tile_gemm<Cfg>(acc, P.w1T + (size_t)q * C * NPAD + n0, C, Ps + lane, kLD);
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int n = n0 + i;
      if (n < C) {
        const float bn = __ldg(P.b1 + q * NPAD + n);
#pragma unroll
        for (int j = 0; j < kTPL; ++j) Hs[n * kLD + lane + 32 * j] = gelu_erf(acc[i][j] + bn);
      }
    }
    __syncthreads();
    tile_gemm<Cfg>(yacc, P.w2T + (size_t)q * C * NPAD + n0, C, Hs + lane, kLD);
    __syncthreads();
  }
  float *yb = y + (size_t)b * C * Tout;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int n = n0 + i;
    if (n < C) {
      const float bn = __ldg(P.b2 + n), sm = __ldg(P.sm + n);
#pragma unroll
      for (int j = 0; j < kTPL; ++j) {
        const int t = lane + 32 * j;
        if (t < nvalid) yb[(size_t)n * Tout + t0 + t] = Us[n * kLD + t] + sm * (yacc[i][j] + bn);
      }
    }
  }
}

// ------------------------------------------------------------------ packing
// Every section of the packed fp32 block is one JOB (a zero-padded vector copy or a transposed slice);
// the whole job table travels as a kernel parameter and ONE launch packs the block (was 40 launches).
struct PackJob {
  const float *src;
  float *dst;
  int transpose;            // 0: dst[e] = e < N ? src[e] (or `fill` when src == NULL) : 0, e < NPAD
                            // 1: dst[k][n] = n < N ? src[n * src_ld + src_off + k] : 0, k < K, n < NPAD
  int src_ld, src_off, N, K, NPAD;
  float fill;
};
constexpr int kMaxPackJobs = 48;
struct PackJobs {
  PackJob job[kMaxPackJobs];
  int n;
};
__global__ void pack_jobs_kernel(const __grid_constant__ PackJobs J) {
  const PackJob &j = J.job[blockIdx.y];
  const int total = j.transpose ? j.K * j.NPAD : j.NPAD;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    if (j.transpose) {
      const int k = e / j.NPAD, n = e % j.NPAD;
      j.dst[e] = n < j.N ? j.src[(size_t)n * j.src_ld + j.src_off + k] : 0.f;
    } else {
      j.dst[e] = e < j.N ? (j.src ? j.src[e] : j.fill) : 0.f;
    }
  }
}

int block_pack_fp32(const otp_block_params *p, int c, float *f, cudaStream_t st) {
  const BlockPackLayout L = block_pack_layout(c);
  const int np = npad_of(c);
  PackJobs J{};
  auto pack_T = [&](const float *src, int src_ld, int src_off, float *dst, int N, int K, int NPAD) {
    J.job[J.n++] = PackJob{src, dst, 1, src_ld, src_off, N, K, NPAD, 0.f};
  };
  auto pack_V = [&](const float *src, float *dst, int n, int npad, float fill) {
    J.job[J.n++] = PackJob{src, dst, 0, 0, 0, n, 0, npad, fill};
  };
  pack_T(p->q_w, c, 0, f + L.wqT, c, c, np);
  pack_T(p->k_w, c, 0, f + L.wkT, c, c, np);
  pack_T(p->proj_w, c, 0, f + L.wpT, c, c, np);
  pack_V(p->v_w, f + L.wv, c * c, c * c, 0.f);
  for (int q = 0; q < 4; ++q) {
    pack_T(p->mlp0_w + (size_t)q * c * c, c, 0, f + L.w1T + (size_t)q * c * np, c, c, np);
    pack_T(p->mlp3_w, 4 * c, q * c, f + L.w2T + (size_t)q * c * np, c, c, np);
    pack_V(p->mlp0_b + q * c, f + L.b1 + q * np, c, np, 0.f);
  }
  pack_V(p->ln1_w, f + L.ln1_w, c, c, 1.f); pack_V(p->ln1_b, f + L.ln1_b, c, c, 0.f);
  pack_V(p->ln2_w, f + L.ln2_w, c, c, 1.f); pack_V(p->ln2_b, f + L.ln2_b, c, c, 0.f);
  pack_V(p->q_norm_w, f + L.qn_w, c, c, 1.f); pack_V(p->q_norm_b, f + L.qn_b, c, c, 0.f);
  pack_V(p->k_norm_w, f + L.kn_w, c, c, 1.f); pack_V(p->k_norm_b, f + L.kn_b, c, c, 0.f);
  pack_V(p->v_norm_w, f + L.vn_w, c, c, 1.f); pack_V(p->v_norm_b, f + L.vn_b, c, c, 0.f);
  pack_V(p->q_conv_w, f + L.dwq, 3 * c, 3 * c, 0.f);
  pack_V(p->k_conv_w, f + L.dwk, 3 * c, 3 * c, 0.f);
  pack_V(p->v_conv_w, f + L.dwv, 3 * c, 3 * c, 0.f);
  pack_V(p->q_b, f + L.bq, c, np, 0.f); pack_V(p->k_b, f + L.bk, c, np, 0.f);
  pack_V(p->v_b, f + L.bv, c, np, 0.f); pack_V(p->proj_b, f + L.bp, c, np, 0.f);
  pack_V(p->mlp3_b, f + L.b2, c, np, 0.f);
  pack_V(p->scale_attn, f + L.sa, c, np, 1.f);
  pack_V(p->scale_mlp, f + L.sm, c, np, 1.f);
  if (J.n > kMaxPackJobs) return fail_arg("pack job table overflow");
  LaunchScope ls(K_PACK, st, 1);
  pack_jobs_kernel<<<dim3(ceil_div(c * np, 256), J.n), 256, 0, st>>>(J);
  return check_launch("block_pack");
}

// ------------------------------------------------------------------ host driver
template <int C_>
static int block_forward_simt_t(const void *packed, const float *x, float *y, int b, int t, int stride,
                                void *ws, cudaStream_t st) {
  using Cfg = BlockCfg<C_>;
  constexpr int C = Cfg::C;
  const BlockWorkspace W = block_workspace(b, C, t, Cfg::NH, stride);
  const BlockPack P = block_pack_view(packed, C);
  char *wsb = static_cast<char *>(ws);
  float *gram = reinterpret_cast<float *>(wsb + W.gram_part);
  float *weffT = reinterpret_cast<float *>(wsb + W.weffT);
  float *beff = reinterpret_cast<float *>(wsb + W.beff);
  float *obuf = reinterpret_cast<float *>(wsb + W.obuf);
  const int threads = Cfg::NWARP * 32;
  const int tiles = ceil_div(W.tout, kTT);
  const size_t red_bytes = (size_t)(Cfg::NWARP + 2) * kRedLd * 4;
  const size_t smem_front = (size_t)C * (kLDX + 3 * kLD) * 4 + red_bytes;
  const size_t smem_apply = (size_t)C * (kLDX + kLD) * 4 + red_bytes;
  const size_t smem_back = (size_t)C * 3 * kLD * 4 + red_bytes;
  static PerDeviceOnce attr;
  if (attr.first()) {
    if (!set_max_smem(block_front_kernel<C_>, smem_front, "block_front_kernel") ||
        !set_max_smem(block_apply_kernel<C_>, smem_apply, "block_apply_kernel") ||
        !set_max_smem(block_back_kernel<C_>, smem_back, "block_back_kernel"))
      return OTP_ERR_CUDA;
  }
  const float qscale = 1.0f / sqrtf((float)Cfg::HS);
  {
    LaunchScope ls(K_BLOCK_FRONT, st);
    block_front_kernel<C_><<<dim3(W.nchunk, b), threads, smem_front, st>>>(
        P, x, gram, t, W.tout, stride, W.tiles_per_chunk, W.nchunk, qscale);
  }
  {
    LaunchScope ls(K_BLOCK_FOLD, st);
    block_fold_kernel<C_, 0><<<dim3(b, FoldCfg<C_>::NBLK), kFoldThreads, 0, st>>>(P.wv, P.bv, gram, W.nchunk, weffT,
                                                                                 beff, C, Cfg::NPAD);
  }
  {
    LaunchScope ls(K_BLOCK_APPLY, st);
    block_apply_kernel<C_><<<dim3(tiles, b), threads, smem_apply, st>>>(P, x, weffT, beff, obuf, t, W.tout,
                                                                         stride);
  }
  {
    LaunchScope ls(K_BLOCK_BACK, st);
    block_back_kernel<C_><<<dim3(tiles, b), threads, smem_back, st>>>(P, x, obuf, y, t, W.tout, stride);
  }
  return check_launch("block_forward_simt");
}

int block_forward_small(const void *packed, const float *x, float *y, int b, int t, int stride, void *ws,
                        cudaStream_t st);   // block_small.cu: thread-per-token kernels for C = 17

int block_forward_simt(const void *packed, const float *x, float *y, int b, int c, int t, int stride,
                       void *ws, cudaStream_t st) {
  if (c == 136) return block_forward_simt_t<136>(packed, x, y, b, t, stride, ws, st);
  return block_forward_small(packed, x, y, b, t, stride, ws, st);
}

}  // namespace otp
