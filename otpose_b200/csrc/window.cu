// f4: the reference's per-clip INPUT WINDOW ASSEMBLY on the GPU -- the five frames of a
// person-clip cropped by one affine warp, normalised and concatenated:
//
//   dataset/PoseTrackDataset.py:389-406   trans = get_affine_transform(center, scale, 0, image_size)
//                                         input_f = transform(cv2.warpAffine(frame_f, trans, (W, H), INTER_LINEAR))
//   utils/transform.py:6-17               transform = ToTensor() -> Normalize(mean, std)
//   script/Common.py:347                  concat_input = cat((x, prev, next, pprev, nnext), 1)      (B, 15, H, W)
//
// so that what crosses PCIe is the uint8 video frames (shared by every person in them) + 48 bytes of transform
// per clip instead of 6.6 MB of fp32 crops per clip.
//
// BIT-EXACT with cv2.warpAffine (OpenCV imgproc/src/imgwarp.cpp; not under /root/reference: opencv-python==4.4.0.44,
// restated in oracle/window_oracle.py and pinned by tests/golden/window_*.npz): the matrix is inverted in double with
// the same operation order (no FMA contraction: __dmul_rn / __dadd_rn), the source position is the same FIXED-POINT
// number (x 1024, + 16, >> 5: 1/32 pixel), the bilinear weights are the same 15-bit integers
// ((32 - fy)(32 - fx) 32, ...), the pixel is (sum + 2^14) >> 15, BORDER_CONSTANT 0.  ToTensor / Normalize are the same
// three fp32 operations (/ 255, - mean, / std; IEEE division), tabulated per CTA for the 256 possible inputs.  One thread = one output pixel of one clip: the
// fixed-point position is computed once and used for all five frames and three channels; fp32 NCHW stores are
// coalesced along x; an optional second output is the 16-bit channels-last (5B, H, W, 3) batch the backbone
// drop-in consumes (model/OTPose.py:317: cat(x.split(3, dim=1), 0) -> frame-major).
#include <cuda_bf16.h>

#include "common.cuh"

namespace otp {
namespace {

struct WindowArgs {
  const unsigned char *frames;
  long long frame_stride;
  int n_frames, src_h, src_w;
  const int *frame_index;
  const double *trans;
  int b, out_h, out_w, swap_rb, nf;
  float mean[3], stdv[3];
  float *out;
  __nv_bfloat16 *out16;
};

constexpr int kWinRows = 8;   // output rows per CTA: amortises the normalisation table and the per-column constants

template <int NF>
__global__ void __launch_bounds__(128) window_kernel(const WindowArgs A) {
  // ToTensor + Normalize of a uint8 has 256 possible results per channel: the three IEEE operations are evaluated
  // once per table entry (bit-exact by construction) instead of two fp32 divisions per output value
  __shared__ float lut[3][256];
  for (int e = threadIdx.x; e < 3 * 256; e += blockDim.x) {
    const int c = e >> 8, u = e & 255;
    const float mean = c == 0 ? A.mean[0] : (c == 1 ? A.mean[1] : A.mean[2]);
    const float sd = c == 0 ? A.stdv[0] : (c == 1 ? A.stdv[1] : A.stdv[2]);
    lut[c][u] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), mean), sd);
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
  if (x >= A.out_w) return;
  // ---- cv::warpAffine: invert the 2x3 matrix (double, same operation order as imgwarp.cpp) ----
  const double *T = A.trans + (size_t)b * 6;
  double m0 = T[0], m1 = T[1], m2 = T[2], m3 = T[3], m4 = T[4], m5 = T[5];
  double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
  D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
  const double a11 = __dmul_rn(m4, D), a22 = __dmul_rn(m0, D);
  m0 = a11;
  m1 = __dmul_rn(m1, -D);
  m3 = __dmul_rn(m3, -D);
  m4 = a22;
  const double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
  const double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
  // ---- fixed-point source position: AB_BITS = 10, INTER_BITS = 5, round_delta = 16 ----
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m0, (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m3, (double)x), 1024.0));
  const size_t P = (size_t)A.out_h * A.out_w;
  const int ci0 = A.swap_rb ? 2 : 0, ci2 = A.swap_rb ? 0 : 2;   // cv2.cvtColor(BGR2RGB) before the warp == channel choice after it
  const int y_end = min(A.out_h, (int)(blockIdx.y + 1) * kWinRows);
  for (int y = blockIdx.y * kWinRows; y < y_end; ++y) {
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m1, (double)y), b1), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m4, (double)y), b2), 1024.0)) + 16;
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = min(max(X >> 5, -32768), 32767), sy = min(max(Y >> 5, -32768), 32767);   // saturate_cast<short>
    const int fx = X & 31, fy = Y & 31;
    const bool x0 = sx >= 0 && sx < A.src_w, x1 = sx + 1 >= 0 && sx + 1 < A.src_w;
    const bool y0 = sy >= 0 && sy < A.src_h, y1 = sy + 1 >= 0 && sy + 1 < A.src_h;
    // border taps (BORDER_CONSTANT 0) get weight 0 and a clamped address
    const int w00 = (y0 && x0) ? (32 - fy) * (32 - fx) * 32 : 0, w01 = (y0 && x1) ? (32 - fy) * fx * 32 : 0;
    const int w10 = (y1 && x0) ? fy * (32 - fx) * 32 : 0, w11 = (y1 && x1) ? fy * fx * 32 : 0;
    const size_t o0 = ((size_t)(y0 ? sy : 0) * A.src_w) * 3, o1 = ((size_t)(y1 ? sy + 1 : 0) * A.src_w) * 3;
    const int c0 = (x0 ? sx : 0) * 3, c1 = (x1 ? sx + 1 : 0) * 3;
    const size_t px = (size_t)y * A.out_w + x;
    // all taps of all NF frames in flight (12 byte loads per frame) before the first use
    int t[NF][4][3];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      const int fi = A.frame_index[b * NF + f];
      const unsigned char *src = A.frames + (size_t)min(max(fi, 0), A.n_frames - 1) * A.frame_stride;
      const unsigned char *r0 = src + o0, *r1 = src + o1;
#pragma unroll
      for (int c = 0; c < 3; ++c) t[f][0][c] = r0[c0 + c], t[f][1][c] = r0[c1 + c], t[f][2][c] = r1[c0 + c], t[f][3][c] = r1[c1 + c];
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      float v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int sc = c == 0 ? ci0 : (c == 1 ? 1 : ci2);
        const int s0 = sc == 0 ? t[f][0][0] : (sc == 1 ? t[f][0][1] : t[f][0][2]), s1 = sc == 0 ? t[f][1][0] : (sc == 1 ? t[f][1][1] : t[f][1][2]);
        const int s2 = sc == 0 ? t[f][2][0] : (sc == 1 ? t[f][2][1] : t[f][2][2]), s3 = sc == 0 ? t[f][3][0] : (sc == 1 ? t[f][3][1] : t[f][3][2]);
        const int u = (s0 * w00 + s1 * w01 + s2 * w10 + s3 * w11 + (1 << 14)) >> 15;
        v[c] = lut[c][u];   // ToTensor: uint8 -> float32 / 255;  Normalize: (x - mean) / std
        if (A.out) A.out[((size_t)b * (3 * NF) + 3 * f + c) * P + px] = v[c];
      }
      if (A.out16) {
        __nv_bfloat16 *d = A.out16 + (((size_t)f * A.b + b) * P + px) * 3;
        d[0] = __float2bfloat16_rn(v[0]);
        d[1] = __float2bfloat16_rn(v[1]);
        d[2] = __float2bfloat16_rn(v[2]);
      }
    }
  }
}

}  // namespace
}  // namespace otp

using namespace otp;

extern "C" int otp_window_assemble(const unsigned char *frames, int n_frames, int src_h, int src_w,
                                   long long frame_stride, const int *frame_index, int frames_per_clip,
                                   const double *trans, int b, int out_h, int out_w, int swap_rb, const float *mean3,
                                   const float *std3, float *out, void *out_bf16_nhwc, otp_stream_t stream) {
  OTP_REQUIRE(b >= 0 && n_frames > 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0);
  OTP_REQUIRE((frames_per_clip == 3 || frames_per_clip == 5 || frames_per_clip == 7) && out_h <= 65535 && b <= 65535);
  OTP_REQUIRE(src_h <= 32767 && src_w <= 32767 && frame_stride >= (long long)src_h * src_w * 3);
  OTP_REQUIRE(mean3 != nullptr && std3 != nullptr);
  if (b == 0) return OTP_OK;
  OTP_REQUIRE(frames && frame_index && trans && (out || out_bf16_nhwc));
  WindowArgs A{};
  A.frames = frames, A.frame_stride = frame_stride, A.n_frames = n_frames, A.src_h = src_h, A.src_w = src_w;
  A.frame_index = frame_index, A.trans = trans, A.b = b, A.out_h = out_h, A.out_w = out_w, A.swap_rb = swap_rb;
  A.nf = frames_per_clip;
  for (int c = 0; c < 3; ++c) A.mean[c] = mean3[c], A.stdv[c] = std3[c];
  A.out = out, A.out16 = static_cast<__nv_bfloat16 *>(out_bf16_nhwc);
  cudaStream_t st = (cudaStream_t)stream;
  LaunchScope ls(K_WINDOW, st);
  // threads per CTA: the candidate that pads the row least (288 = 3 x 96)
  int bs = 128;
  for (int c : {96, 64})
    if (ceil_div(out_w, c) * c < ceil_div(out_w, bs) * bs) bs = c;
  const dim3 grid(ceil_div(out_w, bs), ceil_div(out_h, kWinRows), b);
  if (frames_per_clip == 5) window_kernel<5><<<grid, bs, 0, st>>>(A);
  else if (frames_per_clip == 3) window_kernel<3><<<grid, bs, 0, st>>>(A);
  else window_kernel<7><<<grid, bs, 0, st>>>(A);
  return check_launch("window_kernel");
}
