/*
 * otpose_b200 -- C ABI of the B200-native OTPose temporal fusion head.
 *
 * Plain C: device pointers, sizes, a CUDA stream handle.  No torch / ATen types.
 * Every entry point returns an otp_status (0 == OK), never throws, never
 * allocates device memory (scratch is passed in, sized by the matching
 * *_workspace_bytes query) and enqueues its kernels on the given stream.
 * otp_last_error() returns a thread-local description of the last failure.
 *
 * All tensors are contiguous fp32 in the reference's layouts (NCHW == (B,C,T)
 * with T = H*W) unless stated.  File:line citations are relative to the
 * reference tree (KyungMinJin/OTPose); each function names the reference
 * interface it replaces.
 */
#ifndef OTPOSE_B200_H_
#define OTPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *otp_stream_t; /* == cudaStream_t */

typedef enum {
  OTP_OK = 0,
  OTP_ERR_ARG = 1,         /* bad shape / null pointer / misaligned buffer          */
  OTP_ERR_UNSUPPORTED = 2, /* legal in the reference but not built here (see msg)   */
  OTP_ERR_WORKSPACE = 3,   /* workspace too small                                   */
  OTP_ERR_CUDA = 4         /* launch / runtime error (message has the CUDA string)  */
} otp_status;

typedef enum {
  OTP_PREC_FP32 = 0, /* fp32 CUDA-core arithmetic everywhere (<=1e-3 vs reference)      */
  OTP_PREC_BF16 = 1, /* bf16 operands on tcgen05 tensor cores, fp32 accumulate / LN /
                        softmax / residual stream                                     */
  OTP_PREC_FP16 = 2  /* same kernels with IEEE-half operands: same speed, 8x finer operand
                        rounding (<=2e-2 vs reference with margin)                    */
} otp_precision;

const char *otp_version(void);
const char *otp_last_error(void);
/* 1 if the current device is sm_100 (tcgen05 paths usable), 0 otherwise, <0 on error. */
int otp_device_is_sm100(void);
/* 1 if the tcgen05 (OTP_PREC_BF16) kernels are built into this library. */
int otp_has_tensor_core_path(void);

/* Launch accounting and optional per-kernel timing (CUDA events recorded on the
 * launching stream around every kernel of this library while enabled).
 *   otp_launch_count        kernels launched by this library since load
 *   otp_profile_enable(1)   clear records and start recording; (0) stop
 *   otp_profile_read        per kernel id: summed device ms and number of records;
 *                           arrays of n >= otp_profile_num_kernels() entries     */
unsigned long long otp_launch_count(void);
int otp_profile_num_kernels(void);
const char *otp_profile_kernel_name(int id);
int otp_profile_enable(int on);
int otp_profile_read(float *total_ms, int *launches, int n);

/* ---------------------------------------------------------------------------
 * a11  get_max_preds / get_final_preds / transform_preds
 * replaces utils/heatmap.py:108-171 + utils/transform.py:76-126 (NumPy on host).
 * heatmaps (N,J,H,W); center, scale (N,2) fp32 (only scale[:,0] is used, as in
 * the reference).  Outputs (all device, any may be NULL):
 *   out_idx     (N,J)   int32 flat argmax, FIRST maximum on ties (np.argmax)
 *   out_coords  (N,J,2) fp32 heat-map coords after the +-0.25 shift
 *   out_preds   (N,J,2) fp32 image coords (get_final_preds()[0])
 *   out_maxvals (N,J)   fp32 (get_final_preds()[1], shape (N,J,1))
 * ------------------------------------------------------------------------- */
int otp_final_preds(const float *heatmaps, int n, int j, int h, int w, const float *center,
                    const float *scale, int32_t *out_idx, float *out_coords, float *out_preds,
                    float *out_maxvals, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a9  modulated deformable convolution forward
 * replaces modulated_deform_conv_cuda_forward (thirdparty/deform_conv/src/
 * deform_conv_cuda.cpp:474-549) and its im2col kernel (deform_conv_cuda_kernel.cu
 * :402-432, 505-571) -- im2col, weight contraction and bias fused, batch in one
 * launch, no `columns`/`ones` scratch.
 *   x (B,C,H,W)  offset (B, dg*2*kh*kw, Ho, Wo)  mask (B, dg*kh*kw, Ho, Wo)
 *   weight (Cout, C, kh, kw)  bias (Cout) or NULL  out (B,Cout,Ho,Wo)
 * out = (accumulate ? out : 0) + alpha * (dcn(x) + bias); alpha/accumulate let
 * the caller fuse the reference's 0.2-weighted sum over dilations
 * (model/OTPose.py:387-392).  groups must be 1 (the only value OTPose uses).
 * ------------------------------------------------------------------------- */
int otp_mdcn_forward(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, int b, int c, int h, int w, int cout, int kh,
                     int kw, int stride, int pad, int dilation, int groups, int deformable_groups,
                     float alpha, int accumulate, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a12 (native part)  modulated deformable convolution backward
 * replaces modulated_deform_conv_cuda_backward (deform_conv_cuda.cpp:551-664)
 * and deform_conv_cuda_kernel.cu:434-503, 573-705.  All gradients are
 * overwritten (not accumulated).  grad_x is accumulated with 64-bit fixed-point
 * atomics (scale 2^30) and is therefore bit-reproducible, unlike the
 * reference's float atomicAdd; grad_weight/grad_bias are reduced in a fixed
 * order.  grad_bias may be NULL.  Built for Cout == 17, groups == 1.
 *   grad_out (B,Cout,Ho,Wo) -> grad_x (B,C,H,W), grad_offset / grad_mask shaped
 *   like offset / mask, grad_weight (Cout,C,kh,kw), grad_bias (Cout).
 * ------------------------------------------------------------------------- */
size_t otp_mdcn_backward_workspace_bytes(int b, int c, int h, int w, int cout, int kh, int kw,
                                         int stride, int pad, int dilation);
int otp_mdcn_backward(const float *x, const float *offset, const float *mask, const float *weight,
                      const float *grad_out, float *grad_x, float *grad_offset, float *grad_mask,
                      float *grad_weight, float *grad_bias, int b, int c, int h, int w, int cout,
                      int kh, int kw, int stride, int pad, int dilation, int groups,
                      int deformable_groups, void *workspace, size_t workspace_bytes,
                      otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a1  fusion prologue, replaces model/OTPose.py:320-330 and :339-359.
 * rough (5B,J,T) ordered cur, prev, next, pprev, nnext.
 *   otp_fusion_sum    -> total_b (B,J,T), squeezed (B,T)  [the reference's
 *                        `squeezed` is this plane repeated J times]
 *   otp_fusion_stack  -> x1, x2 (B, 8J, T) (channel j*8+m, OTPose.py:356-359)
 *                        with the encoders' positional embedding already added
 *                        (ConvVideoTransformer.py:140-155; pe1/pe2 (8J, >=T) with
 *                        row stride pe_stride, NULL = no PE), intersection and
 *                        prev_b (B,J,T).  margin (B,4) int64.
 * ------------------------------------------------------------------------- */
int otp_fusion_sum(const float *rough, int b, int j, int t, float *total_b, float *squeezed,
                   otp_stream_t stream);
int otp_fusion_stack(const float *rough, const int64_t *margin, const float *squeezed,
                     const float *context, const float *pe1, const float *pe2, int pe_stride, int b,
                     int j, int t, float *x1, float *x2, float *intersection, float *prev_b,
                     otp_stream_t stream);
/* Window extension (BASELINE config 5, T = 3 / 5 / 7 frames): `frames` = 2*NP+1,
 * rough (frames*B,J,T) ordered cur, prev1, next1, prev2, next2, ..., margin (B, 2*NP)
 * in the same pair order.  prev_b / next_b sum every past / future frame,
 * close_b the nearest pair, far_b the remaining pairs (far_b = cur at frames == 3).
 * The reference has no code path for frames != 5 (`supplement = 5`,
 * model/OTPose.py:188, 317-321); frames == 5 is exactly otp_fusion_sum / _stack. */
int otp_fusion_sum_frames(const float *rough, int frames, int b, int j, int t, float *total_b,
                          float *squeezed, otp_stream_t stream);
int otp_fusion_stack_frames(const float *rough, const int64_t *margin, const float *squeezed,
                            const float *context, const float *pe1, const float *pe2, int pe_stride,
                            int frames, int b, int j, int t, float *x1, float *x2,
                            float *intersection, float *prev_b, otp_stream_t stream);

/* a0 + a1 fused (backbone -> head hand-off): HRNet.final_layer -- the 1x1 conv
 * Cin -> J the backbone ends with (model/HRNet.py:108-114, 150; called at model/OTPose.py:319) --
 * and the frame sum of model/OTPose.py:324-326 in one pass over the backbone's last feature map.
 *   features   (frames*B, Cin, H, W), frames ordered cur, prev1, next1, ... (OTPose.py:317-321);
 *              feat_dtype: OTP_PREC_FP32 / _BF16 / _FP16 element type; channels_last = 0: NCHW
 *              contiguous (what the reference's backbone returns), 1: NHWC contiguous (a
 *              channels-last cuDNN backbone); 16-byte aligned, Cin % 8 == 0, Cin <= 64
 *   weight     (J, Cin) fp32 = final_layer.weight viewed 2-D (FINAL_CONV_KERNEL = 1), bias (J) or NULL
 *   rough      (frames*B, J, T) fp32 out = the rough_heatmaps OTPose.forward returns
 *   total_b    (B, J, T), squeezed (B, T) fp32 out, bit-identical to otp_fusion_sum_frames(rough)
 * fp32 accumulation whatever the element type. */
int otp_final_layer_fusion_sum(const void *features, int feat_dtype, int channels_last, const float *weight,
                               const float *bias, int frames, int b, int cin, int joints, int t, float *rough,
                               float *total_b, float *squeezed, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a2-a5  ConvTransformer building blocks, replaces model/blocks.py:95-110
 * (LayerNorm), :264-279 (TransformerBlock.forward), :400-452 (MaskedMHCA.forward),
 * :289-298 (AffineDropPath, eval) and ConvVideoTransformer.py:140-179.
 * ------------------------------------------------------------------------- */
typedef struct {
  /* reference state-dict entries of one TransformerBlock, device fp32 */
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;                   /* (1,C,1)            */
  const float *q_conv_w, *k_conv_w, *v_conv_w;                  /* (C,1,3) depthwise   */
  const float *q_norm_w, *q_norm_b, *k_norm_w, *k_norm_b, *v_norm_w, *v_norm_b;
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *proj_w, *proj_b; /* (C,C,1), (C)    */
  const float *mlp0_w, *mlp0_b, *mlp3_w, *mlp3_b;               /* (4C,C,1),(4C),(C,4C,1),(C) */
  const float *scale_attn, *scale_mlp;                          /* (1,C,1) or NULL (=1) */
} otp_block_params;

/* Packed (transposed / padded / bf16) copy of one block's weights. */
size_t otp_block_packed_bytes(int c, int n_head);
int otp_block_pack(const otp_block_params *p, int c, int n_head, void *packed, size_t packed_bytes,
                   otp_stream_t stream);

/* y (B,C,T') = TransformerBlock(x (B,C,T)), T' = T for stride 1, (T-1)/2+1 for
 * stride 2 (depthwise k=3 pad=1 stride 2 + MaxPool1d(3,2,1) skip).  Eval-mode
 * semantics (dropout / drop-path inactive).  x and y must not alias. */
size_t otp_block_workspace_bytes(int b, int c, int t, int n_head, int stride, int precision);
int otp_block_forward(const void *packed, const float *x, float *y, int b, int c, int t, int n_head,
                      int stride, int precision, void *workspace, size_t workspace_bytes,
                      otp_stream_t stream);

/* a2-a5, the whole C = 17 / one-head flow encoder (model/OTPose.py:214-216, 331-335;
 * model/ConvVideoTransformer.py:147-170) in the 16-bit modes: y (B,17,T) = stem blocks (x + pe[:, :t]), all
 * `nblocks` stride-1 TransformerBlocks in ONE launch -- a thread-block cluster per clip keeps the residual
 * stream in shared memory, the channel Gram is reduced through distributed shared memory in a fixed order
 * (IEEE half mma.sync operands, fp32 accumulate / LayerNorms / softmax).  packed_blocks[i] = otp_block_pack of
 * block i (HOST array of device pointers); pe may be NULL.  `supported` = 1 when (c, n_head, t, nblocks) is
 * built (c = 17, n_head = 1, t <= 13824, nblocks <= 8); otherwise call otp_block_forward per block. */
int otp_flow_encoder_supported(int c, int n_head, int t, int nblocks);
size_t otp_flow_encoder_workspace_bytes(int b, int t);
int otp_flow_encoder_forward(const void *const *packed_blocks, int nblocks, const float *x, const float *pe,
                             int pe_stride, float *y, int b, int t, void *workspace, size_t workspace_bytes,
                             otp_stream_t stream);

/* y = x + pe[:, :t]   (x,y (B,C,T); pe (C, >=T), row stride pe_stride) */
int otp_add_pos_embd(const float *x, const float *pe, int pe_stride, float *y, int b, int c, int t,
                     otp_stream_t stream);
/* nn.Upsample(scale_factor=scale, mode='linear'), align_corners=False:
 * x (B,C,Tin) -> y (B,C,Tin*scale)   (ConvVideoTransformer.py:108, 179) */
int otp_upsample_linear(const float *x, float *y, int b, int c, int t_in, int scale,
                        otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a6  stack/view + final_layer (1x1 conv over the 3-scale pyramid), replaces
 * model/OTPose.py:362-373.  s0 (B,C,T), s1 (B,C,T1), s2 (B,C,T2) are the raw
 * branch outputs; the x2 / x4 linear upsampling of s1 / s2 is applied on the fly.
 * weight (Cout, 3C), bias (Cout); out has batch stride out_bstride elements so
 * it can be a channel slice of the concatenated `branches` buffer.
 * ------------------------------------------------------------------------- */
int otp_pyramid_conv1x1(const float *s0, const float *s1, const float *s2, int b, int c, int t,
                        int t1, int t2, const float *weight, const float *bias, int cout,
                        float *out, long long out_bstride, otp_stream_t stream);

/* a6, 16-bit tensor-core variant (t % 8 == 0, cout <= 32): the upsampled / stacked operand tile is
 * built in shared memory in 16 bit and contracted by tcgen05 UMMAs with the packed weight image
 * (otp_pyramid_conv1x1_tc_pack of the (cout, 3c) fp32 weight for the same precision). */
int otp_pyramid_conv1x1_tc_supported(int c, int t, int cout);
size_t otp_pyramid_conv1x1_tc_pack_bytes(int c);
int otp_pyramid_conv1x1_tc_pack(const float *weight, int c, int cout, int precision, void *packed,
                                size_t packed_bytes, otp_stream_t stream);
int otp_pyramid_conv1x1_tc(const float *s0, const float *s1, const float *s2, int b, int c, int t,
                           const void *packed, const float *bias, int cout, float *out,
                           long long out_bstride, int precision, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a7/a8  small-channel Conv2d (stride 1, square kernel k in {1,3}, dilation d,
 * padding d*(k/2)), replaces the nn.Conv2d (+folded eval BatchNorm + ReLU) calls
 * of model/RSB.py:106-139 and the offset/mask convs of model/OTPose.py:168-177.
 *   y = act( conv(x [+ x_add]) + bias [+ residual] )
 * x, x_add, residual, y are channel slices: (ptr, batch stride in elements).
 * ------------------------------------------------------------------------- */
int otp_conv2d(const float *x, long long x_bstride, const float *x_add, long long x_add_bstride,
               const float *weight, const float *bias, const float *residual,
               long long residual_bstride, float *y, long long y_bstride, int b, int cin, int h,
               int w, int cout, int k, int dilation, int relu, otp_stream_t stream);

/* a12 (training config): backward of otp_conv2d for the weights and the bias -- what autograd runs through
 * ATen's convolution backward in the reference for the offset / mask convs of model/OTPose.py:168-177 when
 * ModulatedDeformConvFunction.backward (functions/deform_conv.py:148-167) hands their gradients back.
 *   grad_weight (cout, cin, k, k) = sum_{b,h,w} grad_out[b,o,h,w] * x[b,c,h+(i-k/2)d, w+(j-k/2)d]
 *   grad_bias   (cout) or NULL    = sum_{b,h,w} grad_out[b,o,h,w]
 * x / grad_out are channel slices (ptr, batch stride in elements) like otp_conv2d; accumulate != 0 adds to
 * the existing gradients.  Deterministic (fixed-order reduction of per-slice partials in `workspace`).
 * grad_input needs no entry point: it is otp_conv2d(grad_out) with the weights transposed and flipped. */
size_t otp_conv2d_wgrad_workspace_bytes(int b, int cin, int h, int w, int cout, int k);
int otp_conv2d_wgrad(const float *x, long long x_bstride, const float *grad_out, long long go_bstride,
                     float *grad_weight, float *grad_bias, int b, int cin, int h, int w, int cout, int k,
                     int dilation, int accumulate, void *workspace, size_t workspace_bytes, otp_stream_t stream);

/* a7 weight packing: eval-mode BatchNorm folded into the convolution in front of it
 * (model/RSB.py:106-139: y = bn(conv(x))):  w_out[o] = w[o] * g_o,  b_out[o] = (b[o] - mean[o]) * g_o + beta[o],
 * g_o = gamma[o] / sqrt(var[o] + eps).  `per_out` = cin * k * k weights per output channel.  With has_bn == 0
 * the weights and the bias are copied (bias may be NULL: zeros). */
int otp_conv_bn_fold(const float *weight, const float *bias, const float *gamma, const float *beta,
                     const float *running_mean, const float *running_var, float eps, int has_bn, int cout,
                     int per_out, float *weight_out, float *bias_out, otp_stream_t stream);

/* f4: input window assembly, replaces dataset/PoseTrackDataset.py:389-406 (cv2.warpAffine of the
 * clip's frames with the get_affine_transform matrix, ToTensor, Normalize) and the torch.cat of script/Common.py:347.
 * frames: device uint8 (n_frames, src_h, src_w, 3) as cv2.imread returns them (frame_stride bytes apart);
 * frame_index: device int32 (B, frames_per_clip) rows of `frames` in the order cur, prev, next, pprev, nnext;
 * trans: device double (B, 6), the FORWARD 2x3 matrix of get_affine_transform(center, scale, 0, (out_w, out_h));
 * swap_rb = 1: cv2.cvtColor(BGR2RGB) first (cfg color_rgb); mean3 / std3: HOST float[3].
 * out: (B, 3 * frames_per_clip, out_h, out_w) fp32 = concat_input (may be NULL); out_bf16_nhwc: the same pixels as the
 * (frames_per_clip * B, out_h, out_w, 3) bfloat16 channels-last batch of model/OTPose.py:317 (may be NULL).
 * The warp is bit-exact with cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0); the fp32 output bit-exact with torchvision. */
int otp_window_assemble(const unsigned char *frames, int n_frames, int src_h, int src_w, long long frame_stride,
                        const int *frame_index, int frames_per_clip, const double *trans, int b, int out_h, int out_w,
                        int swap_rb, const float *mean3, const float *std3, float *out, void *out_bf16_nhwc,
                        otp_stream_t stream);

/* a7, 16-bit modes: a whole RSB_BLOCK (model/RSB.py:26-103: conv_bn_relu1, the ten dense-connected 3x3
 * conv_bn_relu2_* convs, conv_bn_relu3 + skip / downsample + ReLU) in one launch -- two for the widest block -- with
 * every intermediate map in shared-memory row rings (csrc/rsb_fused.cu).  `weights` / `biases`: HOST arrays of 13
 * device pointers to the BN-FOLDED fp32 convs (otp_conv_bn_fold) in the order conv_bn_relu1, conv_bn_relu2_{1_1,
 * 2_1, 2_2, 3_1, 3_2, 3_3, 4_1, 4_2, 4_3, 4_4}, conv_bn_relu3, downsample (entry 12 ignored without downsample).
 * x (B,cin,H,W) / y (B,planes,H,W) fp32 with batch strides in elements (channel slices of wider buffers). */
int otp_rsb_block_supported(int cin, int planes, int h, int w);
size_t otp_rsb_block_pack_bytes(int cin, int planes, int has_downsample);
int otp_rsb_block_pack(const float *const *weights, const float *const *biases, int cin, int planes,
                       int has_downsample, void *packed, size_t packed_bytes, otp_stream_t stream);
size_t otp_rsb_block_workspace_bytes(int b, int cin, int planes, int h, int w);
int otp_rsb_block_forward(const void *packed, const float *x, long long x_bstride, float *y, long long y_bstride,
                          int b, int cin, int planes, int has_downsample, int h, int w, void *workspace,
                          size_t workspace_bytes, otp_stream_t stream);

/* a7, 16-bit tensor-core variant of otp_conv2d (dilation 1, w % 8 == 0, <= 96 channels):
 * implicit GEMM on tcgen05 over three pre-shifted 16-bit copies of the input rows (no im2col
 * tile), fp32 accumulate / bias / residual / output.  `packed` = otp_conv2d_tc_pack of the
 * (BatchNorm-folded) fp32 weight for the same precision (OTP_PREC_FP16 / OTP_PREC_BF16). */
int otp_conv2d_tc_supported(int cin, int cout, int h, int w, int k); /* 1 if otp_conv2d_tc handles the shape */
size_t otp_conv2d_tc_pack_bytes(int cin, int cout, int k);
int otp_conv2d_tc_pack(const float *weight, int cin, int cout, int k, int precision, void *packed,
                       size_t packed_bytes, otp_stream_t stream);
int otp_conv2d_tc(const float *x, long long x_bstride, const float *x_add, long long x_add_bstride,
                  const void *packed, const float *bias, const float *residual,
                  long long residual_bstride, float *y, long long y_bstride, int b, int cin, int h,
                  int w, int cout, int k, int relu, int precision, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * a8 + a9 + a10 fused (tensor-core precisions only): for one dilation d,
 *   offsets = conv3x3_d(trans; w_off), masks = conv3x3_d(trans; w_msk),
 *   out = (accumulate ? out : 0) + alpha * (modulated_dcn_d(x; offsets, masks; dcn_w) + dcn_b)
 * replaces model/OTPose.py:381-392 for that dilation (two nn.Conv2d + the DCN op +
 * the weighted sum) without materialising offsets / masks.  Built for the
 * reference shapes: 17 joints (deformable groups), 32 feature channels, 3x3 taps.
 *   w_off (306,32,3,3), w_msk (153,32,3,3) are packed once per dilation by
 *   otp_offset_mask_pack into `packed` (otp_offset_mask_pack_bytes() bytes);
 *   trans (B,32,H,W), x (B,17,H,W), dcn_w (17,17,3,3), dcn_b (17) or NULL, out (B,17,H,W).
 * ------------------------------------------------------------------------- */
size_t otp_offset_mask_pack_bytes(void);
int otp_offset_mask_pack(const float *w_off, const float *w_msk, int joints, int cin, void *packed,
                         size_t packed_bytes, otp_stream_t stream);
int otp_offset_mask_dcn_forward(const void *packed, const float *trans, const float *x, const float *dcn_w,
                                const float *dcn_b, float *out, int b, int h, int w, int dilation, float alpha,
                                int accumulate, int precision, otp_stream_t stream);

/* ---------------------------------------------------------------------------
 * Phase trace of the tensor-core block kernels (no reference counterpart; used
 * to write profiles/): while enabled, CTA 0 of tc_back records
 * (clock64 << 8 | event id) for epilogue warp 0 (row 0), the control warp
 * (row 1) and epilogue warp 15 (row 2) into a 3 x 2048 device array;
 * otp_debug_trace_read copies it to the host array `out` of n >= 6144 entries
 * (synchronises the device).
 * ------------------------------------------------------------------------- */
int otp_debug_trace(int on);
/* UMMA rate probe: `ctas` CTAs each issue reps x ksteps M128 x n x K16 UMMAs; cycles_host gets
 * 2 values per CTA: clock64 cycles of the issue loop, and of issue + completion. */
int otp_debug_umma_rate(int n, int ksteps, int reps, int ctas, long long *cycles_host);
int otp_debug_trace_read(unsigned long long *out, int n);

/* ---------------------------------------------------------------------------
 * Self-test of the tcgen05 / TMEM plumbing (no reference counterpart): one CTA
 * computes D[128,n] (fp32, row-major) = A . B^T over `ksteps` K=16 steps from
 * caller-built bf16 shared-memory operand images and explicit UMMA descriptor
 * fields (byte offsets; *_mn_major selects the MN-major operand view).
 * ------------------------------------------------------------------------- */
int otp_debug_umma_gemm(const void *a_img, int a_bytes, const void *b_img, int b_bytes, float *d, int n,
                        int ksteps, unsigned a_off, unsigned a_lbo, unsigned a_sbo, unsigned a_kstep,
                        unsigned b_off, unsigned b_lbo, unsigned b_sbo, unsigned b_kstep, int a_mn_major,
                        int b_mn_major, int repeat, otp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OTPOSE_B200_H_ */
