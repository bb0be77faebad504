"""CPU restatement (torch fp32 + numpy) of the OTPose temporal fusion head.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function takes a flat
``state_dict`` (reference key names -> torch tensors) and plain tensors, so it
can be driven by the same weights the CUDA drop-in modules hold.  All file:line
citations are relative to the reference tree (KyungMinJin/OTPose).

Third-party arithmetic that is not in the reference tree (conv1d/conv2d, bmm,
softmax, erf-GELU, max_pool1d, linear interpolation, batch-norm in eval mode)
is PyTorch ATen (reference pins torch==1.7.0, requirements.txt:57); this file
calls the same ATen ops on CPU.  The deformable convolution is restated
literally from the reference CUDA kernel (``mdcn_forward_literal``) and
cross-checked against ``torchvision.ops.deform_conv2d`` in the tests.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# model/blocks.py
# --------------------------------------------------------------------------


def layer_norm_ct(x, weight, bias, eps=1e-5):
    """Channel LayerNorm on (B, C, T).  model/blocks.py:95-110.

    mean over dim=1, biased variance, (x-mu)/sqrt(var+eps) * w + b with
    w, b shaped (1, C, 1).
    """
    mu = torch.mean(x, dim=1, keepdim=True)
    res = x - mu
    sigma = torch.mean(res ** 2, dim=1, keepdim=True)
    out = res / torch.sqrt(sigma + eps)
    return out * weight + bias


def sinusoid_encoding(n_position, d_hid):
    """model/blocks.py:114-125 -- float64 numpy table cast to fp32, (1, C, T)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table.astype(np.float32)).unsqueeze(0).transpose(1, 2).contiguous()


def masked_mhca(sd, p, x, n_head, stride):
    """MaskedMHCA.forward, model/blocks.py:400-452 (eval: dropouts are identity).

    ``p`` is the key prefix of the attention module (e.g. ``stem.0.attn.``).
    The attention is a channel x channel Gram matrix per head (q, k, v are
    viewed (B, nh, hs, T) and NOT transposed, blocks.py:427-440) and the head
    re-assembly is the ``transpose(2,3).contiguous().view(B, C, -1)`` scramble
    (blocks.py:447).
    """
    B, C, T = x.shape
    hs = C // n_head
    scale = 1.0 / math.sqrt(hs)

    def dw_ln_pw(name):
        y = F.conv1d(x, sd[p + f"{name}_conv.weight"], None, stride=stride, padding=1, groups=C)
        y = layer_norm_ct(y, sd[p + f"{name}_norm.weight"], sd[p + f"{name}_norm.bias"])
        return F.conv1d(y, sd[p + f"{name}.weight"], sd[p + f"{name}.bias"])

    q = dw_ln_pw("query")
    k = dw_ln_pw("key")
    v = dw_ln_pw("value")
    k = k.view(B, n_head, hs, -1)
    q = q.view(B, n_head, hs, -1)
    v = v.view(B, n_head, hs, -1)
    att = (q * scale) @ k.transpose(-2, -1)
    att = F.softmax(att, dim=-1)
    out = att @ v
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    return F.conv1d(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def transformer_block(sd, p, x, n_head, stride):
    """TransformerBlock.forward, model/blocks.py:264-279 (eval mode).

    pool_skip = Identity | MaxPool1d(3, 2, 1) (blocks.py:234-240); AffineDropPath
    in eval is a per-channel scale (blocks.py:289-298); MLP = Conv1d(C,4C,1) ->
    GELU(erf) -> Conv1d(4C,C,1) (blocks.py:248-254).
    """
    h = layer_norm_ct(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"])
    out = masked_mhca(sd, p + "attn.", h, n_head, stride)
    skip = x if stride == 1 else F.max_pool1d(x, stride + 1, stride=stride, padding=(stride + 1) // 2)
    sa = sd.get(p + "drop_path_attn.scale")
    sm = sd.get(p + "drop_path_mlp.scale")
    out = skip + (out if sa is None else sa * out)
    h = layer_norm_ct(out, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    h = F.conv1d(h, sd[p + "mlp.0.weight"], sd[p + "mlp.0.bias"])
    h = F.gelu(h)
    h = F.conv1d(h, sd[p + "mlp.3.weight"], sd[p + "mlp.3.bias"])
    return out + (h if sm is None else sm * h)


# --------------------------------------------------------------------------
# model/ConvVideoTransformer.py
# --------------------------------------------------------------------------


def conv_transformer(sd, p, x, n_head, arch, max_len=None, training=False):
    """ConvTransformer.forward, model/ConvVideoTransformer.py:123-184.

    arch[0] (conv embedding) must be 0 -- the only value OTPose uses
    (model/OTPose.py:201-202).  Returns a tuple of 1 + arch[2] (B, C, T) maps.
    """
    assert arch[0] == 0
    B, C, H, W = x.shape
    T = H * W
    x = x.flatten(2)
    pe = sd[p + "pos_embd"]
    max_len = pe.shape[-1] if max_len is None else max_len
    if training:
        assert T <= max_len, "Reached max length."
    elif T >= max_len:
        pe = F.interpolate(pe, T, mode="linear", align_corners=False)
    x = x + pe[:, :, :T]
    for i in range(arch[1]):
        x = transformer_block(sd, f"{p}stem.{i}.", x, n_head, 1)
    outs = (x,)
    for i in range(arch[2]):
        x = transformer_block(sd, f"{p}branch.{i}.", x, n_head, 2)
        outs += (F.interpolate(x, scale_factor=float(2 ** (i + 1)), mode="linear"),)
    return outs


# --------------------------------------------------------------------------
# model/RSB.py
# --------------------------------------------------------------------------


def conv_bn_relu(sd, p, x, padding, has_relu):
    """conv_bn_relu.forward, model/RSB.py:106-139, BatchNorm in eval mode."""
    x = F.conv2d(x, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=padding)
    x = F.batch_norm(x, sd[p + "bn.running_mean"], sd[p + "bn.running_var"],
                     sd[p + "bn.weight"], sd[p + "bn.bias"], False, 0.1, 1e-5)
    return F.relu(x) if has_relu else x


def rsb_block(sd, p, x, has_downsample):
    """RSB_BLOCK.forward, model/RSB.py:81-103."""
    out = conv_bn_relu(sd, p + "conv_bn_relu1.", x, 0, True)
    bc = out.shape[1] // 4
    spx = torch.split(out, bc, 1)

    def c(name, t):
        return conv_bn_relu(sd, f"{p}conv_bn_relu2_{name}.", t, 1, True)

    o11 = c("1_1", spx[0])
    o21 = c("2_1", spx[1] + o11)
    o22 = c("2_2", o21)
    o31 = c("3_1", spx[2] + o21)
    o32 = c("3_2", o31 + o22)
    o33 = c("3_3", o32)
    o41 = c("4_1", spx[3] + o31)
    o42 = c("4_2", o41 + o32)
    o43 = c("4_3", o42 + o33)
    o44 = c("4_4", o43)
    out = torch.cat((o11, o22, o33, o44), 1)
    out = conv_bn_relu(sd, p + "conv_bn_relu3.", out, 0, False)
    if has_downsample:
        x = conv_bn_relu(sd, p + "downsample.", x, 0, False)
    return F.relu(out + x)


def chain_rsb(sd, p, x, num_blocks):
    """CHAIN_RSB_BLOCKS.forward, model/RSB.py:10-23 (block 0 has the 1x1 downsample)."""
    for i in range(num_blocks):
        x = rsb_block(sd, f"{p}layers.{i}.", x, has_downsample=(i == 0))
    return x


# --------------------------------------------------------------------------
# thirdparty/deform_conv
# --------------------------------------------------------------------------


def mdcn_forward_literal(x, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
    """Literal fp32 restatement of the reference modulated deformable conv.

    Follows thirdparty/deform_conv/src/deform_conv_cuda_kernel.cu:402-432
    (dmcn_im2col_bilinear), :505-571 (modulated_deformable_im2col_gpu_kernel) and
    src/deform_conv_cuda.cpp:531-548 (weight @ columns + bias), groups == 1.
    Vectorised over pixels with numpy fp32 so every per-sample arithmetic step
    keeps the kernel's operation order:  val = w1*v1 + w2*v2 + w3*v3 + w4*v4,
    col = val * mask.
    """
    x = x.detach().cpu().numpy().astype(np.float32)
    offset = offset.detach().cpu().numpy().astype(np.float32)
    mask = mask.detach().cpu().numpy().astype(np.float32)
    w = weight.detach().cpu().numpy().astype(np.float32)
    B, C, H, W = x.shape
    Co, Ci, kh, kw = w.shape
    assert Ci == C
    Ho = (H + 2 * padding - (dilation * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * padding - (dilation * (kw - 1) + 1)) // stride + 1
    cpg = C // deformable_groups
    hcol, wcol = np.meshgrid(np.arange(Ho), np.arange(Wo), indexing="ij")
    h_in = (hcol * stride - padding).astype(np.float32)
    w_in = (wcol * stride - padding).astype(np.float32)
    cols = np.zeros((B, C * kh * kw, Ho * Wo), dtype=np.float32)
    one = np.float32(1)
    for b in range(B):
        for c in range(C):
            g = c // cpg
            img = x[b, c]
            for i in range(kh):
                for j in range(kw):
                    tap = i * kw + j
                    off_h = offset[b, g * 2 * kh * kw + 2 * tap]
                    off_w = offset[b, g * 2 * kh * kw + 2 * tap + 1]
                    m = mask[b, g * kh * kw + tap]
                    h_im = h_in + np.float32(i * dilation) + off_h
                    w_im = w_in + np.float32(j * dilation) + off_w
                    valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
                    h_low = np.floor(h_im)
                    w_low = np.floor(w_im)
                    lh = h_im - h_low
                    lw = w_im - w_low
                    hh = one - lh
                    hw = one - lw
                    hl = h_low.astype(np.int64)
                    wl = w_low.astype(np.int64)
                    hh_i = hl + 1
                    wh_i = wl + 1

                    def tap_val(hi, wi, ok):
                        ok = ok & valid
                        v = img[np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1)]
                        return np.where(ok, v, np.float32(0))

                    v1 = tap_val(hl, wl, (hl >= 0) & (wl >= 0))
                    v2 = tap_val(hl, wh_i, (hl >= 0) & (wh_i <= W - 1))
                    v3 = tap_val(hh_i, wl, (hh_i <= H - 1) & (wl >= 0))
                    v4 = tap_val(hh_i, wh_i, (hh_i <= H - 1) & (wh_i <= W - 1))
                    w1, w2, w3, w4 = hh * hw, hh * lw, lh * hw, lh * lw
                    val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4
                    val = np.where(valid, val, np.float32(0))
                    cols[b, c * kh * kw + tap] = (val * m).reshape(-1)
    out = np.einsum("ok,bkp->bop", w.reshape(Co, -1), cols).reshape(B, Co, Ho, Wo)
    if bias is not None:
        out = out + bias.detach().cpu().numpy().astype(np.float32).reshape(1, Co, 1, 1)
    return torch.from_numpy(out.astype(np.float32))


def mdcn_forward(x, offset, mask, weight, bias, stride, padding, dilation):
    """Executable DCN oracle used inside ``head_forward``: torchvision's
    deform_conv2d (same offset/mask channel layout as the reference op;
    checked equal to ``mdcn_forward_literal`` in tests/test_oracle.py)."""
    from torchvision.ops import deform_conv2d
    return deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding,
                         dilation=dilation, mask=mask)


# --------------------------------------------------------------------------
# model/OTPose.py  (forward after the backbone)
# --------------------------------------------------------------------------


def fusion_prologue(rough_heatmaps, margin):
    """model/OTPose.py:320-330, 339-354 -- everything that does not need weights.

    ``rough_heatmaps`` is (5B, J, H, W) ordered cur, prev, next, pprev, nnext
    (OTPose.py:320-321); ``margin`` is (B, 4) integer (left, right, lleft, rright).
    """
    B = rough_heatmaps.shape[0] // 5
    cur, prev, nxt, pprev, nnext = rough_heatmaps.split(B, dim=0)
    total_b = cur + prev + nxt + pprev + nnext
    squeezed = torch.sum(total_b, axis=1)
    J = cur.shape[1]
    squeezed = torch.stack([squeezed for _ in range(J)], dim=1)
    intersection = total_b * squeezed
    prev = torch.div(prev, (margin.T[0] + 1)[:, None, None, None])
    nxt = torch.div(nxt, (margin.T[1] + 1)[:, None, None, None])
    pprev = torch.div(pprev, (margin.T[2] + 1)[:, None, None, None])
    nnext = torch.div(nnext, (margin.T[3] + 1)[:, None, None, None])
    prev_b = cur + (prev + pprev)
    next_b = cur + (nxt + nnext)
    close_b = cur + (nxt + prev)
    far_b = cur + (nnext + pprev)
    return dict(total_b=total_b, squeezed=squeezed, intersection=intersection,
                prev_b=prev_b, next_b=next_b, close_b=close_b, far_b=far_b,
                prev_int=prev_b * squeezed, next_int=next_b * squeezed,
                close_int=close_b * squeezed, far_int=far_b * squeezed)


def fusion_prologue_frames(rough_heatmaps, margin):
    """Window extension of ``fusion_prologue`` for BASELINE config 5 (T = 3 / 5 / 7 frames).

    The reference hard-codes 5 frames (``supplement = 5``, model/OTPose.py:188, 317-321), so
    this is a DEFINITION, not a restatement: frames ordered cur, prev1, next1, prev2, next2,
    ...; margin (B, frames-1) in the same order; sums taken left to right so that the 5-frame
    instance is the reference's expression bit for bit (tests/test_oracle.py asserts equality
    with ``fusion_prologue``).  prev_b / next_b collect every past / future frame, close_b the
    nearest pair, far_b the remaining pairs (= cur alone for 3 frames).
    """
    frames = margin.shape[1] + 1
    assert frames % 2 == 1 and frames >= 3
    B = rough_heatmaps.shape[0] // frames
    fr = rough_heatmaps.split(B, dim=0)
    cur, J = fr[0], fr[0].shape[1]
    total_b = cur
    for f in fr[1:]:
        total_b = total_b + f
    squeezed = torch.sum(total_b, axis=1)
    squeezed = torch.stack([squeezed for _ in range(J)], dim=1)
    sc = [torch.div(fr[i + 1], (margin.T[i] + 1)[:, None, None, None]) for i in range(frames - 1)]
    prevs, nexts = sc[0::2], sc[1::2]
    ps, ns = prevs[0], nexts[0]
    for k in range(1, len(prevs)):
        ps, ns = ps + prevs[k], ns + nexts[k]
    prev_b, next_b = cur + ps, cur + ns
    close_b = cur + (nexts[0] + prevs[0])
    if len(prevs) >= 2:
        fs = nexts[1] + prevs[1]
        for k in range(2, len(prevs)):
            fs = fs + (nexts[k] + prevs[k])
        far_b = cur + fs
    else:
        far_b = cur.clone()
    return dict(total_b=total_b, squeezed=squeezed, intersection=total_b * squeezed,
                prev_b=prev_b, next_b=next_b, close_b=close_b, far_b=far_b,
                prev_int=prev_b * squeezed, next_int=next_b * squeezed,
                close_int=close_b * squeezed, far_int=far_b * squeezed)


def head_forward(sd, rough_heatmaps, margin, dilations=(3, 6, 9, 12, 15), num_rsb_blocks=2,
                 return_intermediates=False):
    """OTPose.forward lines 320-394 (model/OTPose.py), given the backbone output.

    Returns the reference 7-tuple (output_heatmaps, rough_heatmaps, intersection,
    prev_b, context_encoding, squeezed, total_b).
    """
    frames = margin.shape[1] + 1           # 5 in the reference; 3 / 7 = the config-5 window extension
    B = rough_heatmaps.shape[0] // frames
    J, H, W = rough_heatmaps.shape[1:]
    f = fusion_prologue(rough_heatmaps, margin) if frames == 5 else fusion_prologue_frames(rough_heatmaps, margin)
    ctx = conv_transformer(sd, "flow_encoder.", f["total_b"], 1, (0, 6, 0))
    context_encoding = torch.stack([s for s in ctx], dim=1).contiguous().view(B, J, H, W)
    x1 = torch.stack((f["intersection"], context_encoding, f["prev_b"], f["far_b"], f["close_b"],
                      f["prev_int"], f["far_int"], f["close_int"]), dim=2).flatten(start_dim=1, end_dim=2)
    x2 = torch.stack((f["intersection"], context_encoding, f["next_b"], f["close_b"], f["far_b"],
                      f["next_int"], f["close_int"], f["far_int"]), dim=2).flatten(start_dim=1, end_dim=2)
    e1 = conv_transformer(sd, "temporal_encoder1.", x1, 2, (0, 6, 2))
    e2 = conv_transformer(sd, "temporal_encoder2.", x2, 2, (0, 6, 2))
    C = x1.shape[1]
    y1 = torch.stack([s for s in e1], dim=1).contiguous().view(B, C * 3, H, W)
    y2 = torch.stack([s for s in e2], dim=1).contiguous().view(B, C * 3, H, W)
    y1 = F.conv2d(y1, sd["final_layer1.weight"], sd["final_layer1.bias"])
    y2 = F.conv2d(y2, sd["final_layer2.weight"], sd["final_layer2.bias"])
    branches = torch.cat([y1, y2], dim=1)
    def_heatmaps = chain_rsb(sd, "def_fuse.", f["total_b"], num_rsb_blocks)
    trans = chain_rsb(sd, "offset_mask_combine_conv.", torch.cat([branches, def_heatmaps], dim=1),
                      num_rsb_blocks)
    warped = []
    inter = dict(x1=x1, x2=x2, e1=e1, e2=e2, branches=branches, def_heatmaps=def_heatmaps, trans=trans,
                 offsets=[], masks=[])
    for i, d in enumerate(dilations):
        offsets = F.conv2d(trans, sd[f"offsets_list.{i}.0.weight"], None, padding=d, dilation=d)
        masks = F.conv2d(trans, sd[f"masks_list.{i}.0.weight"], None, padding=d, dilation=d)
        warped.append(mdcn_forward(def_heatmaps, offsets, masks,
                                   sd[f"modulated_deform_conv_list.{i}.deform_conv.weight"],
                                   sd[f"modulated_deform_conv_list.{i}.deform_conv.bias"], 1, d, d))
        if return_intermediates:
            inter["offsets"].append(offsets)
            inter["masks"].append(masks)
    ww = 1 / len(dilations)
    out = ww * warped[0]
    for w_ in warped[1:]:
        out = out + ww * w_
    res = (out, rough_heatmaps, f["intersection"], f["prev_b"], context_encoding, f["squeezed"], f["total_b"])
    return (res, inter) if return_intermediates else res


# --------------------------------------------------------------------------
# utils/heatmap.py + utils/transform.py
# --------------------------------------------------------------------------


def get_max_preds(batch_heatmaps):
    """utils/heatmap.py:143-171 (numpy; argmax returns the first maximum)."""
    assert batch_heatmaps.ndim == 4
    n, j, _, width = batch_heatmaps.shape
    flat = batch_heatmaps.reshape((n, j, -1))
    idx = np.argmax(flat, 2)
    maxvals = np.amax(flat, 2).reshape((n, j, 1))
    idx = idx.reshape((n, j, 1))
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % width
    preds[:, :, 1] = np.floor(preds[:, :, 1] / width)
    pred_mask = np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    preds *= pred_mask
    return preds, maxvals


def _third_point(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def affine_transform_inv(center, scale, output_size):
    """get_affine_transform(center, scale, 0, output_size, inv=1), utils/transform.py:76-105.

    rot == 0 and shift == 0 as ``transform_preds`` calls it (utils/heatmap.py:137).
    cv2.getAffineTransform (opencv-python, third-party) is the exact solve of the
    three-point correspondence in float64; restated with numpy.linalg.solve.
    """
    scale = np.asarray(scale, dtype=np.float64)
    center = np.asarray(center)
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + np.array([0, src_w * -0.5])
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    # inv=1: map dst -> src.   [x y 1] @ M^T = src
    a = np.concatenate([dst.astype(np.float64), np.ones((3, 1))], axis=1)
    m = np.linalg.solve(a, src.astype(np.float64))
    return m.T  # (2, 3)


def transform_preds(coords, center, scale, output_size):
    """utils/heatmap.py:135-140."""
    target = np.zeros(coords.shape)
    trans = affine_transform_inv(center, scale, output_size)
    for p in range(coords.shape[0]):
        pt = np.array([coords[p, 0], coords[p, 1], 1.0]).T
        target[p, 0:2] = np.dot(trans, pt)[:2]
    return target


def get_final_preds(batch_heatmaps, center, scale):
    """utils/heatmap.py:108-132: argmax + quarter-pixel shift + back-projection."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    hh, ww = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    for n in range(coords.shape[0]):
        for p in range(coords.shape[1]):
            hm = batch_heatmaps[n][p]
            px = int(math.floor(coords[n][p][0] + 0.5))
            py = int(math.floor(coords[n][p][1] + 0.5))
            if 1 < px < ww - 1 and 1 < py < hh - 1:
                diff = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]])
                coords[n][p] += np.sign(diff) * .25
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [ww, hh])
    return preds, maxvals


def final_preds_full(batch_heatmaps, center, scale):
    """Same as get_final_preds but also returns the flat argmax index and the
    heatmap-space coordinates -- the integer / exactly-representable parts the
    CUDA kernel must match bit-exactly."""
    n, j, hh, ww = batch_heatmaps.shape
    idx = np.argmax(batch_heatmaps.reshape(n, j, -1), 2).astype(np.int32)
    coords, maxvals = get_max_preds(batch_heatmaps)
    for a in range(n):
        for p in range(j):
            hm = batch_heatmaps[a][p]
            px = int(math.floor(coords[a][p][0] + 0.5))
            py = int(math.floor(coords[a][p][1] + 0.5))
            if 1 < px < ww - 1 and 1 < py < hh - 1:
                diff = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]])
                coords[a][p] += np.sign(diff) * .25
    preds = coords.copy()
    for i in range(n):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [ww, hh])
    return idx, coords, preds, maxvals
