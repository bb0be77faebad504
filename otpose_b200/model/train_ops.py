"""Training-mode forward of the head's modules, differentiable.

STATUS -- read this before citing it.  The inference path of this package is hand-written CUDA
behind the C ABI.  The TRAINING path is a first version with a different split:

* NATIVE forward + backward (C ABI kernels): the modulated deformable convolution
  (``otp_mdcn_forward`` / ``otp_mdcn_backward``), the dilated offset / mask convolutions that feed it, every
  convolution of the RSB chains and the two pyramid 1x1 convs (``otp_conv2d`` / ``otp_conv2d_wgrad`` behind
  ``Conv2dFunction``).
* LIBRARY ops under autograd (ATen: cuBLAS / cuDNN / elementwise): the TransformerBlocks (pointwise /
  depthwise convs, Gram attention, LayerNorm, GELU, dropout), BatchNorm + ReLU of the RSB chains and the
  fusion prologue, composed below from the same arithmetic the reference's modules execute (model/blocks.py:95-110, 264-279, 289-316, 400-452; model/RSB.py:81-139;
  model/OTPose.py:320-375).  They are explicit library calls, not a fallback of the CUDA path: eval-mode
  calls never reach this file and a training call never reaches the fused inference kernels.

Randomness: dropout (p = 0.1 after GELU, after W2 and after the attention projection) and per-sample
drop-path use torch's Philox generator on the device (``torch.cuda.manual_seed`` replays them).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def layer_norm_ct(x, ln):
    """Channel LayerNorm over (B, C, T): biased variance, eps inside the sqrt (model/blocks.py:95-110)."""
    mu = x.mean(dim=1, keepdim=True)
    d = x - mu
    var = (d * d).mean(dim=1, keepdim=True)
    y = d * torch.rsqrt(var + ln.eps)
    return y * ln.weight + ln.bias if ln.affine else y


def drop_path(x, p, training):
    """Per-sample stochastic depth: x / keep * floor(keep + U) (model/blocks.py:301-316)."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    mask = torch.floor(keep + torch.rand((x.shape[0],) + (1,) * (x.dim() - 1), dtype=x.dtype, device=x.device))
    return x / keep * mask


def attention(attn, h):
    """MaskedMHCA.forward (model/blocks.py:400-452): depthwise conv -> LayerNorm -> pointwise for q, k, v;
    channel-Gram softmax per head over ALL tokens; the transpose(2,3).contiguous().view scramble; proj."""
    b, c, _ = h.shape
    nh, hs = attn.n_head, attn.n_channels

    def branch(conv, norm, lin):
        y = F.conv1d(h, conv.weight, None, stride=conv.stride, padding=conv.padding, groups=c)
        return F.conv1d(layer_norm_ct(y, norm), lin.weight, lin.bias)

    q = branch(attn.query_conv, attn.query_norm, attn.query).view(b, nh, hs, -1)
    k = branch(attn.key_conv, attn.key_norm, attn.key).view(b, nh, hs, -1)
    v = branch(attn.value_conv, attn.value_norm, attn.value).view(b, nh, hs, -1)
    att = torch.softmax((q * attn.scale) @ k.transpose(-2, -1), dim=-1)
    att = attn.attn_drop(att)
    out = (att @ v).transpose(2, 3).contiguous().view(b, c, -1)
    return attn.proj_drop(F.conv1d(out, attn.proj.weight, attn.proj.bias))


def transformer_block(blk, x):
    """TransformerBlock.forward (model/blocks.py:264-279) with dropout / drop-path as the modules are set."""
    def path(dp, y):
        scale = getattr(dp, "scale", None)
        if scale is None:
            return y
        return drop_path(scale * y, dp.drop_prob, dp.training)

    out = blk.pool_skip(x) + path(blk.drop_path_attn, attention(blk.attn, layer_norm_ct(x, blk.ln1)))
    return out + path(blk.drop_path_mlp, blk.mlp(layer_norm_ct(out, blk.ln2)))


def conv_transformer(enc, x):
    """ConvTransformer.forward in training mode (model/ConvVideoTransformer.py:123-184): the positional
    embedding is sliced (T <= max_len is asserted, :141-146), branch outputs are upsampled linearly."""
    b, c, h, w = x.shape
    t = h * w
    x = x.reshape(b, c, t)
    if enc.use_abs_pe:
        assert t <= enc.max_len, "Reached max length."
        x = x + enc.pos_embd[:, :, :t]
    for blk in enc.stem:
        x = transformer_block(blk, x)
    outs = (x,)
    for i, blk in enumerate(enc.branch):
        x = transformer_block(blk, x)
        outs += (F.interpolate(x, scale_factor=float(2 ** (i + 1)), mode="linear"),)
    return outs


def conv_bn_relu(m, x):
    """conv_bn_relu.forward (model/RSB.py:106-139); BatchNorm uses batch statistics when m.bn.training.
    The convolution runs on the library's own kernels, forward and backward (Conv2dFunction: otp_conv2d /
    otp_conv2d_wgrad); BatchNorm and ReLU are library ops."""
    from .conv2d_fn import conv2d as conv2d_native
    x = conv2d_native(x, m.conv.weight, m.conv.bias, 1)
    if m.has_bn:
        x = m.bn(x)
    return F.relu(x) if m.has_relu else x


def rsb_block(blk, x):
    """RSB_BLOCK.forward (model/RSB.py:81-103): ten dense-connected 3x3 branches between two 1x1 convs."""
    spx = torch.split(conv_bn_relu(blk.conv_bn_relu1, x), blk.branch_ch, 1)

    def c(name, t):
        return conv_bn_relu(getattr(blk, "conv_bn_relu2_" + name), t)

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
    out = conv_bn_relu(blk.conv_bn_relu3, torch.cat((o11, o22, o33, o44), 1))
    skip = x if blk.downsample is None else conv_bn_relu(blk.downsample, x)
    return F.relu(out + skip)


def chain_rsb(chain, x):
    for blk in chain.layers:
        x = rsb_block(blk, x)
    return x


def fusion_prologue(rough, margin):
    """model/OTPose.py:320-330, 339-354 (5 frames: cur, prev, next, pprev, nnext)."""
    b = rough.shape[0] // 5
    cur, prev, nxt, pprev, nnext = rough.split(b, dim=0)
    total_b = cur + prev + nxt + pprev + nnext
    squeezed = total_b.sum(dim=1, keepdim=True).expand_as(total_b)
    div = (margin.to(rough.dtype) + 1.0).t()[:, :, None, None, None]
    prev, nxt, pprev, nnext = prev / div[0], nxt / div[1], pprev / div[2], nnext / div[3]
    prev_b, next_b = cur + (prev + pprev), cur + (nxt + nnext)
    close_b, far_b = cur + (nxt + prev), cur + (nnext + pprev)
    return dict(total_b=total_b, squeezed=squeezed, intersection=total_b * squeezed, prev_b=prev_b, next_b=next_b,
                close_b=close_b, far_b=far_b)


def head_forward_train(model, rough_heatmaps, margin):
    """OTPose.forward lines 320-394 (model/OTPose.py) under autograd.  Returns the reference 7-tuple."""
    from ..thirdparty.deform_conv import modulated_deform_conv
    from .conv2d_fn import conv2d as conv2d_native
    if margin.shape[1] != 4:
        raise NotImplementedError("the training path follows the reference's 5-frame window")
    rough = rough_heatmaps.float()
    b = rough.shape[0] // 5
    j, h, w = rough.shape[1:]
    f = fusion_prologue(rough, margin)
    sq = f["squeezed"]
    ctx = torch.stack(conv_transformer(model.flow_encoder, f["total_b"]), dim=1).contiguous().view(b, j, h, w)

    def stack8(*maps):
        return torch.stack(maps, dim=2).flatten(1, 2)

    x1 = stack8(f["intersection"], ctx, f["prev_b"], f["far_b"], f["close_b"], f["prev_b"] * sq, f["far_b"] * sq,
                f["close_b"] * sq)
    x2 = stack8(f["intersection"], ctx, f["next_b"], f["close_b"], f["far_b"], f["next_b"] * sq, f["close_b"] * sq,
                f["far_b"] * sq)
    c8 = x1.shape[1]
    y1 = torch.stack(conv_transformer(model.temporal_encoder1, x1), dim=1).contiguous().view(b, 3 * c8, h, w)
    y2 = torch.stack(conv_transformer(model.temporal_encoder2, x2), dim=1).contiguous().view(b, 3 * c8, h, w)
    fl1, fl2 = model.final_layer1, model.final_layer2      # 1x1 convs 408 -> 17 on the library's own kernels
    branches = torch.cat([conv2d_native(y1, fl1.weight, fl1.bias, 1), conv2d_native(y2, fl2.weight, fl2.bias, 1)], dim=1)
    def_heatmaps = chain_rsb(model.def_fuse, f["total_b"])
    trans = chain_rsb(model.offset_mask_combine_conv, torch.cat([branches, def_heatmaps], dim=1))
    ww = 1.0 / len(model.deformable_conv_dilations)
    out = None
    for i, dd in enumerate(model.deformable_conv_dilations):
        # native stage: dilated offset / mask convs and the modulated DCN, forward and backward on the C ABI
        offsets = conv2d_native(trans, model.offsets_list[i][0].weight, None, dd)
        masks = conv2d_native(trans, model.masks_list[i][0].weight, None, dd)
        dcn = model.modulated_deform_conv_list[i].deform_conv
        warped = modulated_deform_conv(def_heatmaps.contiguous(), offsets, masks, dcn.weight, dcn.bias, dcn.stride,
                                       dcn.padding, dcn.dilation, dcn.groups, dcn.deformable_groups)
        out = ww * warped if out is None else out + ww * warped
    return out, rough_heatmaps, f["intersection"], f["prev_b"], ctx, sq, f["total_b"]
