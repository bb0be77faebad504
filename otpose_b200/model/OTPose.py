"""Drop-in ``OTPose`` for the temporal fusion head (reference model/OTPose.py:180-257
constructor / state-dict layout, :307-394 forward).

Everything after the backbone runs in the CUDA kernels behind the C ABI:

    rough_heatmaps (5B,J,H,W), margin (B,4)
      -> otp_fusion_sum                       total_b, squeezed           (OTPose.py:324-326)
      -> flow_encoder (6 fused blocks, C=J)   context_encoding             (:331-335)
      -> otp_fusion_stack                     x1, x2 (+pos_embd), intersection, prev_b (:330-359)
      -> temporal_encoder1/2 (6+2 blocks)     3-scale pyramids             (:360-361)
      -> otp_pyramid_conv1x1 x2               branches (upsample + stack + final_layer fused) (:362-375)
      -> def_fuse, offset_mask_combine_conv   RSB chains via otp_conv2d    (:376-378)
      -> per dilation: offset conv, mask conv, otp_mdcn_forward accumulating 0.2*dcn (:381-392)

The HRNet backbone is out of scope (cuDNN convolutions): pass any
module producing (5B, J, H, W) heat maps as ``backbone`` to use ``forward``;
``forward_head`` is the hot path itself.
"""
from __future__ import annotations

import logging

import torch
from torch import nn

from .. import _lib
from .ConvVideoTransformer import ConvTransformer
from .RSB import CHAIN_RSB_BLOCKS, conv_bn_relu
from .layers import DeformableCONV
from ..thirdparty.deform_conv import ModulatedDeformConv

logger = logging.getLogger(__name__)


class AttrDict(dict):
    """Minimal yacs-CfgNode stand-in: attribute and item access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def default_cfg(heatmap_hw=(96, 72), num_joints=17, dilations=(3, 6, 9, 12, 15)):
    """The head-relevant keys of configs/17/model_RSN.yaml + configs/default.py."""
    h, w = heatmap_hw
    return AttrDict(MODEL=AttrDict(
        EXTRA=AttrDict(FINAL_CONV_KERNEL=1, PRETRAINED_LAYERS=["*"]),
        HEATMAP_SIZE=[w, h], NUM_JOINTS=num_joints, FREEZE_HRNET_WEIGHTS=True, PRETRAINED="",
        DEFORMABLE_CONV=AttrDict(DILATION=list(dilations), AGGREGATION_TYPE="weighted_sum"),
        DEFORMABLE_CONV_CH=32, OFFSET_MASK_COMBINE_CONV=2))


def _mask_conv(nc, kh, kw, dd, dg):
    return nn.Conv2d(nc, dg * 1 * kh * kw, kernel_size=(3, 3), stride=(1, 1), dilation=(dd, dd),
                     padding=(1 * dd, 1 * dd), bias=False)


def _offset_conv(nc, kh, kw, dd, dg):
    return nn.Conv2d(nc, dg * 2 * kh * kw, kernel_size=(3, 3), stride=(1, 1), dilation=(dd, dd),
                     padding=(1 * dd, 1 * dd), bias=False)


class OTPose(nn.Module):
    def __init__(self, cfg, backbone=None, precision="fp32", cuda_graph=False, **kwargs):
        super().__init__()
        extra = cfg['MODEL']['EXTRA']
        self.num_frames = 8
        self.pe_w, self.pe_h = cfg.MODEL.HEATMAP_SIZE
        self.num_joints = cfg.MODEL.NUM_JOINTS
        self.patch_size = 1
        self.num_patches = self.pe_h * self.pe_w
        self.patch_dim = self.num_joints
        self.temporal_encoding_dim = self.patch_dim * self.num_frames
        self.precision = precision
        self.cuda_graph = bool(cuda_graph)   # opt-in: replay forward_head from a captured CUDA graph
        self.graph_clone_outputs = True      # graph mode returns copies; False = the graph's static buffers, which
                                             # the NEXT forward_head call overwrites (zero-copy serving loops)
        self._graphs = {}
        self._side_streams = {}
        self._pyramid_cache = {}
        self.overlap_branches = True    # def_fuse on a side stream, concurrent with the flow encoder
        self.overlap_encoders = True    # temporal_encoder2 on the side stream, concurrent with temporal_encoder1
        if extra['FINAL_CONV_KERNEL'] != 1:
            raise NotImplementedError("final_layer kernels are built for FINAL_CONV_KERNEL = 1")
        if backbone is not None:
            self.rough_pose_estimation_net = backbone

        self.scale_arch = (0, 6, 2)
        self.flow_scale_arch = (0, 6, 0)
        self.max_seq_len = self.num_patches
        d = self.temporal_encoding_dim
        self.temporal_encoder1 = ConvTransformer(d, d, n_head=2, n_embd_ks=3, max_len=self.num_patches,
                                                 arch=self.scale_arch, proj_pdrop=0.1, path_pdrop=0.1,
                                                 h=self.pe_h, precision=precision)
        self.temporal_encoder2 = ConvTransformer(d, d, n_head=2, n_embd_ks=3, max_len=self.num_patches,
                                                 arch=self.scale_arch, proj_pdrop=0.1, path_pdrop=0.1,
                                                 h=self.pe_h, precision=precision)
        self.flow_encoder = ConvTransformer(self.patch_dim, self.patch_dim, 1, 3, self.num_patches,
                                            arch=self.flow_scale_arch, proj_pdrop=0.1, path_pdrop=0.1,
                                            h=self.pe_h, precision=precision)
        self.deformable_conv_dilations = list(cfg.MODEL.DEFORMABLE_CONV.DILATION)
        self.deformable_aggregation_type = cfg.MODEL.DEFORMABLE_CONV.AGGREGATION_TYPE
        if self.deformable_aggregation_type != "weighted_sum":
            raise NotImplementedError("only AGGREGATION_TYPE = weighted_sum defines an output (OTPose.py:387)")
        self.final_layer1 = nn.Conv2d(d * 3, self.num_joints, kernel_size=1, stride=1, padding=0)
        self.final_layer2 = nn.Conv2d(d * 3, self.num_joints, kernel_size=1, stride=1, padding=0)
        self.pretrained_layers = extra['PRETRAINED_LAYERS']

        k = 3
        def_ch = cfg.MODEL.DEFORMABLE_CONV_CH
        n_rsb = cfg.MODEL.OFFSET_MASK_COMBINE_CONV
        self.offset_mask_combine_conv = CHAIN_RSB_BLOCKS(self.num_joints * 3, def_ch, n_rsb)
        self.def_fuse = CHAIN_RSB_BLOCKS(self.num_joints, self.num_joints, n_rsb)
        self.offsets_list = nn.ModuleList(
            nn.Sequential(_offset_conv(def_ch, k, k, dd, self.num_joints)) for dd in self.deformable_conv_dilations)
        self.masks_list = nn.ModuleList(
            nn.Sequential(_mask_conv(def_ch, k, k, dd, self.num_joints)) for dd in self.deformable_conv_dilations)
        self.modulated_deform_conv_list = nn.ModuleList(
            DeformableCONV(self.num_joints, k, dd) for dd in self.deformable_conv_dilations)
        self.init_weights()

    def init_weights(self):
        """Head part of the reference init (model/OTPose.py:431-468)."""
        for name, m in self.named_modules():
            if name.split('.')[0] == "rough_pose_estimation_net":
                continue
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, ModulatedDeformConv):
                with torch.no_grad():
                    m.weight.zero_()
                    for k in range(m.weight.size(0)):
                        m.weight[k, k, m.weight.size(2) // 2, m.weight.size(3) // 2] = 1.0
                    if m.bias is not None:
                        m.bias.zero_()

    def _offset_mask_packed(self, i):
        """16-bit operand images of offsets_list[i] / masks_list[i] for the fused tensor-core
        kernel; rebuilt when either weight changed."""
        wo, wm = self.offsets_list[i][0].weight, self.masks_list[i][0].weight
        key = (id(wo), wo.data_ptr(), wo._version, id(wm), wm.data_ptr(), wm._version)
        cache = self.__dict__.setdefault("_om_cache", {})
        if cache.get(i, (None,))[0] != key:
            lib = _lib.load()
            buf = torch.empty(lib.otp_offset_mask_pack_bytes(), dtype=torch.uint8, device=wo.device)
            with torch.cuda.device(wo.device):
                _lib.check(lib.otp_offset_mask_pack(_lib.dptr(wo.detach()), _lib.dptr(wm.detach()), self.num_joints,
                                                    wo.shape[1], buf.data_ptr(), buf.numel(),
                                                    _lib.stream_ptr(wo.device)), "otp_offset_mask_pack")
            cache[i] = (key, buf, (wo, wm))          # the sources stay alive with the entry
        return cache[i][1]

    # ------------------------------------------------------------------
    def forward(self, x, **kwargs):
        """Reference signature: x (B, 15, Himg, Wimg), margin=(B, 4)."""
        assert "margin" in kwargs
        if not hasattr(self, "rough_pose_estimation_net"):
            raise RuntimeError("OTPose was built without a backbone; call forward_head(rough_heatmaps, margin)")
        x = torch.cat(x.split(3, dim=1), 0)
        rough_heatmaps = self.rough_pose_estimation_net(x)
        return self.forward_head(rough_heatmaps, kwargs["margin"])

    def _pyramid_packed(self, i, fl, prec):
        """16-bit operand image of final_layer{i+1}.weight (cached; rebuilt when the parameter changes)."""
        key = (id(fl.weight), fl.weight.data_ptr(), fl.weight._version, prec)
        ent = self._pyramid_cache.get(i)
        if ent is None or ent[0] != key:
            lib = _lib.load()
            j, c3 = fl.weight.shape[0], fl.weight.shape[1]
            nbytes = lib.otp_pyramid_conv1x1_tc_pack_bytes(c3 // 3)
            buf = torch.empty(nbytes, dtype=torch.uint8, device=fl.weight.device)
            _lib.check(lib.otp_pyramid_conv1x1_tc_pack(_lib.dptr(fl.weight.detach().view(j, c3).contiguous()), c3 // 3, j,
                                                       prec, buf.data_ptr(), nbytes,
                                                       _lib.stream_ptr(fl.weight.device)), "otp_pyramid_conv1x1_tc_pack")
            ent = self._pyramid_cache[i] = (key, buf, fl.weight)
        return ent[1]

    # ------------------------------------------------------------------ a0: HRNet.final_layer boundary
    @torch.no_grad()
    def forward_from_features(self, features, margin, weight=None, bias=None):
        """Wider boundary (the backbone -> head hand-off): take the backbone's last
        feature map ``y_list[0]`` (frames*B, Cin, H, W) and apply ``HRNet.final_layer`` (1x1 conv Cin -> J,
        model/HRNet.py:108-114, 150) here.  ``features`` may be fp32 / bf16 / fp16, NCHW-contiguous (what
        the reference's backbone returns) or ``torch.channels_last`` (a channels-last cuDNN backbone): one
        kernel (``otp_final_layer_fusion_sum``) reads it once and writes the rough heat maps together with
        ``total_b`` / ``squeezed`` of model/OTPose.py:324-326, so the head starts without re-reading them.
        ``weight`` / ``bias`` default to ``self.rough_pose_estimation_net.final_layer`` when a backbone with
        that attribute is attached.  Returns the reference 7-tuple; element 1 is the rough heat maps
        computed here (fp32)."""
        _lib.require_cuda(features)
        if weight is None:
            fl = self.rough_pose_estimation_net.final_layer
            weight, bias = fl.weight, fl.bias
        n, cin, h, w = features.shape
        j = weight.shape[0]
        if tuple(weight.shape[1:]) != (cin, 1, 1):
            raise NotImplementedError("final_layer: built for FINAL_CONV_KERNEL = 1")
        frames = margin.shape[1] + 1
        if frames not in (3, 5, 7) or n % frames:
            raise NotImplementedError(f"frame window of {frames} not built (3, 5 or 7)")
        dcode = {torch.float32: _lib.PREC_FP32, torch.bfloat16: _lib.PREC_BF16, torch.float16: _lib.PREC_FP16}
        feats = features
        if feats.dtype not in dcode:
            feats = feats.float()
        if feats.is_contiguous():
            nhwc = 0
        elif feats.is_contiguous(memory_format=torch.channels_last):
            nhwc = 1
        else:
            feats, nhwc = feats.contiguous(), 0
        b, t, dev = n // frames, h * w, feats.device
        f32 = dict(dtype=torch.float32, device=dev)
        rough = torch.empty((n, j, h, w), **f32)
        total_b = torch.empty((b, j, h, w), **f32)
        squeezed = torch.empty((b, 1, h, w), **f32)
        fused = j == 17 and cin % 8 == 0 and cin <= 64
        if n and not fused:       # other widths: the generic fp32 conv kernel, then the head from the rough maps
            lib = _lib.load()
            feats = feats.float().contiguous()
            wt = weight.detach().float().contiguous()
            bs = (bias.detach().float().contiguous() if bias is not None
                  else torch.zeros(j, dtype=torch.float32, device=dev))
            with torch.cuda.device(dev):
                _lib.check(lib.otp_conv2d(_lib.dptr(feats), cin * t, None, 0, _lib.dptr(wt), _lib.dptr(bs), None, 0,
                                          rough.data_ptr(), j * t, n, cin, h, w, j, 1, 1, 0, _lib.stream_ptr(dev)),
                           "otp_conv2d")
            return self.forward_head(rough, margin)
        if n:
            lib = _lib.load()
            wt = weight.detach().float().reshape(j, cin).contiguous()
            bs = bias.detach().float().contiguous() if bias is not None else None
            with torch.cuda.device(dev):
                _lib.check(lib.otp_final_layer_fusion_sum(
                    feats.data_ptr(), dcode[feats.dtype], nhwc, _lib.dptr(wt), _lib.dptr(bs, allow_none=True), frames,
                    b, cin, j, t, rough.data_ptr(), total_b.data_ptr(), squeezed.data_ptr(), _lib.stream_ptr(dev)),
                    "otp_final_layer_fusion_sum")
        if self.cuda_graph and not self.training and n:
            return self.forward_head(rough, margin)      # the captured graph starts from the rough heat maps
        return self._forward_head_eager(rough, margin, _pre=(total_b, squeezed))

    # ------------------------------------------------------------------ CUDA graph replay (opt-in)
    def invalidate_graphs(self):
        """Drop captured graphs (call after changing parameters in place; ``load_state_dict``, ``.to()``
        and ``.train()`` do it themselves)."""
        self._graphs = {}

    def load_state_dict(self, *a, **k):
        self.invalidate_graphs()
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._graphs = {}
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        self._graphs = {}
        return super().train(mode)

    @torch.no_grad()
    def _forward_head_graphed(self, rough_heatmaps, margin):
        """The ~190 launches of one head forward captured once per input shape in a CUDA graph and
        replayed: the inputs are copied into the graph's static buffers (75 MB device copy at 32 clips,
        ~25 us), the returned tensors are the graph's static outputs -- valid until the next call."""
        rough = rough_heatmaps if rough_heatmaps.dtype == torch.float32 else rough_heatmaps.float()
        dev = rough.device
        key = (dev.index, tuple(rough.shape), tuple(margin.shape), self.precision)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            s_rough = torch.empty_like(rough, memory_format=torch.contiguous_format)
            s_margin = torch.empty(margin.shape, dtype=torch.int64, device=dev)
            s_rough.copy_(rough)
            s_margin.copy_(margin)
            warm = torch.cuda.Stream(device=dev)          # warm-up off the capture: weight packing,
            warm.wait_stream(torch.cuda.current_stream(dev))   # workspace growth, function attributes
            with torch.cuda.stream(warm):
                self._forward_head_eager(s_rough, s_margin)
                self._forward_head_eager(s_rough, s_margin)
            torch.cuda.current_stream(dev).wait_stream(warm)
            warm.synchronize()
            _lib.workspace.drop_stream(dev, warm)         # the warm-up stream's scratch buffers are not kept
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._forward_head_eager(s_rough, s_margin)
            ent = self._graphs[key] = (graph, s_rough, s_margin, outs)
        graph, s_rough, s_margin, outs = ent
        s_rough.copy_(rough, non_blocking=True)
        s_margin.copy_(margin, non_blocking=True)
        graph.replay()
        if self.graph_clone_outputs:
            # callers may keep results across steps (a validation loop accumulating predictions): hand out copies
            # (`squeezed`, a stride-0 expand of one plane, is materialised as the reference's J-fold stack)
            return (outs[0].clone(), rough_heatmaps) + tuple(o.clone() for o in outs[2:])
        return (outs[0], rough_heatmaps) + tuple(outs[2:])

    def forward_head(self, rough_heatmaps, margin, _debug=None):
        """model/OTPose.py:320-394.  rough_heatmaps (5B, J, H, W) fp32 CUDA ordered
        cur, prev, next, pprev, nnext; margin (B, 4) integer.  Returns the reference
        7-tuple.  ``squeezed`` is returned as a stride-0 expand of the (B,1,H,W) plane
        (same values as the reference's J-fold stack).  With ``cuda_graph=True`` the launches are
        replayed from a captured CUDA graph (see ``_forward_head_graphed``)."""
        if self.training:
            # differentiable path (model/train_ops.py): native DCN + offset / mask conv forward and backward,
            # library ops under autograd for the blocks / RSB chains -- see the status note there
            from . import train_ops
            _lib.require_cuda(rough_heatmaps)
            return train_ops.head_forward_train(self, rough_heatmaps, margin)
        if (self.cuda_graph and _debug is None and not self.training and rough_heatmaps.is_cuda
                and rough_heatmaps.shape[0] > 0 and not torch.cuda.is_current_stream_capturing()):
            return self._forward_head_graphed(rough_heatmaps, margin)
        return self._forward_head_eager(rough_heatmaps, margin, _debug)

    @torch.no_grad()
    def _forward_head_eager(self, rough_heatmaps, margin, _debug=None, _pre=None):
        if self.training:
            raise NotImplementedError("CUDA OTPose head implements eval-mode forward; call .eval()")
        _lib.require_cuda(rough_heatmaps)
        lib = _lib.load()
        rough = rough_heatmaps
        if rough.dtype != torch.float32 or not rough.is_contiguous():
            rough = rough.float().contiguous()
        nf, j, h, w = rough.shape
        # frame window: 5 in the reference (`supplement = 5`, OTPose.py:188, 317-321); margin carries
        # one column per supplementary frame, so (B, 2) / (B, 6) select the 3- / 7-frame extension
        frames = margin.shape[1] + 1
        if frames not in (3, 5, 7):
            raise NotImplementedError(f"frame window of {frames} not built (3, 5 or 7)")
        assert nf % frames == 0 and nf // frames == margin.shape[0], "rough_heatmaps / margin disagree on B"
        assert j == self.num_joints and (h, w) == (self.pe_h, self.pe_w)
        b, t, dev = nf // frames, h * w, rough.device
        if t % 4 != 0:
            raise ValueError("H*W must be divisible by 4 (two stride-2 branch levels are upsampled back)")
        margin = margin.to(device=dev, dtype=torch.int64).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        c8 = self.temporal_encoding_dim
        if _pre is None:
            total_b = torch.empty((b, j, h, w), **f32)
            squeezed = torch.empty((b, 1, h, w), **f32)
        else:                     # forward_from_features computed them together with the rough heat maps
            total_b, squeezed = _pre
        intersection = torch.empty((b, j, h, w), **f32)
        prev_b = torch.empty((b, j, h, w), **f32)
        out = torch.empty((b, j, h, w), **f32)
        if b == 0:
            ctx = torch.empty((b, j, h, w), **f32)
            return out, rough_heatmaps, intersection, prev_b, ctx, squeezed.expand(b, j, h, w), total_b
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            if _pre is None:
                _lib.check(lib.otp_fusion_sum_frames(_lib.dptr(rough), frames, b, j, t, total_b.data_ptr(),
                                                     squeezed.data_ptr(), st), "otp_fusion_sum_frames")
            # Operand format of the small convolutions behind the encoders (pyramid 1x1, RSB chains, offset /
            # mask convs): IEEE half in BOTH 16-bit modes.  Their inputs are bounded maps (1x1-projected
            # LayerNorm-ed features, BatchNorm + ReLU outputs; the conversions saturate), their weights are
            # shared by every pixel, and a bfloat16-rounded weight is a token-coherent perturbation that the
            # learned-offset sampling amplifies (scripts/emulate_operand_rounding.py) -- "bf16" keeps bfloat16
            # where the range matters, in the encoder blocks (with hi + lo MLP weights).
            refine = "fp32" if self.precision == "fp32" else "fp16"
            if getattr(self, "_rsb_precision", None) != refine:
                for mod in self.modules():
                    if isinstance(mod, conv_bn_relu):
                        mod.precision = refine
                self._rsb_precision = refine
            # def_fuse only needs total_b: its 25 small convs run on a side stream, concurrently with
            # the (equally launch/latency-bound) C=17 flow encoder and the fusion prologue
            cat = torch.empty((b, 3 * j, h, w), **f32)       # [final_layer1 | final_layer2 | def_heatmaps]
            main = torch.cuda.current_stream(dev)
            side = self._side_streams.get(dev.index)
            if side is None:
                side = self._side_streams[dev.index] = torch.cuda.Stream(device=dev)
            if self.overlap_branches:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    def_heatmaps = self.def_fuse(total_b)
                    cat[:, 2 * j:].copy_(def_heatmaps)
                def_heatmaps.record_stream(main)
            else:
                def_heatmaps = self.def_fuse(total_b)
                cat[:, 2 * j:].copy_(def_heatmaps)
            ctx = self.flow_encoder(total_b)[0]                                  # (B, J, T)
            x1 = torch.empty((b, c8, t), **f32)
            x2 = torch.empty((b, c8, t), **f32)
            pe1, ps1 = self.temporal_encoder1.pos_embd_for(t)
            pe2, ps2 = self.temporal_encoder2.pos_embd_for(t)
            assert ps1 == ps2
            _lib.check(lib.otp_fusion_stack_frames(
                _lib.dptr(rough), margin.data_ptr(), squeezed.data_ptr(), _lib.dptr(ctx),
                _lib.dptr(pe1, allow_none=True), _lib.dptr(pe2, allow_none=True), ps1, frames, b, j, t,
                x1.data_ptr(), x2.data_ptr(), intersection.data_ptr(), prev_b.data_ptr(), st),
                "otp_fusion_stack_frames")
            pprec = _lib.precision_code(refine)

            def encode(i, enc, fl, xin):
                stq = _lib.stream_ptr(dev)
                s0, s1, s2 = enc.forward_tokens(xin)
                if (pprec != _lib.PREC_FP32 and s1.shape[-1] * 2 == t and s2.shape[-1] * 4 == t
                        and lib.otp_pyramid_conv1x1_tc_supported(c8, t, j)):
                    _lib.check(lib.otp_pyramid_conv1x1_tc(
                        _lib.dptr(s0), _lib.dptr(s1), _lib.dptr(s2), b, c8, t, self._pyramid_packed(i, fl, pprec).data_ptr(),
                        _lib.dptr(fl.bias.detach()), j, cat.data_ptr() + 4 * i * j * t, 3 * j * t, pprec, stq),
                        "otp_pyramid_conv1x1_tc")
                else:
                    _lib.check(lib.otp_pyramid_conv1x1(
                        _lib.dptr(s0), _lib.dptr(s1), _lib.dptr(s2), b, c8, t, s1.shape[-1], s2.shape[-1],
                        _lib.dptr(fl.weight.detach().view(j, 3 * c8)), _lib.dptr(fl.bias.detach()), j,
                        cat.data_ptr() + 4 * i * j * t, 3 * j * t, stq), "otp_pyramid_conv1x1")

            # the two temporal encoders are independent: encoder 2 (and its pyramid conv) runs on the side
            # stream, so the ramp / tail of every persistent kernel of one is filled by CTAs of the other
            both = self.overlap_branches and self.overlap_encoders
            if both:
                side.wait_stream(main)                      # x2 is ready
                with torch.cuda.stream(side):
                    encode(1, self.temporal_encoder2, self.final_layer2, x2)
            encode(0, self.temporal_encoder1, self.final_layer1, x1)
            if not both:
                encode(1, self.temporal_encoder2, self.final_layer2, x2)
            if self.overlap_branches:
                main.wait_stream(side)
            del x1, x2                                       # x2 stays allocated until the side stream has joined
            trans = self.offset_mask_combine_conv(cat)
            if _debug is not None:
                _debug.update(cat=cat, trans=trans, def_heatmaps=def_heatmaps)
            cdef = trans.shape[1]
            ww = 1.0 / len(self.deformable_conv_dilations)
            prec = _lib.precision_code(refine)
            fused = prec != _lib.PREC_FP32 and j == 17 and cdef == 32
            if fused:
                # offsets / masks never reach HBM: implicit-GEMM conv on tcgen05 feeding the DCN from TMEM
                for i, dd in enumerate(self.deformable_conv_dilations):
                    dcn = self.modulated_deform_conv_list[i].deform_conv
                    _lib.check(lib.otp_offset_mask_dcn_forward(
                        self._offset_mask_packed(i).data_ptr(), _lib.dptr(trans), _lib.dptr(def_heatmaps),
                        _lib.dptr(dcn.weight.detach()), _lib.dptr(dcn.bias.detach() if dcn.bias is not None else None,
                                                                  allow_none=True),
                        out.data_ptr(), b, h, w, dd, ww, int(i > 0), prec, st), "otp_offset_mask_dcn_forward")
            else:
                from ..thirdparty.deform_conv import modulated_deform_conv
                k2 = 9
                offsets = torch.empty((b, 2 * k2 * j, h, w), **f32)
                masks = torch.empty((b, k2 * j, h, w), **f32)
                for i, dd in enumerate(self.deformable_conv_dilations):
                    for conv, dst in ((self.offsets_list[i][0], offsets), (self.masks_list[i][0], masks)):
                        _lib.check(lib.otp_conv2d(
                            _lib.dptr(trans), cdef * t, None, 0, _lib.dptr(conv.weight.detach()), None, None, 0,
                            dst.data_ptr(), dst.shape[1] * t, b, cdef, h, w, dst.shape[1], 3, dd, 0, st),
                            "otp_conv2d")
                    dcn = self.modulated_deform_conv_list[i].deform_conv
                    modulated_deform_conv(def_heatmaps, offsets, masks, dcn.weight, dcn.bias, dcn.stride,
                                          dcn.padding, dcn.dilation, dcn.groups, dcn.deformable_groups,
                                          alpha=ww, out=out, accumulate=(i > 0))
        return (out, rough_heatmaps, intersection, prev_b, ctx.view(b, j, h, w),
                squeezed.expand(b, j, h, w), total_b)
