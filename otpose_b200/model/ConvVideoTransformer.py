"""Drop-in ``ConvTransformer`` (reference model/ConvVideoTransformer.py:16-184).

Same constructor, same state-dict keys (``pos_embd``, ``stem.{i}.*``,
``branch.{i}.*``) and the same return value -- a tuple of ``1 + arch[2]``
``(B, C, T)`` maps, the branch outputs linearly upsampled back to ``T`` -- with
every block running through the C-ABI CUDA passes.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib
from .blocks import TransformerBlock, get_sinusoid_encoding


class ConvTransformer(nn.Module):
    def __init__(self, n_in, n_embd, n_head, n_embd_ks, max_len, arch, mha_win_size=[-1] * 6, h=72,
                 scale_factor=2, with_ln=True, attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0,
                 use_abs_pe=True, use_rel_pe=False, precision="fp32"):
        super().__init__()
        assert len(arch) == 3
        if arch[0] != 0:
            raise NotImplementedError("conv embedding stage (arch[0] > 0) is not used by OTPose")
        if n_in != n_embd:
            raise ValueError("without an embedding stage n_in must equal n_embd")
        if scale_factor != 2:
            raise NotImplementedError("branch kernels are built for scale_factor = 2")
        self.arch, self.max_len, self.scale_factor = arch, max_len, scale_factor
        self.n_embd, self.n_head, self.h = n_embd, n_head, h
        self.use_abs_pe, self.use_rel_pe = use_abs_pe, use_rel_pe
        self.fpn_strides = [scale_factor ** i for i in range(arch[-1] + 1)]
        self.mha_win_size = [mha_win_size] * len(self.fpn_strides) if isinstance(mha_win_size, int) \
            else mha_win_size
        self.precision = precision
        self.fused_stem = True      # 16-bit modes, C = 17: one cluster launch for the whole stem (block_flow.cu)
        if use_abs_pe:
            self.register_buffer("pos_embd", get_sinusoid_encoding(max_len, n_embd) / (n_embd ** 0.5))
        self.embd = nn.ModuleList()
        self.embd_norm = nn.ModuleList()
        self.stem = nn.ModuleList(
            TransformerBlock(n_embd, n_head, n_ds_strides=(1, 1), attn_pdrop=attn_pdrop, proj_pdrop=proj_pdrop,
                             path_pdrop=path_pdrop, mha_win_size=self.mha_win_size[0], use_rel_pe=use_rel_pe)
            for _ in range(arch[1]))
        self.branch = nn.ModuleList(
            TransformerBlock(n_embd, n_head, n_ds_strides=(scale_factor, scale_factor), attn_pdrop=attn_pdrop,
                             proj_pdrop=proj_pdrop, path_pdrop=path_pdrop,
                             mha_win_size=self.mha_win_size[1 + i], use_rel_pe=use_rel_pe)
            for i in range(arch[2]))
        self.upsample = nn.ModuleList(nn.Upsample(scale_factor=2 ** (i + 1), mode="linear")
                                      for i in range(arch[2]))
        for m in self.modules():   # reference __init_weights__: Conv1d bias = 0
            if isinstance(m, nn.Conv1d) and m.bias is not None:
                nn.init.constant_(m.bias, 0.0)

    # ------------------------------------------------------------------
    def pos_embd_for(self, t: int):
        """(pe (C, >=t) contiguous rows, row stride) as the eval forward uses it
        (ConvVideoTransformer.py:147-155: re-interpolated when t >= max_len)."""
        if not self.use_abs_pe:
            return None, 0
        pe = self.pos_embd
        if t >= self.max_len and t != pe.shape[-1]:
            pe = F.interpolate(pe, t, mode="linear", align_corners=False)
        return pe[0], pe.shape[-1]

    def forward_tokens(self, x):
        """Blocks only: x (B, C, T) with the positional embedding already added.
        Returns the raw pyramid [(B,C,T), (B,C,T/2), (B,C,T/4)] (no upsampling)."""
        for blk in self.stem:
            x = blk(x, precision=self.precision)
        outs = [x]
        for blk in self.branch:
            x = blk(x, precision=self.precision)
            outs.append(x)
        return outs

    def forward(self, x):
        if self.training:      # differentiable path (train_ops: library ops under autograd)
            from . import train_ops
            _lib.require_cuda(x)
            return train_ops.conv_transformer(self, x)
        _lib.require_cuda(x)
        b, c, h, w = x.shape
        t = h * w
        lib = _lib.load()
        x = x.reshape(b, c, t)
        if not x.is_contiguous():
            x = x.contiguous()
        st = _lib.stream_ptr(x.device)
        if (self.precision != "fp32" and b > 0 and not self.branch and len(self.stem) > 0 and self.fused_stem
                and lib.otp_flow_encoder_supported(c, self.n_head, t, len(self.stem))):
            # the narrow flow encoder in the 16-bit modes: positional embedding + every stem block in ONE
            # cluster launch (csrc/block_flow.cu); operands are IEEE half in both 16-bit modes
            import ctypes as C
            pe, stride = self.pos_embd_for(t) if self.use_abs_pe else (None, 0)
            packs = [blk.packed_weights() for blk in self.stem]
            ptrs = (C.c_void_p * len(packs))(*[p_.data_ptr() for p_ in packs])
            nws = lib.otp_flow_encoder_workspace_bytes(b, t)
            ws = _lib.workspace.get(nws, x.device, "flow")
            y = torch.empty_like(x)
            with torch.cuda.device(x.device):
                _lib.check(lib.otp_flow_encoder_forward(ptrs, len(packs), _lib.dptr(x), _lib.dptr(pe, allow_none=True),
                                                        stride, y.data_ptr(), b, t, ws.data_ptr(), ws.numel(), st),
                           "otp_flow_encoder_forward")
            return (y,)
        with torch.cuda.device(x.device):
            if self.use_abs_pe and b > 0:
                pe, stride = self.pos_embd_for(t)
                y = torch.empty_like(x)
                _lib.check(lib.otp_add_pos_embd(_lib.dptr(x), _lib.dptr(pe), stride, y.data_ptr(), b, c, t, st),
                           "otp_add_pos_embd")
                x = y
            pyr = self.forward_tokens(x)
            outs = (pyr[0],)
            for i, s in enumerate(pyr[1:]):
                scale = 2 ** (i + 1)
                up = torch.empty((b, c, s.shape[-1] * scale), dtype=torch.float32, device=x.device)
                if b > 0:
                    _lib.check(lib.otp_upsample_linear(_lib.dptr(s), up.data_ptr(), b, c, s.shape[-1], scale, st),
                               "otp_upsample_linear")
                outs += (up,)
        return outs
