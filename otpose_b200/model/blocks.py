"""Drop-in module interfaces of the reference's ``model/blocks.py`` for the classes
the OTPose head instantiates: ``LayerNorm`` (blocks.py:67-110), ``MaskedMHCA``
(:319-452), ``AffineDropPath`` (:283-298), ``TransformerBlock`` (:185-280) and
``get_sinusoid_encoding`` (:114-125).

Same constructor signatures, parameter names and shapes (so reference
checkpoints load with ``load_state_dict``); the arithmetic of a whole
TransformerBlock runs in the fused CUDA passes behind ``otp_block_forward``
(include/otpose_b200.h).  The sub-modules are parameter containers: their math
is fused into the block kernels and they have no stand-alone forward.

``LocalMaskedMHCA`` / ``MaskedConv1D`` / ``Scale`` are never constructed by
OTPose (``mha_win_size = [-1]*6``) and are not provided.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from .. import _lib
from ..utils.synthetic import sinusoid_table


def get_sinusoid_encoding(n_position, d_hid):
    """(1, d_hid, n_position) sinusoid table (reference model/blocks.py:114-125)."""
    return sinusoid_table(n_position, d_hid)


class _Fused(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - documented contract
        raise NotImplementedError(
            f"{type(self).__name__} is fused into TransformerBlock's CUDA passes; call the block")


class LayerNorm(_Fused):
    """Channel LayerNorm over (B, C, T): weight / bias of shape (1, C, 1)."""

    def __init__(self, num_channels, eps=1e-5, affine=True, device=None, dtype=None):
        super().__init__()
        kw = {"device": device, "dtype": dtype}
        self.num_channels, self.eps, self.affine = num_channels, eps, affine
        if eps != 1e-5:
            raise NotImplementedError("fused LayerNorm kernels are built for eps = 1e-5")
        if affine:
            self.weight = nn.Parameter(torch.ones([1, num_channels, 1], **kw))
            self.bias = nn.Parameter(torch.zeros([1, num_channels, 1], **kw))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)


class AffineDropPath(_Fused):
    """Per-channel residual scale (+ stochastic depth in training)."""

    def __init__(self, num_dim, drop_prob=0.0, init_scale_value=1e-4):
        super().__init__()
        self.scale = nn.Parameter(init_scale_value * torch.ones((1, num_dim, 1)), requires_grad=True)
        self.drop_prob = drop_prob


class MaskedMHCA(_Fused):
    """Depthwise-conv + LayerNorm + pointwise q/k/v, channel-Gram attention, proj."""

    def __init__(self, n_embd, n_head, n_qx_stride=1, n_kv_stride=1, attn_pdrop=0.0, proj_pdrop=0.0):
        super().__init__()
        assert n_embd % n_head == 0
        assert (n_qx_stride == 1) or (n_qx_stride % 2 == 0)
        assert (n_kv_stride == 1) or (n_kv_stride % 2 == 0)
        self.n_embd, self.n_head = n_embd, n_head
        self.n_channels = n_embd // n_head
        self.scale = 1.0 / (self.n_channels ** 0.5)
        self.n_qx_stride, self.n_kv_stride = n_qx_stride, n_kv_stride

        def dw(stride):
            k = stride + 1 if stride > 1 else 3
            return nn.Conv1d(n_embd, n_embd, k, stride=stride, padding=k // 2, groups=n_embd, bias=False)

        self.query_conv = dw(n_kv_stride)   # the reference strides q by n_kv_stride (blocks.py:360)
        self.query_norm = LayerNorm(n_embd)
        self.key_conv = dw(n_kv_stride)
        self.key_norm = LayerNorm(n_embd)
        self.value_conv = dw(n_kv_stride)
        self.value_norm = LayerNorm(n_embd)
        self.key = nn.Conv1d(n_embd, n_embd, 1)
        self.query = nn.Conv1d(n_embd, n_embd, 1)
        self.value = nn.Conv1d(n_embd, n_embd, 1)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.proj_drop = nn.Dropout(proj_pdrop)
        self.proj = nn.Conv1d(n_embd, n_embd, 1)


class TransformerBlock(nn.Module):
    """Pre-LN block: ``u = pool_skip(x) + s_a*attn(ln1(x)); y = u + s_m*mlp(ln2(u))``."""

    def __init__(self, n_embd, n_head, n_ds_strides=(1, 1), n_out=None, n_hidden=None, act_layer=nn.GELU,
                 attn_pdrop=0.0, proj_pdrop=0.0, path_pdrop=0.0, mha_win_size=-1, use_rel_pe=False):
        super().__init__()
        assert len(n_ds_strides) == 2
        if mha_win_size > 1:
            raise NotImplementedError("LocalMaskedMHCA (mha_win_size > 1) is not used by OTPose")
        if n_ds_strides[0] != n_ds_strides[1] or n_ds_strides[0] not in (1, 2):
            raise NotImplementedError("block kernels are built for strides (1,1) and (2,2)")
        if (n_out not in (None, n_embd)) or (n_hidden not in (None, 4 * n_embd)) or act_layer is not nn.GELU:
            raise NotImplementedError("block kernels are built for n_out=n_embd, n_hidden=4*n_embd, GELU")
        self.n_embd, self.n_head, self.stride = n_embd, n_head, n_ds_strides[0]
        self.ln1 = LayerNorm(n_embd)
        self.ln2 = LayerNorm(n_embd)
        self.attn = MaskedMHCA(n_embd, n_head, n_qx_stride=n_ds_strides[0], n_kv_stride=n_ds_strides[1],
                               attn_pdrop=attn_pdrop, proj_pdrop=proj_pdrop)
        if n_ds_strides[0] > 1:
            k, s = n_ds_strides[0] + 1, n_ds_strides[0]
            self.pool_skip = nn.MaxPool1d(k, stride=s, padding=(s + 1) // 2)
        else:
            self.pool_skip = nn.Identity()
        self.mlp = nn.Sequential(nn.Conv1d(n_embd, 4 * n_embd, 1), act_layer(), nn.Dropout(proj_pdrop),
                                 nn.Conv1d(4 * n_embd, n_embd, 1), nn.Dropout(proj_pdrop))
        if path_pdrop > 0.0:
            self.drop_path_attn = AffineDropPath(n_embd, drop_prob=path_pdrop)
            self.drop_path_mlp = AffineDropPath(n_embd, drop_prob=path_pdrop)
        else:
            self.drop_path_attn = nn.Identity()
            self.drop_path_mlp = nn.Identity()
        self._packed = None
        self._packed_key = None

    # ---- weight packing ----------------------------------------------------
    def _param_list(self):
        a = self.attn
        sa = getattr(self.drop_path_attn, "scale", None)
        sm = getattr(self.drop_path_mlp, "scale", None)
        return [("ln1_w", self.ln1.weight), ("ln1_b", self.ln1.bias), ("ln2_w", self.ln2.weight),
                ("ln2_b", self.ln2.bias), ("q_conv_w", a.query_conv.weight), ("k_conv_w", a.key_conv.weight),
                ("v_conv_w", a.value_conv.weight), ("q_norm_w", a.query_norm.weight),
                ("q_norm_b", a.query_norm.bias), ("k_norm_w", a.key_norm.weight), ("k_norm_b", a.key_norm.bias),
                ("v_norm_w", a.value_norm.weight), ("v_norm_b", a.value_norm.bias),
                ("q_w", a.query.weight), ("q_b", a.query.bias), ("k_w", a.key.weight), ("k_b", a.key.bias),
                ("v_w", a.value.weight), ("v_b", a.value.bias), ("proj_w", a.proj.weight),
                ("proj_b", a.proj.bias), ("mlp0_w", self.mlp[0].weight), ("mlp0_b", self.mlp[0].bias),
                ("mlp3_w", self.mlp[3].weight), ("mlp3_b", self.mlp[3].bias),
                ("scale_attn", sa), ("scale_mlp", sm)]

    def packed_weights(self) -> torch.Tensor:
        """Packed device copy of this block's weights, rebuilt when any parameter changed."""
        plist = self._param_list()
        # (address, version) per parameter; the cache entry keeps the source tensors alive, so an address
        # cannot be handed to a replacement Parameter while the entry exists (init_weights re-binds parameters)
        key = tuple((id(p), p.data_ptr(), p._version) if p is not None else None for _, p in plist)
        if self._packed is None or key != self._packed_key:
            lib = _lib.load()
            dev = self.ln1.weight.device
            params = _lib.BlockParams()
            for name, p in plist:
                setattr(params, name, _lib.dptr(p.detach() if p is not None else None,
                                                allow_none=name.startswith("scale_")))
            nbytes = lib.otp_block_packed_bytes(self.n_embd, self.n_head)
            if nbytes == 0:
                _lib.check(2, "otp_block_packed_bytes")
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.otp_block_pack(C.byref(params), self.n_embd, self.n_head, buf.data_ptr(),
                                              nbytes, _lib.stream_ptr(dev)), "otp_block_pack")
            self._packed, self._packed_key = buf, key
            self._packed_src = [p for _, p in plist]
        return self._packed

    def out_len(self, t: int) -> int:
        return t if self.stride == 1 else (t - 1) // 2 + 1

    def forward(self, x, pos_embd=None, precision="fp32"):
        """x (B, C, T) fp32 CUDA -> (B, C, T') (reference model/blocks.py:264-279).  Eval mode: the fused
        CUDA passes.  Training mode: the differentiable library-op composition of ``train_ops`` (dropout,
        drop-path; see that module's status note)."""
        if self.training:
            from . import train_ops
            _lib.require_cuda(x)          # same contract as the kernels: no CPU path, in either mode
            return train_ops.transformer_block(self, x)
        if pos_embd is not None:
            raise NotImplementedError("pos_embd argument is unused by ConvTransformer")
        _lib.require_cuda(x)
        b, c, t = x.shape
        assert c == self.n_embd
        lib = _lib.load()
        prec = _lib.precision_code(precision)
        y = torch.empty((b, c, self.out_len(t)), dtype=torch.float32, device=x.device)
        if b == 0:
            return y
        packed = self.packed_weights()
        nws = lib.otp_block_workspace_bytes(b, c, t, self.n_head, self.stride, prec)
        ws = _lib.workspace.get(nws, x.device, "block")
        with torch.cuda.device(x.device):
            _lib.check(lib.otp_block_forward(packed.data_ptr(), _lib.dptr(x), y.data_ptr(), b, c, t,
                                             self.n_head, self.stride, prec, ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr(x.device)), "otp_block_forward")
        return y
