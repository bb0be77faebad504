"""``ModulatedDeformConv`` / ``modulated_deform_conv`` with the reference's interface
(thirdparty/deform_conv/modules/deform_conv.py:85-131 and
functions/deform_conv.py:109-180), forward through ``otp_mdcn_forward``.
"""
from __future__ import annotations

import math

import torch
from torch import nn
from torch.nn.modules.utils import _pair

from ... import _lib


def _out_size(size, k, stride, padding, dilation):
    return (size + 2 * padding - (dilation * (k - 1) + 1)) // stride + 1


def modulated_deform_conv(input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1,
                          groups=1, deformable_groups=1, *, alpha=1.0, out=None, accumulate=False):
    """Functional form, same positional arguments as the reference's
    ``ModulatedDeformConvFunction.apply`` (functions/deform_conv.py:112-122).

    Keyword-only extensions: ``out = (out if accumulate else 0) + alpha * result``
    lets a caller fuse the 0.2-weighted sum over dilations (model/OTPose.py:387-392).
    Raises ``NotImplementedError`` for non-CUDA input like the reference (:131-132).
    """
    if not input.is_cuda:
        raise NotImplementedError
    if any(t is not None and t.requires_grad and torch.is_grad_enabled()
           for t in (input, offset, mask, weight, bias)):
        raise NotImplementedError("modulated_deform_conv backward is a 'next' row (SURVEY.md 8 a12); "
                                  "call under torch.no_grad()")
    b, c, h, w = input.shape
    cout, cin_g, kh, kw = weight.shape
    if c != cin_g * groups:
        raise RuntimeError(f"Input shape and kernel channels wont match: ({c} vs {cin_g * groups}).")
    ho, wo = _out_size(h, kh, stride, padding, dilation), _out_size(w, kw, stride, padding, dilation)
    if tuple(offset.shape) != (b, deformable_groups * 2 * kh * kw, ho, wo):
        raise RuntimeError(f"offset shape {tuple(offset.shape)} != {(b, deformable_groups * 2 * kh * kw, ho, wo)}")
    if tuple(mask.shape) != (b, deformable_groups * kh * kw, ho, wo):
        raise RuntimeError(f"mask shape {tuple(mask.shape)} != {(b, deformable_groups * kh * kw, ho, wo)}")
    if out is None:
        assert not accumulate
        out = input.new_empty((b, cout, ho, wo))
    lib = _lib.load()
    with torch.cuda.device(input.device):
        _lib.check(lib.otp_mdcn_forward(
            _lib.dptr(input), _lib.dptr(offset), _lib.dptr(mask), _lib.dptr(weight.detach()),
            _lib.dptr(bias.detach() if bias is not None else None, allow_none=True), _lib.dptr(out),
            b, c, h, w, cout, kh, kw, stride, padding, dilation, groups, deformable_groups,
            float(alpha), int(accumulate), _lib.stream_ptr(input.device)), "otp_mdcn_forward")
    return out


class ModulatedDeformConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deformable_groups=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deformable_groups, self.with_bias = groups, deformable_groups, bias
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1. / math.sqrt(n)
        self.weight.data.uniform_(-stdv, stdv)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, offset, mask):
        return modulated_deform_conv(x, offset, mask, self.weight, self.bias, self.stride, self.padding,
                                     self.dilation, self.groups, self.deformable_groups)


def deform_conv(*a, **k):
    raise NotImplementedError("unmodulated DeformConv (v1) is never constructed by OTPose and is not built")


class DeformConv(nn.Module):
    """Name kept for ``isinstance`` checks in the reference's init_weights (model/OTPose.py:449)."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("unmodulated DeformConv (v1) is never constructed by OTPose and is not built")


class DeformConvPack(DeformConv):
    pass


class ModulatedDeformConvPack(ModulatedDeformConv):
    def __init__(self, *a, **k):
        raise NotImplementedError("ModulatedDeformConvPack is never constructed by OTPose and is not built")
