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
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                       for t in (input, offset, mask, weight, bias)):
        if out is not None or accumulate or alpha != 1.0:
            raise NotImplementedError("alpha/out/accumulate are inference-only extensions; "
                                      "call under torch.no_grad() or drop them to train")
        return ModulatedDeformConvFunction.apply(input, offset, mask, weight, bias, stride, padding, dilation,
                                                 groups, deformable_groups)
    return _mdcn_forward(input, offset, mask, weight, bias, stride, padding, dilation, groups,
                         deformable_groups, alpha, out, accumulate)


def _mdcn_forward(input, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups,
                  alpha=1.0, out=None, accumulate=False):
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


class ModulatedDeformConvFunction(torch.autograd.Function):
    """Autograd wrapper with the reference's argument order (functions/deform_conv.py:109-180):
    forward = ``otp_mdcn_forward``, backward = ``otp_mdcn_backward`` (one fused launch + a
    fixed-order reduction; deterministic, unlike the reference's atomicAdd col2im)."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                deformable_groups=1):
        ctx.conf = (stride, padding, dilation, groups, deformable_groups)
        ctx.with_bias = bias is not None
        input, offset, mask = input.contiguous(), offset.contiguous(), mask.contiguous()
        weight = weight.contiguous()
        ctx.save_for_backward(input, offset, mask, weight)
        return _mdcn_forward(input.detach(), offset.detach(), mask.detach(), weight.detach(),
                             bias.detach() if bias is not None else None, stride, padding, dilation, groups,
                             deformable_groups)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        input, offset, mask, weight = ctx.saved_tensors
        stride, padding, dilation, groups, dg = ctx.conf
        b, c, h, w = input.shape
        cout, _, kh, kw = weight.shape
        grad_output = grad_output.contiguous().float()
        grad_input, grad_offset = torch.empty_like(input), torch.empty_like(offset)
        grad_mask, grad_weight = torch.empty_like(mask), torch.empty_like(weight)
        grad_bias = weight.new_empty(cout) if ctx.with_bias else None
        lib = _lib.load()
        nbytes = lib.otp_mdcn_backward_workspace_bytes(b, c, h, w, cout, kh, kw, stride, padding, dilation)
        with torch.cuda.device(input.device):
            ws = _lib.workspace.get(nbytes, input.device, "mdcn_bwd")
            _lib.check(lib.otp_mdcn_backward(
                _lib.dptr(input), _lib.dptr(offset), _lib.dptr(mask), _lib.dptr(weight), _lib.dptr(grad_output),
                _lib.dptr(grad_input), _lib.dptr(grad_offset), _lib.dptr(grad_mask), _lib.dptr(grad_weight),
                _lib.dptr(grad_bias, allow_none=True), b, c, h, w, cout, kh, kw, stride, padding, dilation,
                groups, dg, ws.data_ptr(), nbytes, _lib.stream_ptr(input.device)), "otp_mdcn_backward")
        return (grad_input, grad_offset, grad_mask, grad_weight, grad_bias, None, None, None, None, None)


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
