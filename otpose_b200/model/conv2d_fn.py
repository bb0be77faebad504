"""Autograd wrapper of the small-channel ``otp_conv2d`` (stride 1, k in {1, 3}, dilation d, padding d*(k//2)):
the dilated offset / mask convs of the reference (``nn.Conv2d(..., dilation=(dd, dd), padding=(dd, dd))``,
model/OTPose.py:168-177) with a native backward, so that together with ``ModulatedDeformConvFunction`` the
whole offset/mask-conv + DCN stage trains without an ATen convolution.

forward      otp_conv2d
grad_input   otp_conv2d on grad_output with the weights transposed and flipped (same dilation / padding)
grad_weight, grad_bias   otp_conv2d_wgrad (deterministic two-pass reduction)
"""
from __future__ import annotations

import torch

from .. import _lib

__all__ = ["Conv2dFunction", "conv2d"]


_CIN_CHUNK = 48   # otp_conv2d keeps the (Cin*k*k, 16) weight tile of a CTA in shared memory: built for small Cin


def _launch_conv(x, weight, bias, dilation):
    """y = conv(x) + bias.  Wide inputs (the grad_input conv of a 32 -> 306 layer has Cin = 306) are run as a
    chain over input-channel chunks, each launch adding the previous partial through the kernel's fused
    residual input; the partials ping-pong between two buffers (the kernel reads its residual through the
    read-only path and declares it __restrict__, so a launch never reads what it writes)."""
    lib = _lib.load()
    b, cin, h, w = x.shape
    cout, k = weight.shape[0], weight.shape[2]
    bufs = [torch.empty((b, cout, h, w), dtype=torch.float32, device=x.device), None]
    cur = 0
    if b:
        p = h * w
        with torch.cuda.device(x.device):
            st = _lib.stream_ptr(x.device)
            for c0 in range(0, cin, _CIN_CHUNK):
                c1 = min(cin, c0 + _CIN_CHUNK)
                wc = weight if (c0 == 0 and c1 == cin) else weight[:, c0:c1].contiguous()
                if c0 == 0:
                    res, out = None, bufs[0]
                else:                      # the partial so far is the residual; the sum goes to the other buffer
                    if bufs[1 - cur] is None:
                        bufs[1 - cur] = torch.empty_like(bufs[cur])
                    res, out = bufs[cur], bufs[1 - cur]
                    cur = 1 - cur
                _lib.check(lib.otp_conv2d(x.data_ptr() + 4 * c0 * p, cin * p, None, 0, _lib.dptr(wc),
                                          _lib.dptr(bias, allow_none=True) if c0 == 0 else None,
                                          res.data_ptr() if res is not None else None, cout * p, out.data_ptr(),
                                          cout * p, b, c1 - c0, h, w, cout, k, dilation, 0, st), "otp_conv2d")
    return bufs[cur]


class Conv2dFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, dilation):
        _lib.require_cuda(x)
        if weight.shape[2] != weight.shape[3] or weight.shape[2] not in (1, 3):
            raise NotImplementedError("conv2d kernels: square kernels of size 1 or 3")
        x = x.contiguous().float()
        weight = weight.contiguous().float()
        bias = bias.contiguous().float() if bias is not None else None
        ctx.dilation = int(dilation)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight)
        return _launch_conv(x, weight, bias, ctx.dilation)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        lib = _lib.load()
        grad_out = grad_out.contiguous().float()
        b, cin, h, w = x.shape
        cout, k = weight.shape[0], weight.shape[2]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wt = weight.flip(2, 3).transpose(0, 1).contiguous()        # (cin, cout, k, k)
            gx = _launch_conv(grad_out, wt, None, ctx.dilation)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw = torch.empty_like(weight)
            gb = torch.empty(cout, dtype=torch.float32, device=x.device) if ctx.has_bias else None
            nws = lib.otp_conv2d_wgrad_workspace_bytes(max(b, 1), cin, h, w, cout, k)
            ws = _lib.workspace.get(nws, x.device, "conv_wgrad")
            with torch.cuda.device(x.device):
                _lib.check(lib.otp_conv2d_wgrad(_lib.dptr(x), cin * h * w, _lib.dptr(grad_out), cout * h * w,
                                                gw.data_ptr(), gb.data_ptr() if gb is not None else None, b, cin, h,
                                                w, cout, k, ctx.dilation, 0, ws.data_ptr(), ws.numel(),
                                                _lib.stream_ptr(x.device)), "otp_conv2d_wgrad")
        return gx, gw, gb, None


def conv2d(x, weight, bias=None, dilation=1):
    """``F.conv2d(x, weight, bias, stride=1, padding=dilation * (k // 2), dilation=dilation)`` on the library's
    kernels, differentiable in x, weight and bias."""
    return Conv2dFunction.apply(x, weight, bias, dilation)
