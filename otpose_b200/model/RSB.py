"""Drop-in ``CHAIN_RSB_BLOCKS`` / ``RSB_BLOCK`` / ``conv_bn_relu`` (reference
model/RSB.py:10-139) -- same constructors and state-dict keys; eval-mode forward
through ``otp_conv2d`` with BatchNorm folded into the conv weights and the
branch adds / residual / ReLU / channel concat fused into the conv launches.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib

__all__ = ["RSB_BLOCK", "CHAIN_RSB_BLOCKS", "conv_bn_relu"]


class conv_bn_relu(nn.Module):
    def __init__(self, in_planes, out_planes, kernel_size, stride, padding, has_bn=True, has_relu=True,
                 efficient=False, groups=1):
        super().__init__()
        if stride != 1 or groups != 1 or kernel_size not in (1, 3) or padding != kernel_size // 2:
            raise NotImplementedError("conv_bn_relu kernels: stride 1, groups 1, k in {1,3}, 'same' padding")
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              groups=groups)
        self.has_bn, self.has_relu, self.efficient = has_bn, has_relu, efficient
        self.bn = nn.BatchNorm2d(out_planes)
        self.relu = nn.ReLU(inplace=True)
        self.kernel_size = kernel_size
        self._folded = None
        self._folded_key = None
        self._tc_packed = {}
        self.precision = "fp32"     # "fp16" / "bf16": tcgen05 implicit-GEMM path (set by OTPose)

    def folded(self):
        """(weight, bias) with the eval-mode BatchNorm affine folded in (RSB.py:120-131)."""
        ps = [self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias, self.bn.running_mean,
              self.bn.running_var]
        key = tuple((id(p), p.data_ptr(), p._version) for p in ps)   # the entry below keeps `ps` alive
        if self._folded is None or key != self._folded_key:
            lib = _lib.load()
            wsrc, dev = self.conv.weight.detach(), self.conv.weight.device
            _lib.require_cuda(wsrc)
            w = torch.empty_like(wsrc, dtype=torch.float32, memory_format=torch.contiguous_format)
            b = torch.empty(wsrc.shape[0], dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):      # one library launch (packing stays inside the library)
                _lib.check(lib.otp_conv_bn_fold(
                    _lib.dptr(wsrc), _lib.dptr(self.conv.bias.detach() if self.conv.bias is not None else None,
                                               allow_none=True),
                    _lib.dptr(self.bn.weight.detach()), _lib.dptr(self.bn.bias.detach()),
                    _lib.dptr(self.bn.running_mean), _lib.dptr(self.bn.running_var), float(self.bn.eps),
                    int(self.has_bn), wsrc.shape[0], wsrc[0].numel(), w.data_ptr(), b.data_ptr(),
                    _lib.stream_ptr(dev)), "otp_conv_bn_fold")
            self._folded = (w, b)
            self._folded_key = key
            self._folded_src = ps
        return self._folded

    def run(self, x, x_bs, out, out_bs, b, h, w, x_add=None, x_add_bs=0, residual=None, residual_bs=0,
            relu=None):
        """Launch on raw (pointer, batch-stride) channel slices."""
        wt, bias = self.folded()
        relu = self.has_relu if relu is None else relu
        lib = _lib.load()
        cin, cout, k = self.conv.in_channels, self.conv.out_channels, self.kernel_size
        if self.precision != "fp32" and lib.otp_conv2d_tc_supported(cin, cout, h, w, k):
            prec = _lib.precision_code(self.precision)
            key = (self._folded_key, prec)
            if self._tc_packed.get("key") != key:
                nbytes = lib.otp_conv2d_tc_pack_bytes(cin, cout, k)
                buf = torch.empty(nbytes, dtype=torch.uint8, device=wt.device)
                _lib.check(lib.otp_conv2d_tc_pack(wt.data_ptr(), cin, cout, k, prec, buf.data_ptr(), nbytes,
                                                  _lib.stream_ptr(wt.device)), "otp_conv2d_tc_pack")
                self._tc_packed = {"key": key, "buf": buf}
            _lib.check(lib.otp_conv2d_tc(x, x_bs, x_add, x_add_bs, self._tc_packed["buf"].data_ptr(), bias.data_ptr(),
                                         residual, residual_bs, out, out_bs, b, cin, h, w, cout, k, int(relu), prec,
                                         _lib.stream_ptr(wt.device)), "otp_conv2d_tc")
            return
        _lib.check(lib.otp_conv2d(x, x_bs, x_add, x_add_bs, wt.data_ptr(), bias.data_ptr(), residual,
                                  residual_bs, out, out_bs, b, self.conv.in_channels, h, w,
                                  self.conv.out_channels, self.kernel_size, 1, int(relu),
                                  _lib.stream_ptr(wt.device)), "otp_conv2d")

    def forward(self, x):
        if self.training:      # differentiable path (train_ops: library ops under autograd, BN batch statistics)
            from . import train_ops
            _lib.require_cuda(x)
            return train_ops.conv_bn_relu(self, x)
        _lib.require_cuda(x)
        b, c, h, w = x.shape
        y = torch.empty((b, self.conv.out_channels, h, w), dtype=torch.float32, device=x.device)
        if b:
            with torch.cuda.device(x.device):
                self.run(_lib.dptr(x), c * h * w, y.data_ptr(), y.shape[1] * h * w, b, h, w)
        return y


class RSB_BLOCK(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1, groups=1, downsample=None, efficient=False):
        super().__init__()
        bc = self.branch_ch = in_planes * 26 // 64
        self.conv_bn_relu1 = conv_bn_relu(in_planes, 4 * bc, 1, stride, 0, groups=groups)
        for name in ("1_1", "2_1", "2_2", "3_1", "3_2", "3_3", "4_1", "4_2", "4_3", "4_4"):
            setattr(self, "conv_bn_relu2_" + name, conv_bn_relu(bc, bc, 3, 1, 1, groups=groups))
        self.conv_bn_relu3 = conv_bn_relu(4 * bc, planes * self.expansion, 1, 1, 0, groups=groups, has_relu=False)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.planes = planes
        self.fused = True           # 16-bit modes: the whole block in one / two launches (csrc/rsb_fused.cu)
        self._fused_packed = None
        self._fused_key = None

    _ORDER = ("conv_bn_relu1", "conv_bn_relu2_1_1", "conv_bn_relu2_2_1", "conv_bn_relu2_2_2", "conv_bn_relu2_3_1",
              "conv_bn_relu2_3_2", "conv_bn_relu2_3_3", "conv_bn_relu2_4_1", "conv_bn_relu2_4_2", "conv_bn_relu2_4_3",
              "conv_bn_relu2_4_4", "conv_bn_relu3")

    def _forward_fused(self, x):
        """RSB.py:81-103 in the 16-bit modes: ``otp_rsb_block_forward`` on the BN-folded convs (IEEE half operands,
        fp32 accumulate; every intermediate map stays in shared memory)."""
        import ctypes as C
        lib = _lib.load()
        b, cin, h, w = x.shape
        dev = x.device
        convs = [getattr(self, n) for n in self._ORDER] + ([self.downsample] if self.downsample is not None else [])
        folded = [m.folded() for m in convs]
        key = tuple(m._folded_key for m in convs)
        has_ds = int(self.downsample is not None)
        with torch.cuda.device(dev):
            if self._fused_packed is None or key != self._fused_key:
                nbytes = lib.otp_rsb_block_pack_bytes(cin, self.planes, has_ds)
                buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                wp = (C.c_void_p * 13)(*([f[0].data_ptr() for f in folded] + [None] * (13 - len(folded))))
                bp = (C.c_void_p * 13)(*([f[1].data_ptr() for f in folded] + [None] * (13 - len(folded))))
                _lib.check(lib.otp_rsb_block_pack(wp, bp, cin, self.planes, has_ds, buf.data_ptr(), nbytes,
                                                  _lib.stream_ptr(dev)), "otp_rsb_block_pack")
                self._fused_packed, self._fused_key = buf, key
            xc = x if x.is_contiguous() else x.contiguous()
            out = torch.empty((b, self.planes, h, w), dtype=torch.float32, device=dev)
            nws = lib.otp_rsb_block_workspace_bytes(b, cin, self.planes, h, w)
            ws = _lib.workspace.get(nws, dev, "rsb")
            _lib.check(lib.otp_rsb_block_forward(self._fused_packed.data_ptr(), _lib.dptr(xc), cin * h * w,
                                                 out.data_ptr(), self.planes * h * w, b, cin, self.planes, has_ds, h, w,
                                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                       "otp_rsb_block_forward")
        return out

    def forward(self, x):
        """RSB.py:81-103.  Intermediate layout: `spx` (B,4bc,H,W) and one (B,10bc,H,W)
        buffer holding the ten branch outputs; the concat (out_1_1, out_2_2, out_3_3,
        out_4_4) is assembled by writing those four convs into a (B,4bc,H,W) buffer."""
        if self.training:
            from . import train_ops
            _lib.require_cuda(x)
            return train_ops.rsb_block(self, x)
        _lib.require_cuda(x)
        b, cin, h, w = x.shape
        if self.fused and self.conv_bn_relu1.precision != "fp32" and b > 0 and x.dtype == torch.float32 \
                and _lib.load().otp_rsb_block_supported(cin, self.planes, h, w):
            return self._forward_fused(x)
        p = h * w
        bc = self.branch_ch
        dev = x.device
        spx = torch.empty((b, 4 * bc, h, w), dtype=torch.float32, device=dev)
        tmp = torch.empty((6, b, bc, h, w), dtype=torch.float32, device=dev)   # 2_1 3_1 3_2 4_1 4_2 4_3
        cat = torch.empty((b, 4 * bc, h, w), dtype=torch.float32, device=dev)  # 1_1 2_2 3_3 4_4
        out = torch.empty((b, self.planes, h, w), dtype=torch.float32, device=dev)
        if b == 0:
            return out
        xp = _lib.dptr(x)
        sp = lambda i: spx.data_ptr() + 4 * i * bc * p           # noqa: E731  channel slice i of spx
        ct = lambda i: cat.data_ptr() + 4 * i * bc * p           # noqa: E731
        tp = lambda i: tmp[i].data_ptr()                         # noqa: E731
        S4, S1 = 4 * bc * p, bc * p
        m = lambda n: getattr(self, "conv_bn_relu2_" + n)        # noqa: E731
        with torch.cuda.device(dev):
            self.conv_bn_relu1.run(xp, cin * p, spx.data_ptr(), S4, b, h, w)
            m("1_1").run(sp(0), S4, ct(0), S4, b, h, w)                                  # out_1_1
            m("2_1").run(sp(1), S4, tp(0), S1, b, h, w, x_add=ct(0), x_add_bs=S4)        # out_2_1
            m("2_2").run(tp(0), S1, ct(1), S4, b, h, w)                                  # out_2_2
            m("3_1").run(sp(2), S4, tp(1), S1, b, h, w, x_add=tp(0), x_add_bs=S1)        # out_3_1
            m("3_2").run(tp(1), S1, tp(2), S1, b, h, w, x_add=ct(1), x_add_bs=S4)        # out_3_2
            m("3_3").run(tp(2), S1, ct(2), S4, b, h, w)                                  # out_3_3
            m("4_1").run(sp(3), S4, tp(3), S1, b, h, w, x_add=tp(1), x_add_bs=S1)        # out_4_1
            m("4_2").run(tp(3), S1, tp(4), S1, b, h, w, x_add=tp(2), x_add_bs=S1)        # out_4_2
            m("4_3").run(tp(4), S1, tp(5), S1, b, h, w, x_add=ct(2), x_add_bs=S4)        # out_4_3
            m("4_4").run(tp(5), S1, ct(3), S4, b, h, w)                                  # out_4_4
            if self.downsample is not None:
                skip = self.downsample(x)
                rp, rbs = skip.data_ptr(), self.planes * p
            else:
                rp, rbs = xp, cin * p
            self.conv_bn_relu3.run(cat.data_ptr(), S4, out.data_ptr(), self.planes * p, b, h, w,
                                   residual=rp, residual_bs=rbs, relu=True)
        return out


class CHAIN_RSB_BLOCKS(nn.Module):
    def __init__(self, in_planes, out_planes, num_blocks, groups=1):
        super().__init__()
        ds = conv_bn_relu(in_planes, out_planes, kernel_size=1, stride=1, padding=0, has_relu=False, groups=groups)
        layers = [RSB_BLOCK(in_planes, out_planes, 1, downsample=ds)]
        for _ in range(1, num_blocks):
            layers.append(RSB_BLOCK(out_planes, out_planes, 1, downsample=None))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)
