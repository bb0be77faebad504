"""GPU: pins the tcgen05 shared-memory descriptor semantics the tensor-core kernels
rely on (core-matrix interleaved layout, K-major and MN-major views, K stepping,
accumulation) with a one-CTA self-test GEMM checked against numpy."""
import numpy as np
import pytest
import torch

from otpose_b200 import _lib

pytestmark = pytest.mark.gpu


def bf16_round(x):
    return torch.from_numpy(x).to(torch.bfloat16)


def cm_image(mat: torch.Tensor, rs: int, cs: int) -> torch.Tensor:
    """bf16 (R, K) -> byte image with addr(r,k) = (r/8)*rs + (k/8)*cs + (r%8)*16 + (k%8)*2."""
    r, k = mat.shape
    rr, kk = np.meshgrid(np.arange(r), np.arange(k), indexing="ij")
    off = (rr // 8) * rs + (kk // 8) * cs + (rr % 8) * 16 + (kk % 8) * 2
    size = int(off.max()) + 2
    size = (size + 15) // 16 * 16
    img = torch.zeros(size // 2, dtype=torch.int16)
    img[torch.from_numpy(off.reshape(-1) // 2)] = mat.contiguous().view(torch.int16).reshape(-1)
    return img


def run(a_img, b_img, n, ksteps, a, b, repeat=1):
    lib = _lib.load()
    d = torch.full((128, n), float("nan"), device="cuda")
    ai, bi = a_img.cuda(), b_img.cuda()
    _lib.check(lib.otp_debug_umma_gemm(ai.data_ptr(), ai.numel() * 2, bi.data_ptr(), bi.numel() * 2, d.data_ptr(),
                                       n, ksteps, a["off"], a["lbo"], a["sbo"], a["kstep"], b["off"], b["lbo"],
                                       b["sbo"], b["kstep"], a["mn"], b["mn"], repeat, None), "otp_debug_umma_gemm")
    torch.cuda.synchronize()
    return d.cpu()


def relerr(x, ref):
    return float((x - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("n,k", [(144, 144), (96, 144), (80, 64), (16, 16), (256, 96)])
@pytest.mark.parametrize("row_major_groups", [True, False])
def test_k_major_operands(n, k, row_major_groups):
    r = np.random.default_rng(n + k)
    a = bf16_round(r.standard_normal((128, k)).astype(np.float32))
    b = bf16_round(r.standard_normal((n, k)).astype(np.float32))
    ref = a.float() @ b.float().T
    if row_major_groups:      # [row group][k chunk][8][16B]
        a_rs, a_cs, b_rs, b_cs = (k // 8) * 128, 128, (k // 8) * 128, 128
    else:                     # [k chunk][row group][8][16B]
        a_rs, a_cs, b_rs, b_cs = 128, 16 * 128, 128, (n // 8) * 128
    got = run(cm_image(a, a_rs, a_cs), cm_image(b, b_rs, b_cs), n, k // 16,
              dict(off=0, lbo=a_cs, sbo=a_rs, kstep=2 * a_cs, mn=0), dict(off=0, lbo=b_cs, sbo=b_rs, kstep=2 * b_cs, mn=0))
    assert relerr(got, ref) < 1e-5
    got2 = run(cm_image(a, a_rs, a_cs), cm_image(b, b_rs, b_cs), n, k // 16,
               dict(off=0, lbo=a_cs, sbo=a_rs, kstep=2 * a_cs, mn=0),
               dict(off=0, lbo=b_cs, sbo=b_rs, kstep=2 * b_cs, mn=0), repeat=3)
    assert relerr(got2, 3 * ref) < 1e-5


def test_mn_major_view_of_the_same_tile_gram():
    """Gram over tokens: D[i, j] = sum_t Q[t, c0+i] * K[t, c0+j] from [token][channel] tiles."""
    r = np.random.default_rng(3)
    ch = 192
    q = bf16_round(r.standard_normal((128, ch)).astype(np.float32))
    kk = bf16_round(r.standard_normal((128, ch)).astype(np.float32))
    rs, cs = (ch // 8) * 128, 128
    qi, ki = cm_image(q, rs, cs), cm_image(kk, rs, cs)
    c0, n = 64, 80
    ref = q.float()[:, c0:c0 + 128].T @ kk.float()[:, c0:c0 + n]
    # MN-major view: SBO = stride between 8-element MN chunks (cs), LBO = stride between
    # 8-row K groups (rs).  (The swapped assignment faults -- scripts/tc_probe.py.)
    op = dict(off=(c0 // 8) * cs, lbo=rs, sbo=cs, kstep=2 * rs, mn=1)
    got = run(qi, ki, n, 128 // 16, op, dict(op))
    assert relerr(got, ref) < 1e-5
    # head-1 window used by the front pass: rows = channels 8..135, cols = channels 64..143
    ref1 = q.float()[:, 8:136].T @ kk.float()[:, 64:144]
    got1 = run(qi, ki, 80, 128 // 16, dict(off=cs, lbo=rs, sbo=cs, kstep=2 * rs, mn=1),
               dict(off=8 * cs, lbo=rs, sbo=cs, kstep=2 * rs, mn=1))
    assert relerr(got1, ref1) < 1e-5


@pytest.mark.parametrize("prec,tol", [("fp16", 1e-3), ("bf16", 8e-3)])
@pytest.mark.parametrize("cin,cout,k,b,h,w,add,res", [
    (20, 20, 3, 2, 96, 72, True, False),     # RSB 3x3 of offset_mask_combine_conv, fused input add
    (13, 13, 3, 2, 24, 16, False, False),
    (6, 6, 3, 3, 16, 8, True, False),        # def_fuse branch convs; map smaller than one tile
    (51, 80, 1, 2, 96, 72, False, False),    # conv_bn_relu1 (1x1, wide)
    (80, 32, 1, 2, 24, 16, False, True),     # conv_bn_relu3 with the block residual, 10 staged items / thread
    (17, 24, 1, 1, 8, 8, False, False),
    (20, 20, 3, 1, 128, 96, False, False),   # BASELINE config 5 map size
])
def test_tensor_core_conv_vs_float64(cin, cout, k, b, h, w, add, res, prec, tol):
    """a7, 16-bit path: implicit-GEMM conv on tcgen05 (three pre-shifted 16-bit copies of the input
    rows, no im2col tile) + folded BatchNorm + ReLU + fused add / residual vs a float64 CPU conv."""
    import torch.nn.functional as F
    from otpose_b200.model.RSB import conv_bn_relu
    torch.manual_seed(cin * 100 + cout)
    m = conv_bn_relu(cin, cout, k, 1, k // 2).eval()
    with torch.no_grad():
        m.bn.running_mean.normal_()
        m.bn.running_var.uniform_(0.5, 2)
        m.bn.weight.uniform_(0.5, 1.5)
        m.bn.bias.normal_()
    x = torch.randn(b, cin, h, w)
    xa = torch.randn_like(x) if add else None
    r = torch.randn(b, cout, h, w) if res else None
    m = m.cuda()
    wt, bias = (t.cpu() for t in m.folded())          # otp_conv_bn_fold (library launch: CUDA tensors only)
    with torch.no_grad():                             # ... checked against the module's own parameters
        g = m.bn.weight.cpu().double() / torch.sqrt(m.bn.running_var.cpu().double() + m.bn.eps)
        w64 = m.conv.weight.cpu().double() * g.view(-1, 1, 1, 1)
        b64 = (m.conv.bias.cpu().double() - m.bn.running_mean.cpu().double()) * g + m.bn.bias.cpu().double()
    assert float((wt.double() - w64).abs().max()) < 1e-6 and float((bias.double() - b64).abs().max()) < 1e-6
    ref = F.conv2d((x + (xa if add else 0)).double(), w64, b64, padding=k // 2)
    ref = F.relu(ref + (r.double() if res else 0))
    assert _lib.load().otp_conv2d_tc_supported(cin, cout, h, w, k) == 1
    xg, xag, rg = x.cuda(), xa.cuda() if add else None, r.cuda() if res else None
    for p, t in (("fp32", 1e-5), (prec, tol)):
        m.precision = p
        y = torch.empty(b, cout, h, w, device="cuda")
        m.run(xg.data_ptr(), cin * h * w, y.data_ptr(), cout * h * w, b, h, w,
              x_add=xag.data_ptr() if add else None, x_add_bs=cin * h * w,
              residual=rg.data_ptr() if res else None, residual_bs=cout * h * w)
        err = float((y.cpu().double() - ref).abs().max() / ref.abs().max())
        assert err < t, (p, err)
    # shapes the tensor-core path does not take fall back to the CUDA-core kernel (w % 8 != 0)
    assert _lib.load().otp_conv2d_tc_supported(cin, cout, 9, 7, k) == 0
