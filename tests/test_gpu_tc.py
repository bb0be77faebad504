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
