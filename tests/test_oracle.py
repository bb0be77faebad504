"""CPU: the oracle restatement vs the golden vectors produced by running the
reference's own code (oracle/make_golden.py).  Tolerance: fp32 CPU vs fp32 CPU of
the same ATen ops, so 1e-5 relative-to-max (op order differences only)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from otpose_b200.utils import synthetic as syn


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def manifest(golden_dir):
    with open(os.path.join(golden_dir, "state_dict_manifest.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["encoder_c136", "encoder_c17", "encoder_c136_odd"])
def test_encoder_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    sd = syn.fill_state_dict(manifest(golden_dir)[name], seed=int(g["seed"]))
    outs = ho.conv_transformer(sd, "", torch.from_numpy(g["x"]), int(g["n_head"]), tuple(g["arch"]))
    assert len(outs) == 1 + int(g["arch"][2])
    for i, o in enumerate(outs):
        assert o.shape == g[f"out{i}"].shape
        assert relmax(o.numpy(), g[f"out{i}"]) < 1e-5


@pytest.mark.parametrize("name", ["rsb_def_fuse", "rsb_combine", "rsb_combine_w8"])
def test_rsb_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    sd = syn.fill_state_dict(manifest(golden_dir)[name], seed=int(g["seed"]))
    out = ho.chain_rsb(sd, "", torch.from_numpy(g["x"]), 2)
    assert relmax(out.numpy(), g["out"]) < 1e-5


def head_state_dict(golden_dir, h, w, seed):
    shapes = dict(manifest(golden_dir)["head"])
    for enc, c in (("temporal_encoder1", 136), ("temporal_encoder2", 136), ("flow_encoder", 17)):
        shapes[enc + ".pos_embd"] = (1, c, h * w)
    return syn.fill_state_dict(shapes, seed=seed)


@pytest.mark.parametrize("name,b,h,w", [("head_16x12", 2, 16, 12), ("head_24x20", 1, 24, 20), ("head_16x16", 1, 16, 16)])
def test_head_matches_reference_forward(golden_dir, name, b, h, w):
    g = load(golden_dir, name)
    sd = head_state_dict(golden_dir, h, w, int(g["seed"]))
    rough = syn.synth_rough_heatmaps(b, 17, h, w, seed=int(g["rough_seed"]))
    margin = syn.synth_margin(b, seed=int(g["margin_seed"]))
    assert np.array_equal(margin.numpy(), g["margin"])
    outs = ho.head_forward(sd, rough, margin)
    names = ("output_heatmaps", "rough_heatmaps", "intersection", "prev_b", "context_encoding",
             "squeezed", "total_b")
    for n, o in zip(names, outs):
        if n == "rough_heatmaps":
            assert o is rough
            continue
        assert o.shape == g[n].shape, n
        assert relmax(o.numpy(), g[n]) < 2e-5, n


def test_final_preds_matches_reference(golden_dir):
    g = load(golden_dir, "final_preds")
    mp, mv = ho.get_max_preds(g["heatmaps"].copy())
    assert np.array_equal(mp, g["max_preds"])
    assert np.array_equal(mv, g["max_vals"])
    fp, fv = ho.get_final_preds(g["heatmaps"].copy(), g["center"], g["scale"])
    assert np.array_equal(fv, g["final_vals"])
    assert fp.dtype == g["final_preds"].dtype
    np.testing.assert_allclose(fp, g["final_preds"], rtol=1e-6, atol=1e-3)
    idx, coords, preds, maxvals = ho.final_preds_full(g["heatmaps"].copy(), g["center"], g["scale"])
    np.testing.assert_array_equal(preds, fp)
    # first-index tie break (joint 6 of sample 0 has two equal maxima)
    assert idx[0, 6] == 7 * 18 + 7


def test_affine_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    center, scale = syn.synth_center_scale(5, seed=3)
    for c, s in zip(center, scale):
        m = ho.affine_transform_inv(c, s, [72, 96])
        # closed form documented in SURVEY section 3.4: isotropic scale s*200/W about the centre
        k = float(np.float32(s[0] * 200.0)) / 72
        np.testing.assert_allclose(m, [[k, 0, c[0] - 36 * k], [0, k, c[1] - 48 * k]], rtol=1e-5, atol=1e-3)


def test_mdcn_literal_equals_torchvision():
    r = np.random.default_rng(5)
    for d in (1, 3, 6):
        x = torch.from_numpy(r.standard_normal((2, 5, 11, 9)).astype(np.float32))
        off = torch.from_numpy((r.standard_normal((2, 5 * 18, 11, 9)) * 4).astype(np.float32))
        msk = torch.from_numpy(r.standard_normal((2, 5 * 9, 11, 9)).astype(np.float32))
        w = torch.from_numpy(r.standard_normal((5, 5, 3, 3)).astype(np.float32))
        bias = torch.from_numpy(r.standard_normal(5).astype(np.float32))
        a = ho.mdcn_forward_literal(x, off, msk, w, bias, 1, d, d, 5)
        b = ho.mdcn_forward(x, off, msk, w, bias, 1, d, d)
        assert relmax(a.numpy(), b.numpy()) < 1e-5


def test_mdcn_known_answers():
    """zero offsets + unit masks + identity centre tap -> output == input + bias;
    integer offsets -> exact shifted copy with zero fill."""
    r = np.random.default_rng(6)
    x = torch.from_numpy(r.standard_normal((1, 4, 8, 7)).astype(np.float32))
    w = torch.zeros(4, 4, 3, 3)
    for k in range(4):
        w[k, k, 1, 1] = 1.0
    bias = torch.tensor([0.5, -1.0, 0.0, 2.0])
    off = torch.zeros(1, 4 * 18, 8, 7)
    msk = torch.ones(1, 4 * 9, 8, 7)
    out = ho.mdcn_forward_literal(x, off, msk, w, bias, 1, 2, 2, 4)
    assert torch.equal(out, x + bias.view(1, 4, 1, 1))
    off[:, 0::2] = 1.0   # dh = +1
    off[:, 1::2] = -2.0  # dw = -2
    out = ho.mdcn_forward_literal(x, off, msk, w, None, 1, 2, 2, 4)
    exp = torch.zeros_like(x)
    exp[:, :, :-1, 2:] = x[:, :, 1:, :-2]
    assert torch.equal(out, exp)


def test_scramble_formula():
    """out2[h*hs + r, s] = o[h, f % hs, f // hs], f = r*T + s."""
    nh, hs, t = 2, 3, 4
    o = torch.arange(nh * hs * t, dtype=torch.float32).view(1, nh, hs, t)
    out2 = o.transpose(2, 3).contiguous().view(1, nh * hs, -1)
    for h in range(nh):
        for r_ in range(hs):
            for s in range(t):
                f = r_ * t + s
                assert out2[0, h * hs + r_, s] == o[0, h, f % hs, f // hs]


def test_frame_window_prologue_reduces_to_the_reference_at_5_frames():
    """BASELINE config 5: the generalised (3/5/7-frame) prologue is a definition, pinned by
    requiring its 5-frame instance to equal the line-for-line restatement of
    model/OTPose.py:320-354 bit for bit."""
    rough = syn.synth_rough_heatmaps(3, 17, 12, 8, seed=4)
    margin = syn.synth_margin(3, seed=8)
    a, b = ho.fusion_prologue(rough, margin), ho.fusion_prologue_frames(rough, margin)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # 3 frames: no far pair, so far_b is the current frame alone
    r3 = syn.synth_rough_heatmaps(2, 17, 12, 8, frames=3, seed=5)
    m3 = syn.synth_margin(2, seed=9, frames=3)
    f3 = ho.fusion_prologue_frames(r3, m3)
    cur, prev, nxt = r3.split(2, dim=0)
    assert torch.equal(f3["far_b"], cur)
    assert torch.equal(f3["total_b"], cur + prev + nxt)
    assert torch.equal(f3["prev_b"], cur + prev / (m3[:, 0] + 1)[:, None, None, None])
    # 7 frames: every supplementary frame contributes to total_b; far_b holds pairs 2 and 3
    r7 = syn.synth_rough_heatmaps(2, 17, 12, 8, frames=7, seed=6)
    m7 = torch.zeros(2, 6, dtype=torch.int64)
    f7 = ho.fusion_prologue_frames(r7, m7)
    fr = r7.split(2, dim=0)
    assert relmax(f7["total_b"], sum(fr)) < 1e-6
    assert relmax(f7["far_b"], fr[0] + fr[3] + fr[4] + fr[5] + fr[6]) < 1e-6
    assert relmax(f7["prev_b"], fr[0] + fr[1] + fr[3] + fr[5]) < 1e-6
