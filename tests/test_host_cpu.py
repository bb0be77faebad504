"""CPU-only checks of the host side: the C-ABI library loads and exports every
symbol the header declares, drop-in modules expose the reference's state-dict
layout, argument validation / error behaviour, synthetic-data determinism."""
import json
import os
import re

import numpy as np
import pytest
import torch

from otpose_b200 import _lib
from otpose_b200.model import ConvTransformer, OTPose, default_cfg
from otpose_b200.model.RSB import CHAIN_RSB_BLOCKS
from otpose_b200.utils import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "otpose_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(otp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/otpose_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signatures out of sync with the header"
    assert b"sm_100a" in lib.otp_version()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libotpose_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_shape_queries_and_unsupported_widths():
    lib = _lib.load()
    assert lib.otp_block_packed_bytes(136, 2) > 4 * (3 * 136 * 136 + 8 * 136 * 136)
    assert lib.otp_block_packed_bytes(17, 1) > 0
    assert lib.otp_block_packed_bytes(100, 4) == 0
    assert b"not built" in lib.otp_last_error()
    assert lib.otp_block_workspace_bytes(2, 136, 48, 2, 1, 0) > 2 * 136 * 48 * 4
    assert lib.otp_block_workspace_bytes(2, 136, 48, 2, 3, 0) == 0
    # argument validation happens before any CUDA call
    assert lib.otp_final_preds(None, 1, 17, 0, 4, None, None, None, None, None, None, None) == 1
    assert lib.otp_mdcn_forward(None, None, None, None, None, None, 1, 17, 8, 8, 17, 3, 3, 1, 1, 1, 2, 17,
                                1.0, 0, None) == 2
    assert lib.otp_conv2d(None, 0, None, 0, None, None, None, 0, None, 0, 1, 4, 8, 8, 4, 5, 1, 0, None) == 2
    # empty batches are a no-op, not an error (reference handles B=0 tensors)
    assert lib.otp_final_preds(None, 0, 17, 8, 8, None, None, None, None, None, None, None) == 0
    assert lib.otp_fusion_sum(None, 0, 17, 64, None, None, None) == 0
    # a0 + a1 hand-off kernel: widths it is not built for are refused (status 2), bad frame windows are
    # argument errors, an empty batch is a no-op
    assert lib.otp_final_layer_fusion_sum(None, 1, 1, None, None, 5, 2, 44, 17, 64, None, None, None, None) == 2
    assert lib.otp_final_layer_fusion_sum(None, 1, 1, None, None, 5, 2, 48, 16, 64, None, None, None, None) == 2
    assert lib.otp_final_layer_fusion_sum(None, 1, 1, None, None, 4, 2, 48, 17, 64, None, None, None, None) == 1
    assert lib.otp_final_layer_fusion_sum(None, 1, 1, None, None, 5, 0, 48, 17, 64, None, None, None, None) == 0


def test_flow_encoder_and_rsb_block_queries():
    """Round-2 entry points (one-launch flow encoder, RSB level kernels): what is built / refused, workspace
    sizes, argument handling before any CUDA call."""
    lib = _lib.load()
    # flow encoder: C = 17 / one head, <= 8 blocks, up to 8 x 1728 tokens per clip
    assert lib.otp_flow_encoder_supported(17, 1, 96 * 72, 6) == 1
    assert lib.otp_flow_encoder_supported(17, 1, 128 * 96, 6) == 1
    assert lib.otp_flow_encoder_supported(17, 1, 35, 1) == 1
    assert lib.otp_flow_encoder_supported(136, 2, 96 * 72, 6) == 0
    assert lib.otp_flow_encoder_supported(17, 1, 8 * 1728 + 1, 6) == 0
    assert lib.otp_flow_encoder_supported(17, 1, 96 * 72, 9) == 0
    assert lib.otp_flow_encoder_workspace_bytes(32, 6912) >= 32 * 6912 * 17 * 4      # fp32 scramble buffer
    assert lib.otp_flow_encoder_workspace_bytes(0, 6912) == 0
    assert lib.otp_flow_encoder_forward(None, 6, None, None, 0, None, 1, 6912, None, 0, None) == 1   # no block table
    # RSB blocks: the three widths of the head (and the 128 x 96 maps of BASELINE config 5); maps wider than 96 are
    # refused (the caller falls back to the per-conv kernels)
    for cin, planes in ((17, 17), (51, 32), (32, 32)):
        assert lib.otp_rsb_block_supported(cin, planes, 96, 72) == 1
        assert lib.otp_rsb_block_supported(cin, planes, 128, 96) == 1
        assert lib.otp_rsb_block_supported(cin, planes, 9, 7) == 1
        assert lib.otp_rsb_block_pack_bytes(cin, planes, 1) > lib.otp_rsb_block_pack_bytes(cin, planes, 0) > 0
    assert lib.otp_rsb_block_supported(51, 32, 96, 104) == 0
    assert lib.otp_rsb_block_supported(128, 32, 96, 72) == 0
    # six 16-bit planes of cp channels + the 16-bit copy of x
    assert lib.otp_rsb_block_workspace_bytes(32, 51, 32, 96, 72) >= 32 * 96 * 72 * 2 * (6 * 24 + 56)
    assert lib.otp_rsb_block_forward(None, None, 0, None, 0, 1, 128, 32, 1, 96, 72, None, 0, None) == 2   # unsupported
    assert lib.otp_rsb_block_forward(None, None, 0, None, 0, 0, 51, 32, 1, 96, 72, None, 0, None) == 0    # empty batch
    assert lib.otp_rsb_block_forward(None, None, 0, None, 0, 1, 51, 32, 1, 96, 72, None, 0, None) == 1    # null pointers


def manifest():
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")) as f:
        return json.load(f)


def shapes(m):
    return {k: list(v.shape) for k, v in m.state_dict().items()}


def test_state_dict_layout_matches_reference():
    man = manifest()
    enc = ConvTransformer(136, 136, n_head=2, n_embd_ks=3, max_len=48, arch=(0, 6, 2), proj_pdrop=0.1,
                          path_pdrop=0.1, h=8)
    assert shapes(enc) == man["encoder_c136"]
    flow = ConvTransformer(17, 17, 1, 3, 48, arch=(0, 6, 0), proj_pdrop=0.1, path_pdrop=0.1, h=8)
    assert shapes(flow) == man["encoder_c17"]
    assert shapes(CHAIN_RSB_BLOCKS(17, 17, 2)) == man["rsb_def_fuse"]
    assert shapes(CHAIN_RSB_BLOCKS(51, 32, 2)) == man["rsb_combine"]
    head = OTPose(default_cfg((16, 12)))
    mine = {k: v for k, v in shapes(head).items() if "pos_embd" not in k}
    assert mine == man["head"]
    assert sum(p.numel() for p in OTPose(default_cfg((96, 72))).parameters()) == 4401878
    # reference init: AffineDropPath scale 1e-4, DCN identity centre tap
    assert torch.allclose(head.temporal_encoder1.stem[0].drop_path_attn.scale, torch.full((1, 136, 1), 1e-4))
    w = head.modulated_deform_conv_list[0].deform_conv.weight
    assert w[3, 3, 1, 1] == 1 and w.sum() == 17


def test_pos_embd_matches_reference_table():
    from oracle.head_oracle import sinusoid_encoding
    enc = ConvTransformer(136, 136, 2, 3, 64, arch=(0, 1, 0), h=8)
    assert torch.equal(enc.pos_embd, sinusoid_encoding(64, 136) / (136 ** 0.5))


def test_cpu_tensors_are_rejected_like_the_reference():
    from otpose_b200.thirdparty.deform_conv import ModulatedDeformConv
    m = ModulatedDeformConv(17, 17, 3, padding=3, dilation=3, deformable_groups=17)
    with pytest.raises(NotImplementedError):   # functions/deform_conv.py:131-132
        m(torch.zeros(1, 17, 8, 8), torch.zeros(1, 306, 8, 8), torch.zeros(1, 153, 8, 8))
    blk = ConvTransformer(17, 17, 1, 3, 64, arch=(0, 1, 0), h=8).eval()
    with pytest.raises(NotImplementedError):
        blk(torch.zeros(1, 17, 8, 8))
    with pytest.raises(NotImplementedError):   # the training path keeps the contract: CUDA tensors only
        ConvTransformer(17, 17, 1, 3, 64, arch=(0, 1, 0), h=8).train()(torch.zeros(1, 17, 8, 8))
    head = OTPose(default_cfg((8, 8))).eval()
    with pytest.raises(NotImplementedError):   # the feature hand-off boundary keeps the same contract
        head.forward_from_features(torch.zeros(5, 48, 8, 8), torch.zeros(1, 4, dtype=torch.int64),
                                   torch.zeros(17, 48, 1, 1), torch.zeros(17))


def test_synthetic_is_deterministic_and_nontrivial():
    a = syn.synth_tensor("temporal_encoder1.stem.0.drop_path_attn.scale", (1, 136, 1))
    b = syn.synth_tensor("temporal_encoder1.stem.0.drop_path_attn.scale", (1, 136, 1))
    assert torch.equal(a, b) and a.min() >= 0.5 and a.max() <= 1.5
    r = syn.synth_rough_heatmaps(2, 17, 24, 18)
    assert r.shape == (10, 17, 24, 18) and torch.equal(r, syn.synth_rough_heatmaps(2, 17, 24, 18))
    assert 0.2 < r.amax() < 1.2
    m = syn.synth_margin(64)
    assert m.dtype == torch.int64 and set(np.unique(m.numpy())) <= {0, 1, 2}


def test_hrnet_drop_in_matches_reference_golden():
    """SURVEY 8f rank 3 / configs[1]: the torch HRNet-W48 drop-in reproduces the reference's own module
    (fixture generated by running model/HRNet.py, oracle/make_golden.py) -- same state-dict layout, fp32 eval
    output equal to rounding, and the Conv+BN-folded inference copy within 1e-5."""
    import numpy as np
    from otpose_b200.model.HRNet import HRNet, hrnet_w48_cfg
    g = np.load(os.path.join(ROOT, "tests", "golden", "hrnet_w48_64x64.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")) as f:
        shapes = json.load(f)["hrnet"]
    net = HRNet(hrnet_w48_cfg()).eval()
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert mine == shapes
    net.load_state_dict(syn.fill_state_dict(shapes, seed=int(g["seed"])))
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        out = net(x)
        feats = net.features(x)
        folded = net.fold(dtype=torch.float32, memory_format=torch.channels_last)(x)
    ref = torch.from_numpy(g["out"])
    scale = ref.abs().max()
    assert out.shape == ref.shape and feats.shape == (1, 48, 16, 16)
    assert (out - ref).abs().max() / scale < 1e-6
    assert (folded - ref).abs().max() / scale < 1e-5


def test_device_side_loss_matches_reference_golden():
    """ST_OHKW_MSELoss without the per-joint host synchronisation equals the reference's
    loop (fixture from the reference's own model/loss.py: values and the gradient wrt the student maps)."""
    import numpy as np
    from otpose_b200.model.loss import ST_OHKW_MSELoss
    g = np.load(os.path.join(ROOT, "tests", "golden", "loss_st_ohkw.npz"))
    s = torch.from_numpy(g["output_s"]).requires_grad_(True)
    res = ST_OHKW_MSELoss(use_target_weight=True)(s, torch.from_numpy(g["output_t"]), torch.from_numpy(g["target"]),
                                                  torch.from_numpy(g["target_weight"]))
    for k in ("ohkm_loss_s", "mse_loss_s", "final_loss"):
        assert abs(float(res[k]) - float(g[k])) <= 1e-6 * max(1.0, abs(float(g[k]))), k
    grad = torch.autograd.grad(res["final_loss"], s)[0]
    assert (grad - torch.from_numpy(g["grad_output_s"])).abs().max() < 1e-7
    with pytest.raises(NotImplementedError):
        ST_OHKW_MSELoss(use_target_weight=False)
