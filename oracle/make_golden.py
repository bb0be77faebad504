#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE's own code on CPU.

Runs only in the build container (needs /root/reference); the GPU box and the
test-suite only read the committed vectors.  What is executed unmodified:

* ``model/ConvVideoTransformer.py`` ConvTransformer, ``model/blocks.py``,
  ``model/RSB.py`` CHAIN_RSB_BLOCKS                     (imported as-is)
* ``model/OTPose.py`` OTPose.__init__/forward lines 307-394 (imported as-is)
* ``utils/heatmap.py`` get_max_preds / get_final_preds, ``utils/transform.py``
* ``model/HRNet.py`` HRNet (the W48 backbone of configs[1]; constructor config =
  configs/Base_PoseTrack17.yaml:46-88)

What has to be shimmed, and how (nothing below changes arithmetic on the path):

* ``matplotlib`` (absent here; imported by utils/heatmap.py:6 for plotting only)
  -> empty stub module.
* ``deform_conv_cuda`` / ``deform_pool_cuda`` (the reference CUDA extension does
  not compile against torch 2.11) -> stub modules, and the
  functional ``modulated_deform_conv`` the module calls
  (thirdparty/deform_conv/modules/deform_conv.py:128-131) is routed to
  ``torchvision.ops.deform_conv2d`` -- same algorithm and channel layout; the
  literal restatement of the reference kernel in oracle/head_oracle.py is
  checked equal to it.
* ``.cuda()`` calls inside OTPose.__init__/forward (no GPU here) -> identity.
* ``HRNet`` backbone (out of scope) -> stub returning the supplied
  ``rough_heatmaps``.

Usage:  python oracle/make_golden.py   (rewrites tests/golden/)
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OTPOSE_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from otpose_b200.utils import synthetic as syn  # noqa: E402  (synthetic weights/inputs only)

GOLD = os.path.join(REPO, "tests", "golden")


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def make_cfg(h, w, joints=17, dilations=(3, 6, 9, 12, 15)):
    return AttrDict(MODEL=AttrDict(
        EXTRA=AttrDict(FINAL_CONV_KERNEL=1, PRETRAINED_LAYERS=["*"]),
        HEATMAP_SIZE=[w, h], NUM_JOINTS=joints, FREEZE_HRNET_WEIGHTS=False, PRETRAINED="",
        DEFORMABLE_CONV=AttrDict(DILATION=list(dilations), AGGREGATION_TYPE="weighted_sum"),
        DEFORMABLE_CONV_CH=32, OFFSET_MASK_COMBINE_CONV=2))


def install_shims():
    for name in ("matplotlib", "matplotlib.pyplot", "thirdparty.deform_conv.deform_conv_cuda",
                 "thirdparty.deform_conv.deform_pool_cuda"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    torch.nn.Module.cuda = lambda self, device=None: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    # thirdparty/__init__ pulls thirdparty.utils -> nms_1d_cpu (a C++ ext that is
    # dead code for this model); pre-seed the package so only deform_conv loads.
    pkg = types.ModuleType("thirdparty")
    pkg.__path__ = [os.path.join(REF, "thirdparty")]
    sys.modules["thirdparty"] = pkg


def import_reference():
    install_shims()
    from torchvision.ops import deform_conv2d
    import thirdparty.deform_conv.modules.deform_conv as dcm

    def tv_mdcn(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
        assert groups == 1
        return deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding,
                             dilation=dilation, mask=mask)

    dcm.modulated_deform_conv = tv_mdcn
    import model.OTPose as ref_otpose
    from model.ConvVideoTransformer import ConvTransformer
    from model.RSB import CHAIN_RSB_BLOCKS
    import utils.heatmap as ref_heatmap

    class StubBackbone(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.rough = None

        def forward(self, x):
            return self.rough

    ref_otpose.HRNet = StubBackbone
    return ref_otpose, ConvTransformer, CHAIN_RSB_BLOCKS, ref_heatmap


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ref_otpose, ConvTransformer, CHAIN_RSB_BLOCKS, ref_heatmap = import_reference()
    manifest = {}

    # ---- encoders (ConvTransformer), C=136 temporal and C=17 flow ----------
    for name, c, nh, arch, (h, w), seed in (("encoder_c136", 136, 2, (0, 6, 2), (8, 6), 11),
                                            ("encoder_c17", 17, 1, (0, 6, 0), (8, 6), 12),
                                            ("encoder_c136_odd", 136, 2, (0, 2, 2), (7, 5), 13)):
        m = ConvTransformer(c, c, n_head=nh, n_embd_ks=3, max_len=h * w, arch=arch,
                            proj_pdrop=0.1, path_pdrop=0.1, h=h).eval()
        sh = shapes_of(m)
        manifest[name] = {k: list(v) for k, v in sh.items()}
        m.load_state_dict(syn.fill_state_dict(sh, seed=2024))
        x = torch.from_numpy(np.random.default_rng(seed).standard_normal((2, c, h, w)).astype(np.float32))
        outs = m(x)
        save(name, x=x, seed=2024, n_head=nh, arch=np.array(arch), **{f"out{i}": o for i, o in enumerate(outs)})

    # ---- RSB chains ---------------------------------------------------------
    for name, cin, cout, seed, hw in (("rsb_def_fuse", 17, 17, 21, (9, 7)), ("rsb_combine", 51, 32, 22, (9, 7)),
                                      ("rsb_combine_w8", 51, 32, 23, (10, 8))):
        m = CHAIN_RSB_BLOCKS(cin, cout, 2).eval()
        sh = shapes_of(m)
        manifest[name] = {k: list(v) for k, v in sh.items()}
        m.load_state_dict(syn.fill_state_dict(sh, seed=2024))
        x = torch.from_numpy(np.random.default_rng(seed).standard_normal((2, cin) + hw).astype(np.float32))
        save(name, x=x, seed=2024, out=m(x))

    # ---- the reference OTPose.forward itself (backbone stubbed) -------------
    # head_16x16: W % 8 == 0, so the 16-bit modes reach the tcgen05 RSB convs (conv_tc) on a golden fixture
    for name, b, h, w in (("head_16x12", 2, 16, 12), ("head_24x20", 1, 24, 20), ("head_16x16", 1, 16, 16)):
        model = ref_otpose.OTPose(make_cfg(h, w), phase="validate").eval()
        sh = shapes_of(model)
        if "head" not in manifest:
            manifest["head"] = {k: list(v) for k, v in sh.items() if "pos_embd" not in k}
        model.load_state_dict(syn.fill_state_dict(sh, seed=2024))
        rough = syn.synth_rough_heatmaps(b, 17, h, w, seed=1234)
        margin = syn.synth_margin(b, seed=1236)
        model.rough_pose_estimation_net.rough = rough
        outs = model(torch.zeros(b, 15, 4, 4), margin=margin)
        names = ("output_heatmaps", "rough_heatmaps", "intersection", "prev_b", "context_encoding",
                 "squeezed", "total_b")
        save(name, seed=2024, rough_seed=1234, margin_seed=1236, margin=margin,
             **{n: o for n, o in zip(names, outs) if n != "rough_heatmaps"})

    # ---- HRNet-W48 backbone (reference model/HRNet.py, imported as-is) -------
    from model.HRNet import HRNet as RefHRNet
    from otpose_b200.model.HRNet import hrnet_w48_cfg
    net = RefHRNet(hrnet_w48_cfg()).eval()
    sh = shapes_of(net)
    manifest["hrnet"] = {k: list(v) for k, v in sh.items()}
    net.load_state_dict(syn.fill_state_dict(sh, seed=2024))
    img = torch.from_numpy(np.random.default_rng(31).standard_normal((1, 3, 64, 64)).astype(np.float32))
    save("hrnet_w48_64x64", x=img, seed=2024, out=net(img))

    # ---- ST_OHKW_MSELoss (reference model/loss.py, imported as-is) -----------
    from model.loss import ST_OHKW_MSELoss as RefLoss
    r = np.random.default_rng(41)
    lb, lj, lh, lw = 3, 17, 12, 10
    out_s = torch.from_numpy(r.random((lb, lj, lh, lw)).astype(np.float32)).requires_grad_(True)
    out_t = torch.from_numpy(r.random((lb, lj, lh, lw)).astype(np.float32))
    gt = r.random((lb, lj, lh, lw)).astype(np.float32) * 0.9
    gt[:, ::3] = gt[:, ::3] / gt[:, ::3].max(axis=(0, 2, 3), keepdims=True)     # every third joint peaks at exactly 1
    gt = torch.from_numpy(gt)
    tw = torch.from_numpy((r.random((lb, lj, 1)) > 0.2).astype(np.float32))
    with torch.enable_grad():
        res = RefLoss(use_target_weight=True)(out_s, out_t, gt, tw)
        grad = torch.autograd.grad(res["final_loss"], out_s)[0]
    save("loss_st_ohkw", output_s=out_s.detach(), output_t=out_t, target=gt, target_weight=tw,
         ohkm_loss_s=res["ohkm_loss_s"].detach(), mse_loss_s=res["mse_loss_s"].detach(),
         final_loss=res["final_loss"].detach(), grad_output_s=grad)

    # ---- get_final_preds / get_max_preds (reference numpy code + cv2) --------
    hm = syn.synth_rough_heatmaps(3, 17, 24, 18, frames=1, seed=77).numpy()
    hm[0, 0] = 0.0                      # all-zero map: maxval <= 0 -> coords zeroed
    hm[0, 1] = -1.0
    hm[0, 2] = 0.0; hm[0, 2, 5, 0] = 1.0           # border maximum: no quarter offset
    hm[0, 3] = 0.0; hm[0, 3, 5, 1] = 1.0           # px == 1: no quarter offset
    hm[0, 4] = 0.0; hm[0, 4, 5, 16] = 1.0          # px == W-2 : offset allowed
    hm[0, 5] = 0.0; hm[0, 5, 5, 17] = 1.0          # px == W-1 : none
    hm[0, 6] = 0.0; hm[0, 6, 7, 7] = 1.0; hm[0, 6, 9, 9] = 1.0   # tie -> first index
    hm[0, 7] = 0.0; hm[0, 7, 10, 10] = 1.0; hm[0, 7, 10, 11] = 0.5; hm[0, 7, 11, 10] = 0.25
    hm[0, 8] = 0.0; hm[0, 8, 10, 10] = 1.0; hm[0, 8, 10, 9] = 0.5; hm[0, 8, 9, 10] = 0.25
    hm[0, 9] = 0.0; hm[0, 9, 22, 10] = 1.0         # py == H-2
    hm[0, 10] = 0.0; hm[0, 10, 23, 10] = 1.0       # py == H-1
    center, scale = syn.synth_center_scale(3, seed=1237)
    mp, mv = ref_heatmap.get_max_preds(hm.copy())
    fp, fv = ref_heatmap.get_final_preds(hm.copy(), center, scale)
    save("final_preds", heatmaps=hm, center=center, scale=scale, max_preds=mp, max_vals=mv,
         final_preds=fp, final_vals=fv)

    with open(os.path.join(GOLD, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    print("wrote manifest with", {k: len(v) for k, v in manifest.items()})


if __name__ == "__main__":
    main()
