"""The drop-in claim exercised end to end:

* the pybind11 module ``deform_conv_cuda`` (otpose_b200/shim/deform_conv_cuda.cpp) exports the entry points
  of the reference's extension (thirdparty/deform_conv/src/deform_conv_cuda.cpp:666-680) and the reference's
  OWN ``ModulatedDeformConvFunction`` (functions/deform_conv.py:109-180) runs on it, forward and backward;
* the reference's OWN ``model/OTPose.py`` -- ``__init__`` and ``forward`` untouched, only the four import
  lines of OTPose.py:13-16 swapped as INTEGRATION.md section 2 shows -- reproduces the golden vectors that the
  unmodified reference produced (tests/golden/head_*.npz).

The reference's Python files are read from ``$OTPOSE_REFERENCE``, ``/root/reference`` or the offline
install ``baseline/_ref`` (scripts/install_reference.py; git-ignored, travels to the GPU box).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from otpose_b200.utils import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "otpose_b200", "lib", "deform_conv_cuda.so")


def reference_root():
    for p in (os.environ.get("OTPOSE_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.exists(os.path.join(p, "model", "OTPose.py")):
            return p
    return None


def load_shim():
    if not os.path.exists(SHIM):
        from otpose_b200 import build
        build.build_shim()
    spec = importlib.util.spec_from_file_location("deform_conv_cuda", SHIM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shim_exports_the_reference_entry_points():
    """CPU: the module loads (against libotpose_b200.so) and carries the five names the reference's
    functions/deform_conv.py looks up."""
    mod = load_shim()
    for name in ("modulated_deform_conv_cuda_forward", "modulated_deform_conv_cuda_backward",
                 "deform_conv_forward_cuda", "deform_conv_backward_input_cuda",
                 "deform_conv_backward_parameters_cuda"):
        assert callable(getattr(mod, name)), name
    with pytest.raises(RuntimeError):      # TORCH_CHECK -> RuntimeError, like the reference's checks
        z = torch.zeros(1, 17, 4, 4)
        mod.modulated_deform_conv_cuda_forward(z, torch.zeros(17, 17, 3, 3), z, z, z, z, z, z, 3, 3, 1, 1, 1, 1, 1, 1,
                                               1, 17, False)


def import_reference_dcn(ref, shim):
    """thirdparty.deform_conv of the reference with `deform_conv_cuda` = the shim (the reference imports it
    as `from .. import deform_conv_cuda`, functions/deform_conv.py:9)."""
    for k in [k for k in sys.modules if k == "thirdparty" or k.startswith("thirdparty.")]:
        del sys.modules[k]
    pkg = types.ModuleType("thirdparty")           # skip thirdparty/__init__ (pulls an unrelated NMS extension)
    pkg.__path__ = [os.path.join(ref, "thirdparty")]
    sys.modules["thirdparty"] = pkg
    sys.modules["thirdparty.deform_conv.deform_conv_cuda"] = shim
    sys.modules["thirdparty.deform_conv.deform_pool_cuda"] = types.ModuleType("deform_pool_cuda")
    import thirdparty.deform_conv as dc
    return dc


@pytest.mark.gpu
@pytest.mark.parametrize("dil", [1, 3])
def test_reference_function_runs_on_the_shim(dil):
    ref = reference_root()
    if ref is None:
        pytest.skip("reference tree not available (run scripts/install_reference.py where /root/reference exists)")
    from torchvision.ops import deform_conv2d
    dc = import_reference_dcn(ref, load_shim())
    g = torch.Generator().manual_seed(5)
    b, c, h, w, k = 2, 17, 12, 10, 3
    x = torch.randn(b, c, h, w, generator=g)
    off = torch.randn(b, 2 * k * k * c, h, w, generator=g) * 2
    msk = torch.randn(b, k * k * c, h, w, generator=g)
    wt = torch.randn(c, c, k, k, generator=g) * 0.2
    bias = torch.randn(c, generator=g)
    go = torch.randn(b, c, h, w, generator=g)
    leaves = [t.cuda().requires_grad_(True) for t in (x, off, msk, wt, bias)]
    out = dc.modulated_deform_conv(*leaves, 1, dil, dil, 1, c)     # the reference's Function.apply
    out.backward(go.cuda())
    ref_leaves = [t.double().requires_grad_(True) for t in (x, off, msk, wt, bias)]
    ref_out = deform_conv2d(ref_leaves[0], ref_leaves[1], ref_leaves[3], ref_leaves[4], stride=1, padding=dil,
                            dilation=dil, mask=ref_leaves[2])
    ref_out.backward(go.double())
    assert (out.detach().cpu().double() - ref_out.detach()).abs().max() / ref_out.abs().max() < 1e-5
    for a, r, name in zip(leaves, ref_leaves, ("input", "offset", "mask", "weight", "bias")):
        err = (a.grad.cpu().double() - r.grad).abs().max() / r.grad.abs().max()
        assert err < 5e-5, (name, float(err))
    # the reference module class on top of it
    m = dc.ModulatedDeformConv(c, c, 3, stride=1, padding=dil, dilation=dil, deformable_groups=c).cuda()
    y = m(leaves[0].detach(), leaves[1].detach(), leaves[2].detach())
    assert y.shape == (b, c, h, w) and torch.isfinite(y).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,b,h,w", [("head_16x12", 2, 16, 12), ("head_16x16", 1, 16, 16)])
def test_reference_otpose_forward_with_the_four_imports_swapped(name, b, h, w):
    """model/OTPose.py of the reference, executed with INTEGRATION.md's import patch and nothing else."""
    ref = reference_root()
    if ref is None:
        pytest.skip("reference tree not available (run scripts/install_reference.py where /root/reference exists)")
    src = open(os.path.join(ref, "model", "OTPose.py")).read()
    patch = {
        "from model.ConvVideoTransformer import ConvTransformer":
            "from otpose_b200.model.ConvVideoTransformer import ConvTransformer",
        "from model.RSB import CHAIN_RSB_BLOCKS": "from otpose_b200.model.RSB import CHAIN_RSB_BLOCKS",
        "from model.layers import DeformableCONV": "from otpose_b200.model.layers import DeformableCONV",
        "from thirdparty.deform_conv import DeformConv, ModulatedDeformConv":
            "from otpose_b200.thirdparty.deform_conv import DeformConv, ModulatedDeformConv",
    }
    for old, new in patch.items():
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    # environment the file needs besides the hot path: matplotlib (plot helpers of utils/heatmap.py), the
    # reference's own `model` / `utils` packages for HRNet and normalize_0_to_1
    for modname in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(modname, types.ModuleType(modname))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("model", "utils") or
             k.startswith(("model.", "utils."))}
    sys.path.insert(0, ref)
    try:
        mod = types.ModuleType("reference_otpose_patched")
        mod.__file__ = os.path.join(ref, "model", "OTPose.py")
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)

        class StubBackbone(torch.nn.Module):      # the HRNet backbone is out of the path: hands over the heat maps
            def __init__(self, *a, **k):
                super().__init__()
                self.rough = None

            def forward(self, x):
                return self.rough

        mod.HRNet = StubBackbone
        from oracle.make_golden import make_cfg
        model = mod.OTPose(make_cfg(h, w), phase="validate").cuda().eval()
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k in ("model", "utils") or k.startswith(("model.", "utils."))]:
            del sys.modules[k]
        sys.modules.update(saved)
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sd = syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=int(g["seed"]))
    model.load_state_dict(sd)
    rough = syn.synth_rough_heatmaps(b, 17, h, w, seed=int(g["rough_seed"])).cuda()
    margin = syn.synth_margin(b, seed=int(g["margin_seed"])).cuda()
    model.rough_pose_estimation_net.rough = rough
    # the reference's own nn.Conv2d layers (final_layer1/2, offset / mask convs) run through cuDNN, which
    # defaults to TF32 (1e-3-level rounding, amplified by the learned-offset sampling): hold them to fp32 so
    # that the comparison is about the drop-in modules
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            outs = model(torch.zeros(b, 15, 4, 4, device="cuda"), margin=margin)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    names = ("output_heatmaps", "rough_heatmaps", "intersection", "prev_b", "context_encoding", "squeezed", "total_b")
    for n, o in zip(names, outs):
        if n == "rough_heatmaps":
            continue
        ref_v = torch.from_numpy(g[n])
        err = (o.cpu() - ref_v).abs().max() / ref_v.abs().max()
        assert err < 1e-3, (n, float(err))
