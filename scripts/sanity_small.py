"""Small end-to-end run of every kernel family (for compute-sanitizer memcheck / racecheck / synccheck)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otpose_b200.model import OTPose, default_cfg  # noqa: E402
from otpose_b200.thirdparty.deform_conv import ModulatedDeformConv  # noqa: E402
from otpose_b200.utils import heatmap, synthetic as syn  # noqa: E402

b, h, w = 2, 24, 16
for prec in ("fp16", "fp32"):
    model = OTPose(default_cfg((h, w)), precision=prec)
    model.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=2024))
    model = model.cuda().eval()
    for frames in (5, 3):
        rough = syn.synth_rough_heatmaps(b, 17, h, w, frames=frames).cuda()
        margin = syn.synth_margin(b, frames=frames).cuda()
        out = model.forward_head(rough, margin)[0]
        r = heatmap.final_preds_cuda(out)
        torch.cuda.synchronize()
        print(prec, frames, float(out.abs().max()), int(r["idx"].sum()))
    # a0 + a1 fused hand-off kernel on a bf16 channels-last feature map (ragged tile: 24 * 16 = 384 = 1.5 CTA tiles)
    feats = torch.randn(5 * b, 48, h, w).to(torch.bfloat16).cuda().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(17, 48, 1, 1) / 7).cuda()
    out = model.forward_from_features(feats, syn.synth_margin(b).cuda(), wt, torch.zeros(17).cuda())[0]
    torch.cuda.synchronize()
    print(prec, "from features", float(out.abs().max()))
m = ModulatedDeformConv(17, 17, 3, padding=3, dilation=3, deformable_groups=17).cuda()
x = torch.randn(2, 17, h, w, device="cuda", requires_grad=True)
off = (torch.randn(2, 306, h, w, device="cuda") * 2).requires_grad_(True)
msk = torch.randn(2, 153, h, w, device="cuda", requires_grad=True)
m(x, off, msk).square().mean().backward()
torch.cuda.synchronize()
print("dcn bwd", float(x.grad.abs().max()), float(m.weight.grad.abs().max()))
