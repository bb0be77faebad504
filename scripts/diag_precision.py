"""GPU diagnostic: per-stage error of the head vs the CPU oracle for each precision mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import head_oracle as ho  # noqa: E402
from otpose_b200.model import OTPose, default_cfg  # noqa: E402
from otpose_b200.utils import synthetic as syn  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


h, w, b = 96, 72, 1
shapes = {k: v.shape for k, v in OTPose(default_cfg((h, w))).state_dict().items()}
sd = syn.fill_state_dict(shapes, seed=2024)
rough, margin = syn.synth_rough_heatmaps(b, 17, h, w), syn.synth_margin(b)
ref, inter = ho.head_forward(sd, rough, margin, return_intermediates=True)
for prec in ("fp32", "bf16", "fp16"):
    m = OTPose(default_cfg((h, w)), precision=prec)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    dbg = {}
    outs = m.forward_head(rough.cuda(), margin.cuda(), _debug=dbg)
    print(prec, "ctx", rel(outs[4], ref[4]), "branches", rel(dbg["cat"][:, :34], inter["branches"]),
          "trans", rel(dbg["trans"], inter["trans"]), "out", rel(outs[0], ref[0]), flush=True)
    e = m.temporal_encoder1.forward_tokens(torch.as_tensor(inter["x1"].flatten(2) +
                                                           sd["temporal_encoder1.pos_embd"][:, :, :h * w]).cuda())
    print("   enc1 s0", rel(e[0], inter["e1"][0]))
