"""GPU experiment: forward_head + final_preds captured in a CUDA graph vs eager launches."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otpose_b200.model import OTPose, default_cfg  # noqa: E402
from otpose_b200.utils import heatmap, synthetic as syn  # noqa: E402

b, H, W = 32, 96, 72
model = OTPose(default_cfg((H, W)), precision="fp16")
model.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=2024))
model = model.cuda().eval()
rough = syn.synth_rough_heatmaps(b, 17, H, W).cuda()
margin = syn.synth_margin(b).cuda()
center, scale = (torch.from_numpy(a).cuda() for a in syn.synth_center_scale(b))


def step():
    out = model.forward_head(rough, margin)[0]
    return out, heatmap.final_preds_cuda(out, center, scale)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager ms/step", timeit(step))
ref_out, ref = step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    gout, gres = step()
print("graph ms/step", timeit(g.replay))
torch.cuda.synchronize()
print("same output:", torch.equal(gout, ref_out), torch.equal(gres["idx"], ref["idx"]))
