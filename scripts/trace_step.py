"""GPU diagnostic: per-launch device durations of one head forward in normal (un-serialised) execution,
via torch.profiler (CUPTI).  usage: python scripts/trace_step.py [batch] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otpose_b200.model import OTPose, default_cfg  # noqa: E402
from otpose_b200.utils import heatmap, synthetic as syn  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
m = OTPose(default_cfg((96, 72)), precision=prec)
m.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in m.state_dict().items()}))
m = m.cuda().eval()
rough, margin = syn.synth_rough_heatmaps(b, 17, 96, 72).cuda(), syn.synth_margin(b).cuda()
for _ in range(3):
    m.forward_head(rough, margin)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    out = m.forward_head(rough, margin)[0]
    heatmap.final_preds_cuda(out)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
tot = 0.0
for e in evs:
    name = e.name.split("(")[0].split("::")[-1][:40]
    print(f"{(e.time_range.start - t0):10.1f} us  +{e.device_time:8.1f} us  {name}")
    tot += e.device_time
print("sum of kernel time", tot, "us; span", evs[-1].time_range.end - t0, "us")
