"""Flow-encoder check at head sizes: fused 16-bit launch vs the fp32 per-block kernels, per stem depth.
usage: check_flow.py [h w frames]"""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from otpose_b200.utils import synthetic as syn
from test_gpu_parity import build_head, rel
import torch.nn as nn
h, w, frames = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (96, 72, 5)
b = 1
model, sd = build_head(h, w, "fp16")
rough = syn.synth_rough_heatmaps(b, 17, h, w, frames=frames, seed=77).cuda()
margin = syn.synth_margin(b, seed=78, frames=frames).cuda()
outs = model.forward_head(rough, margin)
model.flow_encoder.fused_stem = False
outs32 = model.forward_head(rough, margin)
model.flow_encoder.fused_stem = True
for i, n in enumerate(("output_heatmaps", "rough", "intersection", "prev_b", "context_encoding")):
    print(n, "fused-flow vs per-block-fp32-flow: %.3e" % rel(outs[i], outs32[i]))
tb = outs[6]
enc = model.flow_encoder
full = list(enc.stem)
for n in range(1, 7):
    enc.stem = nn.ModuleList(full[:n])
    enc.precision = "fp16"; got = enc(tb)[0]
    enc.precision = "fp32"; ref = enc(tb)[0]
    d = (got - ref).abs()
    print(n, "blocks: rel err %.3e" % rel(got, ref), "max|ref| %.3f" % float(ref.abs().max()),
          "rms err %.3e" % float(d.pow(2).mean().sqrt()), "argmax token", int(d.max(1)[0].argmax()), flush=True)
