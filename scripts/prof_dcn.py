"""One head forward at bench size (32 clips, 96x72, fp16 operands, eager) -- ncu target for the fused offset/mask/DCN kernel."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from otpose_b200.utils import synthetic as syn
from test_gpu_parity import build_head
b, h, w = 32, 96, 72
model, sd = build_head(h, w, "fp16")
rough = syn.synth_rough_heatmaps(b, 17, h, w).cuda()
margin = syn.synth_margin(b).cuda()
for _ in range(2):
    outs = model.forward_head(rough, margin)
torch.cuda.synchronize()
print("ok")
