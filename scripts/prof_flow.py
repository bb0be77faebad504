"""One flow-encoder launch at bench size (32 clips, 96x72) for ncu / timing."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from otpose_b200.utils import synthetic as syn
from otpose_b200.model.ConvVideoTransformer import ConvTransformer
b, h, w = 32, 96, 72
m = ConvTransformer(17, 17, n_head=1, n_embd_ks=3, max_len=h * w, arch=(0, 6, 0), proj_pdrop=0.1, path_pdrop=0.1, h=h, precision="fp16")
m.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed=11))
m = m.cuda().eval()
x = torch.randn(b, 17, h, w, device="cuda")
for _ in range(3): y = m(x)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): y = m(x)[0]
e1.record(); torch.cuda.synchronize()
print("flow encoder: %.1f us per launch" % (e0.elapsed_time(e1) * 1000 / 20))
