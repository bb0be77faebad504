"""RSB chains at bench size (32 clips, 96x72), fused 16-bit kernels: timing per chain (and a target for ncu)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from otpose_b200.utils import synthetic as syn
from otpose_b200.model.RSB import CHAIN_RSB_BLOCKS
from otpose_b200 import _lib
b, h, w = 32, 96, 72
lib = _lib.load()
for cin, cout in ((17, 17), (51, 32)):
    m = CHAIN_RSB_BLOCKS(cin, cout, 2)
    m.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed=5))
    m = m.cuda().eval()
    for mod in m.modules():
        if hasattr(mod, "precision"): mod.precision = "fp16"
    x = torch.randn(b, cin, h, w, device="cuda")
    for fused in (True, False):
        for blk in m.layers: blk.fused = fused
        for _ in range(3): y = m(x)
        torch.cuda.synchronize()
        n0 = lib.otp_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = m(x)
        nl = lib.otp_launch_count() - n0
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("chain %d->%d fused=%s: %.1f us per chain (CUDA graph replay), %d launches" % (cin, cout, fused, e0.elapsed_time(e1) * 100, nl), flush=True)
