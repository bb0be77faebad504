"""GPU diagnostic: tc conv time vs batch (slope = per-tile cost, intercept = fixed cost / CPU launch bound)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otpose_b200.model.RSB import conv_bn_relu  # noqa: E402

h, w = 96, 72
for cin, cout, k in ((6, 6, 3), (20, 20, 3), (51, 80, 1)):
    for prec in ("fp16", "fp32"):
        m = conv_bn_relu(cin, cout, k, 1, k // 2).cuda().eval()
        m.precision = prec
        res = []
        for b in (8, 32, 128, 256):
            x = torch.randn(b, cin, h, w, device="cuda")
            y = torch.empty(b, cout, h, w, device="cuda")
            run = lambda: m.run(x.data_ptr(), cin * h * w, y.data_ptr(), cout * h * w, b, h, w)   # noqa: E731
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            res.append((b, e0.elapsed_time(e1) / 20 * 1e3))
        print(f"conv {cin}->{cout} k{k} {prec}: " + "  ".join(f"B={b}: {t:7.1f} us" for b, t in res), flush=True)
