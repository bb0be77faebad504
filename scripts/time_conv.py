"""GPU diagnostic: time RSB-shaped convs (tensor-core vs CUDA-core path) back to back."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from otpose_b200.model.RSB import conv_bn_relu  # noqa: E402

torch.manual_seed(0)
b, h, w = 32, 96, 72


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for cin, cout, k in ((20, 20, 3), (6, 6, 3), (13, 13, 3), (51, 80, 1), (80, 32, 1), (17, 24, 1)):
    x = torch.randn(b, cin, h, w, device="cuda")
    for prec in ("fp32", "fp16"):
        m = conv_bn_relu(cin, cout, k, 1, k // 2).cuda().eval()
        m.precision = prec
        y = torch.empty(b, cout, h, w, device="cuda")
        run = lambda: m.run(x.data_ptr(), cin * h * w, y.data_ptr(), cout * h * w, b, h, w)   # noqa: E731
        print(f"conv {cin:3d}->{cout:3d} k{k} {prec}: {timeit(run):8.1f} us / launch", flush=True)
