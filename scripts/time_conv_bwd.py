"""Time the native backward of the offset conv (32 -> 306, 3x3, dilation 3) at bench size.

    python scripts/time_conv_bwd.py [clips]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otpose_b200.model.conv2d_fn import conv2d  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cin, cout, h, w, d = 32, 306, 96, 72, 3
x = torch.randn(b, cin, h, w, device="cuda", requires_grad=True)
wt = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.05).requires_grad_(True)
go = torch.randn(b, cout, h, w, device="cuda")
for what, leaves in (("grad_input", [x]), ("grad_weight", [wt]), ("both", [x, wt])):
    x.requires_grad_(any(x is t for t in leaves))          # the Function computes the gradients its inputs ask for
    wt.requires_grad_(any(wt is t for t in leaves))
    y = conv2d(x, wt, None, d)
    for _ in range(2):
        torch.autograd.grad(y, leaves, go, retain_graph=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        torch.autograd.grad(y, leaves, go, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    flop = 2.0 * b * h * w * cin * cout * 9 * len(leaves)
    ms = e0.elapsed_time(e1) / 5
    print(f"{what}: {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s (fp32 CUDA cores)")
