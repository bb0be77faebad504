import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from otpose_b200.model.RSB import conv_bn_relu
torch.manual_seed(0)
def run(cin, cout, k, b, h, w, add, res, prec):
    m = conv_bn_relu(cin, cout, k, 1, k // 2).cuda().eval()
    with torch.no_grad():
        m.bn.running_mean.normal_(); m.bn.running_var.uniform_(0.5, 2); m.bn.weight.uniform_(0.5,1.5); m.bn.bias.normal_()
    x = torch.randn(b, cin, h, w, device='cuda'); xa = torch.randn_like(x) if add else None
    r = torch.randn(b, cout, h, w, device='cuda') if res else None
    outs = {}
    for p in ('fp32', prec):
        m.precision = p
        y = torch.empty(b, cout, h, w, device='cuda')
        m.run(x.data_ptr(), cin*h*w, y.data_ptr(), cout*h*w, b, h, w, x_add=xa.data_ptr() if add else None, x_add_bs=cin*h*w,
              residual=r.data_ptr() if res else None, residual_bs=cout*h*w)
        outs[p] = y
    wt, bias = m.folded()
    ref = F.conv2d(x + (xa if add else 0), wt, bias, padding=k//2) + (r if res else 0)
    ref = F.relu(ref)
    e32 = float((outs['fp32']-ref).abs().max()/ref.abs().max()); e16 = float((outs[prec]-ref).abs().max()/ref.abs().max())
    print(cin, cout, k, b, h, w, add, res, prec, 'fp32 err %.2e tc err %.2e' % (e32, e16), flush=True)
for prec in ('fp16', 'bf16'):
    run(20, 20, 3, 2, 96, 72, True, False, prec)
    run(13, 13, 3, 2, 96, 72, False, False, prec)
    run(6, 6, 3, 3, 16, 8, True, False, prec)
    run(51, 80, 1, 2, 96, 72, False, False, prec)
    run(80, 32, 1, 2, 24, 16, False, True, prec)
    run(17, 24, 1, 1, 8, 8, False, False, prec)
