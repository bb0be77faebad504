"""Run one C=136 TransformerBlock (bench shape: 32 clips x 6912 tokens, fp16 operands) a few
times -- the target of `ncu -k regex:tc_(front|back|apply)` captures (profiles/).

    python scripts/prof_block.py [batch] [tokens] [stride] [iters]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otpose_b200.model.blocks import TransformerBlock  # noqa: E402
from otpose_b200.utils import synthetic as syn  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
t = int(sys.argv[2]) if len(sys.argv) > 2 else 6912
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
blk = TransformerBlock(136, 2, n_ds_strides=(stride, stride), proj_pdrop=0.1, path_pdrop=0.1)
blk.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in blk.state_dict().items()}, seed=3))
blk = blk.cuda().eval()
x = torch.from_numpy(np.random.default_rng(0).standard_normal((b, 136, t)).astype(np.float32)).cuda()
with torch.no_grad():
    for _ in range(iters):
        y = blk(x, precision="fp16")
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.abs().max()))
