"""Phase timeline of tc_front (CTA (0,0), compute warp 0) from otp_debug_trace: per event the median
duration since the previous event over the CTA's tiles.

    NVCC_EXTRA=-DOTP_FRONT_TRACE python -m otpose_b200.build --force   # the tracer is compiled out by default
    python scripts/trace_front.py [batch] [tokens] [stride]
"""
import ctypes as C
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otpose_b200 import _lib  # noqa: E402
from otpose_b200.model.blocks import TransformerBlock  # noqa: E402
from otpose_b200.utils import synthetic as syn  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
t = int(sys.argv[2]) if len(sys.argv) > 2 else 6912
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
blk = TransformerBlock(136, 2, n_ds_strides=(stride, stride), proj_pdrop=0.1, path_pdrop=0.1)
blk.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in blk.state_dict().items()}, seed=3))
blk = blk.cuda().eval()
x = torch.from_numpy(np.random.default_rng(0).standard_normal((b, 136, t)).astype(np.float32)).cuda()
lib = _lib.load()
with torch.no_grad():
    for _ in range(3):
        blk(x, precision="fp16")
    torch.cuda.synchronize()
    lib.otp_debug_trace(1)
    blk(x, precision="fp16")
    torch.cuda.synchronize()
    lib.otp_debug_trace(0)
buf = (C.c_ulonglong * 8192)()
_lib.check(lib.otp_debug_trace_read(buf, 8192), "trace_read")
arr = np.frombuffer(buf, dtype=np.uint64).reshape(4, 2048)
ev = [(int(v) & 0xFF, int(v) >> 8) for v in arr[3] if v]
tiles, cur = [], []
for e, c in ev:
    if e == 0 and cur:
        tiles.append(cur)
        cur = []
    cur.append((e, c))
if cur:
    tiles.append(cur)
print("tiles:", len(tiles), "tile period:", [tiles[i + 1][0][1] - tiles[i][0][1] for i in range(len(tiles) - 1)])
durs = {}
order = []
for tl in tiles[1:]:
    for (e0, c0), (e1, c1) in zip(tl, tl[1:]):
        k = (e0, e1)
        if k not in durs:
            durs[k] = []
            order.append(k)
        durs[k].append(c1 - c0)
print("  ".join(f"{a}->{b_}:{int(statistics.median(v))}" for (a, b_), v in ((k, durs[k]) for k in order)))
