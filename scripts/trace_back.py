"""Phase timeline of tc_back (CTA 0) from otp_debug_trace: per event the median cycle offset
from the tile start and the median duration since the previous event of the same role.

    python scripts/trace_back.py [batch] [tokens]
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
blk = TransformerBlock(136, 2, n_ds_strides=(1, 1), proj_pdrop=0.1, path_pdrop=0.1)
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
buf = (C.c_ulonglong * 4096)()
_lib.check(lib.otp_debug_trace_read(buf, 4096), "trace_read")
arr = np.frombuffer(buf, dtype=np.uint64).reshape(2, 2048)
for role, name in enumerate(("epilogue warp 0", "control")):
    ev = [(int(v) & 0xFF, int(v) >> 8) for v in arr[role] if v]
    tiles, cur = [], []
    for e, c in ev:
        if e == 0 and cur:
            tiles.append(cur)
            cur = []
        cur.append((e, c))
    if cur:
        tiles.append(cur)
    print(f"== {name}: {len(tiles)} tiles, tile period (cycles):",
          [tiles[i + 1][0][1] - tiles[i][0][1] for i in range(len(tiles) - 1)])
    offs, durs = {}, {}
    for tl in tiles[1:]:   # skip the cold first tile
        t0 = tl[0][1]
        for i, (e, c) in enumerate(tl):
            offs.setdefault(e, []).append(c - t0)
            if i:
                durs.setdefault(e, []).append(c - tl[i - 1][1])
    order = [e for e, _ in tiles[-1]]
    print("event: offset / since-previous (median cycles)")
    print("  ".join(f"{e}:{int(statistics.median(offs[e]))}/{int(statistics.median(durs.get(e, [0])))}" for e in order))
