"""(build the library with NVCC_EXTRA=-DOTP_BACK_TRACE first)  Phase timeline of tc_back (CTA 0) from otp_debug_trace: per event the median cycle offset
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
buf = (C.c_ulonglong * 6144)()
_lib.check(lib.otp_debug_trace_read(buf, 6144), "trace_read")
arr = np.frombuffer(buf, dtype=np.uint64).reshape(3, 2048)
names = ("E0", "CT", "E15")
per_role = []
for role in range(3):
    ev = [(int(v) & 0xFF, int(v) >> 8) for v in arr[role] if v]
    tiles, cur = [], []
    for e, c in ev:
        if e == 0 and cur:
            tiles.append(cur)
            cur = []
        cur.append((e, c))
    if cur:
        tiles.append(cur)
    per_role.append(tiles)
    print(f"== {names[role]}: {len(tiles)} tiles, tile period (cycles):",
          [tiles[i + 1][0][1] - tiles[i][0][1] for i in range(len(tiles) - 1)])
# merged absolute timeline of one steady-state tile (clock64 is per SM: directly comparable)
k = min(len(t) for t in per_role) // 2
origin = per_role[1][k][0][1]
merged = sorted((c - origin, names[r], e) for r in range(3) for e, c in per_role[r][k])
print(f"-- tile #{k}: cycle offset from the control warp's tile start, role:event")
print("  ".join(f"{c}:{n}:{e}" for c, n, e in merged if e not in (60, 61)))
# control warp: cycles spent waiting for weight pieces (events 60 -> 61) per tile
for tl in per_role[1][2:8]:
    waits = [b[1] - a[1] for a, b in zip(tl, tl[1:]) if a[0] == 60 and b[0] == 61]
    print("piece waits: n =", len(waits), "total", sum(waits), "max", max(waits) if waits else 0,
          "over 200 cycles:", [w for w in waits if w > 200])
