"""UMMA (tcgen05.mma kind::f16, M=128, K=16) rate on this GPU for the N the block kernels use."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otpose_b200 import _lib  # noqa: E402

lib = _lib.load()
for ctas in (1, 148):
    for n, ks in ((48, 9), (64, 9), (96, 9), (144, 9), (144, 3), (144, 4), (192, 9), (256, 9), (80, 8)):
        reps = 64
        buf = (C.c_longlong * (2 * ctas))()
        _lib.check(lib.otp_debug_umma_rate(n, ks, reps, ctas, buf), "rate")
        a = np.frombuffer(buf, dtype=np.int64).reshape(ctas, 2)
        per = a.mean(0) / (reps * ks)
        print(f"ctas={ctas:4d} N={n:3d} ksteps={ks}: issue {per[0]:6.1f} cyc/UMMA, issue+drain {per[1]:6.1f} cyc/UMMA "
              f"(ideal {n / 2:.0f})")
