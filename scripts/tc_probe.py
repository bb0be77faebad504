"""Exploration helper (GPU): try UMMA descriptor hypotheses one per process so a
faulting variant cannot poison the others.  usage: python scripts/tc_probe.py [variant]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = {
    # name: (a_mn, b_mn, a(lbo,sbo), b(lbo,sbo))  -- 'rs'/'cs' symbolic
    "A_mn_v1": (1, 0, ("rs", "cs"), None),
    "A_mn_v2": (1, 0, ("cs", "rs"), None),
    "B_mn_v1": (0, 1, None, ("rs", "cs")),
    "B_mn_v2": (0, 1, None, ("cs", "rs")),
    "AB_mn_v1": (1, 1, ("rs", "cs"), ("rs", "cs")),
    "AB_mn_v2": (1, 1, ("cs", "rs"), ("cs", "rs")),
}


def one(name):
    import numpy as np
    import torch
    from test_gpu_tc import bf16_round, cm_image, relerr, run
    a_mn, b_mn, am, bm = VARIANTS[name]
    r = np.random.default_rng(3)
    ch, c0, n = 192, 64, 80
    rs, cs = (ch // 8) * 128, 128
    sym = {"rs": rs, "cs": cs}
    q = bf16_round(r.standard_normal((128, ch)).astype(np.float32))     # [token][channel]
    kk = bf16_round(r.standard_normal((128, ch)).astype(np.float32))
    if a_mn:   # A' = [M=channel][K=token] view of q
        a_img, a_ref = cm_image(q, rs, cs), q.float()[:, c0:c0 + 128].T
        a = dict(off=(c0 // 8) * cs, lbo=sym[am[0]], sbo=sym[am[1]], kstep=2 * rs, mn=1)
        ksteps = 8
    else:      # A = [M=token? no: rows][K] plain K-major 128 x 128
        a_img, a_ref = cm_image(q[:, :128].contiguous(), 16 * 128, 128), q.float()[:, :128]
        a = dict(off=0, lbo=128, sbo=16 * 128, kstep=256, mn=0)
        ksteps = 8
    if b_mn:
        b_img, b_ref = cm_image(kk, rs, cs), kk.float()[:, c0:c0 + n].T        # [N][K=token]
        b = dict(off=(c0 // 8) * cs, lbo=sym[bm[0]], sbo=sym[bm[1]], kstep=2 * rs, mn=1)
    else:
        w = bf16_round(r.standard_normal((n, 128)).astype(np.float32))
        b_img, b_ref = cm_image(w, 16 * 128, 128), w.float()
        b = dict(off=0, lbo=128, sbo=16 * 128, kstep=256, mn=0)
    ref = a_ref @ b_ref.T
    got = run(a_img, b_img, n, ksteps, a, b)
    print(json.dumps({"variant": name, "relerr": relerr(got, ref)}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1])
    else:
        for v in VARIANTS:
            p = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True, timeout=120)
            out = [l for l in p.stdout.splitlines() if l.startswith("{")]
            print(v, out[-1] if out else "CRASH: " + (p.stderr.strip().splitlines() or ["?"])[-1][:160], flush=True)
