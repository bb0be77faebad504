#!/usr/bin/env python
"""Offline install of the UNMODIFIED reference into ``baseline/_ref`` (git-ignored, travels to the GPU box).

The sanctioned route -- ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref
/root/reference`` -- fails here: the reference's setup.py is detectron2's and needs a ``detectron2/`` source
tree that the repository does not contain ("error while generating package metadata"; recorded in
DESIGN.md).  This script therefore places the reference's importable Python packages for the hot path
(``model/``, ``thirdparty/deform_conv``'s Python side, ``utils/heatmap.py``, ``utils/transform.py``) under
``baseline/_ref`` byte for byte.  Nothing under ``baseline/_ref`` is tracked, imported by the product
(``otpose_b200/``) or read by ``bench.py``; only ``tests/test_gpu_shim.py`` uses it, to run the reference's
own ``OTPose.forward`` / ``ModulatedDeformConvFunction`` on top of the drop-ins.

    python scripts/install_reference.py [/root/reference]
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("OTPOSE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["model/__init__.py", "model/OTPose.py", "model/ConvVideoTransformer.py", "model/blocks.py", "model/RSB.py",
         "model/layers.py", "model/HRNet.py", "model/base_backbone.py", "model/loss.py",
         "thirdparty/__init__.py", "thirdparty/deform_conv/__init__.py",
         "thirdparty/deform_conv/functions/__init__.py", "thirdparty/deform_conv/functions/deform_conv.py",
         "thirdparty/deform_conv/functions/deform_pool.py", "thirdparty/deform_conv/modules/__init__.py",
         "thirdparty/deform_conv/modules/deform_conv.py", "thirdparty/deform_conv/modules/deform_pool.py",
         "utils/__init__.py", "utils/heatmap.py", "utils/transform.py"]


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found")
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, REF], capture_output=True, text=True)
    if r.returncode == 0:
        print("pip install succeeded into", DST)
        return
    print("pip install failed (", (r.stderr or r.stdout).strip().splitlines()[-1][:120], "); copying the packages")
    n = 0
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    print(f"placed {n} reference files under {DST}")


if __name__ == "__main__":
    main()
