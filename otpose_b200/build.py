"""Build the C-ABI shared library ``otpose_b200/lib/libotpose_b200.so`` in-tree.

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libotpose_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"] + os.environ.get("NVCC_EXTRA", "").split()


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def _compile(src, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + _deps())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


SHIM_SRC = os.path.join(HERE, "shim", "deform_conv_cuda.cpp")
SHIM = os.path.join(LIBDIR, "deform_conv_cuda.so")


def build_shim(force: bool = False) -> str:
    """Optional target: the pybind11 module ``deform_conv_cuda`` with the reference's own entry points
    (thirdparty/deform_conv/src/deform_conv_cuda.cpp:666-680) on top of the C ABI, built in-tree as
    ``otpose_b200/lib/deform_conv_cuda.so`` (g++ against the installed torch headers; no nvcc needed)."""
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce
    build()
    deps = [SHIM_SRC, os.path.join(HERE, "..", "include", "otpose_b200.h")]
    if not force and os.path.exists(SHIM) and os.path.getmtime(SHIM) >= max(os.path.getmtime(p) for p in deps):
        return SHIM
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + \
          [f"-I{sysconfig.get_paths()['include']}", f"-I{os.path.join(HERE, '..', 'include')}"]
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", SHIM_SRC, "-o", SHIM,
           "-DTORCH_EXTENSION_NAME=deform_conv_cuda", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc,
           f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python", "-lc10", "-lc10_cuda",
           f"-L{LIBDIR}", "-lotpose_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"shim build failed:\n{r.stdout}\n{r.stderr[-4000:]}")
    return SHIM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--shim" in sys.argv:
        print(build_shim(force="--force" in sys.argv))
