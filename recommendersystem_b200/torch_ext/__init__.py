"""`torch.ops.graft.*`: the TORCH_LIBRARY extension over libgraft.so's C ABI (graft_torch.cpp).

build() compiles it in-tree (recommendersystem_b200/torch_ext/graft_torch.so) and links it against ../libgraft.so;
load() registers the operators (CUDA kernels = the C ABI, Meta kernels for tracing, autograd formulas).  Importing this
module without the built library raises: there is no Python fallback for the operators."""
from __future__ import annotations

import os
import subprocess
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "graft_torch.so")
_loaded = False


def build(verbose=False):
    """Compiles graft_torch.cpp with the host compiler against torch's headers (no nvcc needed: nothing computes here)."""
    from torch.utils import cpp_extension as ce
    src = os.path.join(_HERE, "graft_torch.cpp")
    lib_dir = os.path.dirname(_HERE)
    if os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib_dir, "..", "include", "graft.h"))):
        return SO_PATH
    inc = [f"-I{p}" for p in ce.include_paths("cuda")]
    libs = [f"-L{p}" for p in ce.library_paths("cuda")]
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *inc, src, "-o", SO_PATH, *libs, f"-L{lib_dir}", "-l:libgraft.so", "-Wl,-rpath,$ORIGIN/..", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
           "-ltorch", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-6000:])
        raise RuntimeError("graft_torch.so build failed")
    if verbose:
        print("built", SO_PATH)
    return SO_PATH


def load():
    global _loaded
    if _loaded:
        return torch.ops.graft
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    torch.ops.load_library(SO_PATH)
    _loaded = True
    return torch.ops.graft
