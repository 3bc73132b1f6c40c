"""Build libpp_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The shared object lands next to this file so that it travels to the GPU box with the repo
snapshot.  nvcc cross-compiles for sm_100a without a GPU.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpp_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--threads", "4",
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def _deps():
    return sources() + sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [
        os.path.join(HERE, "..", "include", "pp_b200.h"), os.path.join(HERE, "..", "include", "pp_b200_debug.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, debug=None):
    """Compile every CUDA source for sm_100a into libpp_b200.so.  Returns the library path.
    ``debug`` (default: environment PP_DEBUG=1) adds -DPP_DEBUG: the development entry points of
    include/pp_b200_debug.h; the product build does not export them."""
    if debug is None:
        debug = os.environ.get("PP_DEBUG", "0") == "1"
    if not force and not is_stale() and not debug:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = ([nvcc] + NVCC_FLAGS + (["-DPP_DEBUG"] if debug else []) + (["-Xptxas", "-v"] if verbose else []) +
           ["-o", LIB_PATH] + sources())
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv, debug=("--debug" in sys.argv) or None))
