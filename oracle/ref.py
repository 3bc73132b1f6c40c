"""Loader for oracle/_ref/pillars*.so (TEST INFRASTRUCTURE ONLY).

That file is the reference's own data/pillars.cpp compiled unmodified (oracle/Makefile, target
``ref``) against the Boost stand-in headers of oracle/boost_shim.  It is built in the build
container (where /root/reference exists), is git-ignored, and travels to the GPU box with the
snapshot.  ``load()`` returns the module or None when the file is absent.
"""
import glob
import importlib.util
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def build(reference_root="/root/reference"):
    """Compile the reference module if its sources are present; returns True on success."""
    if not os.path.exists(os.path.join(reference_root, "data", "pillars.cpp")):
        return False
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "refpy", "REF=" + reference_root])
    return True


def load():
    global _mod
    if _mod is not None:
        return _mod
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "pillars*.so")))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("pillars", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod
