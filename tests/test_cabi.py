"""The C-ABI shared library: builds for sm_100a, loads, and exports every symbol
include/pp_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import pp_b200
from pp_b200 import _lib, build

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pp_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), "libpp_b200.so does not export %s" % s
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms      # the ctypes binding covers exactly the header


def test_version_and_error_strings():
    L = _lib.load()
    assert L.pp_version() == 100
    assert L.pp_error_string(0) == b"ok"
    assert b"workspace" in L.pp_error_string(2)


def test_workspace_queries_and_argument_validation_without_gpu():
    L = _lib.load()
    grid = pp_b200.PPConfig().grid()
    n = L.pp_pillarize_workspace_bytes(4, 280000, grid, 24000)
    assert 20e6 < n < 200e6
    assert L.pp_pillarize_workspace_bytes(0, 10, grid, 24000) == 0          # n_sweeps < 1
    assert L.pp_pillarize_workspace_bytes(65, 10, grid, 24000) == 0         # > PP_MAX_SWEEPS
    bad = pp_b200.PPConfig(x_step=0.0).grid()
    assert L.pp_pillarize_workspace_bytes(1, 10, bad, 24000) == 0
    assert L.pp_pfn_workspace_bytes(4, 24000, 64, 600, 600) > 4 * 24000 * 2 * 64 * 4
    assert L.pp_assign_targets_workspace_bytes(4, 540000, 400, None) >= 4 * 540000 * 12
    # invalid arguments are rejected before any CUDA call
    assert L.pp_pillarize(None, 7, 4, 1, _lib.i64_array([0, 0]), 1, grid, 200, 24000, None, None, None,
                          None, None, None, 0, None) == 1
    assert L.pp_pfn_forward(None, 1, 9, 1, 1, 64, None, None, None, None, None, None, None, 1, 0.1, 1e-5,
                            None, None, 0, None) == 1
    assert L.pp_make_ious(None, None, None, None, -1, 0, None, None, None) == 1


def test_product_never_imports_the_oracle():
    root = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_host_mirrors_fail_loudly_without_gpu():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pp_b200 import pillars, model, pipeline
    with pytest.raises(_lib.PPError):
        pillars.create_pillars(np.zeros((3, 4)), np.zeros((4, 4, 9)), np.zeros((4, 3)), 4, 4,
                               .2, .2, -60, -60, -10, 60, 60, 10, 600)
    with pytest.raises(_lib.PPError):
        model.PPFeatureNet(9, 64)(torch.zeros(1, 9, 4, 4))
    with pytest.raises(_lib.PPError):
        pipeline.InputPath()


def test_input_path_workspace_and_validation_without_gpu():
    """pp_input_path: workspace query covers both stages; bad arguments are rejected before any CUDA call."""
    L = _lib.load()
    grid = pp_b200.PPConfig().grid()
    k1 = L.pp_pillarize_workspace_bytes(4, 280000, grid, 24000)
    n = L.pp_input_path_workspace_bytes(4, 280000, grid, 200, 24000, 64, 600, 600, 0)
    assert n > k1 + 4 * 24000 * 64 * 4 + 24000 * 3 * 64 * 4    # K1 state + ext rows + padding table of the sparse PFN
    prep = L.pp_mean_prepared_bytes(24000, 200)
    assert prep >= 24000 * 200 * 48 + 100 * 8                  # operand image (48 B per slot) + moments
    assert L.pp_input_path_workspace_bytes(4, 280000, grid, 200, 24000, 64, 600, 600, 1) >= n + prep
    assert L.pp_input_path_workspace_bytes(4, 280000, grid, 200, 24000, 0, 600, 600, 0) == 0
    assert L.pp_input_path_workspace_bytes(0, 280000, grid, 200, 24000, 64, 600, 600, 0) == 0
    assert L.pp_mean_prepared_bytes(0, 200) == 0
    assert L.pp_mean_prepare(None, 24000, 200, None, 0, None) == 1     # PP_ERR_INVALID_ARG before any CUDA call
    args = [None, 0, 4, 1, _lib.i64_array([0, 0]), 1, grid, 200, 24000, None, None, 64,
            None, None, None, None, None, None, None, 1, 0.1, 1e-5, 600, 600,
            None, None, None, None, None, None, 0, 3, None]
    assert L.pp_input_path(*args) == 1                           # NULL weights: PP_ERR_INVALID_ARG


def test_packed_host_batch_layout_is_aligned_and_disjoint():
    """One pinned blob per batch: every section 256-byte aligned, sections disjoint, sizes as declared."""
    from pp_b200.pipeline import InputPath
    off, total = InputPath._blob_layout(271631, 5, 400)
    order = ["points", "corners", "centers", "wlh", "yaw", "cls"]
    assert list(off) == order
    end = 0
    for name in order:
        o, n = off[name]
        assert o % 256 == 0 and o >= end
        end = o + n
    assert total >= end and total % 256 == 0
    assert off["points"][1] == 271631 * 5 * 4 and off["corners"][1] == 400 * 8 * 8 and off["cls"][1] == 400 * 4
    off0, total0 = InputPath._blob_layout(0, 5, 0)             # empty batch still has one row per section
    assert off0["points"][1] == 5 * 4 and total0 > 0


def test_training_side_entries_validate_arguments_without_gpu():
    """pp_loss / pp_loss_list / pp_pfn_backward / pp_scatter_backward / pp_aggregate_sweeps /
    pp_input_path_backward / pp_assign_targets_list: workspace queries and NULL / shape rejection before any CUDA
    call (return code 1 = PP_ERR_INVALID_ARG)."""
    L = _lib.load()
    grid = pp_b200.PPConfig().grid()
    assert L.pp_loss_workspace_bytes(4, 300, 300, 6) > 4 * 300 * 300 * 6 * 4          # holds the positives list
    assert L.pp_loss_workspace_bytes(0, 300, 300, 6) == 0
    assert L.pp_pfn_backward_workspace_bytes(4, 24000, 64) > 0
    assert L.pp_pfn_backward_workspace_bytes(4, 24000, 32) == 0                        # C == 64 only
    assert L.pp_input_path_backward_workspace_bytes(4, 24000, 64) > 4 * 24000 * 64 * 8  # per-sweep suffix arg-max rows
    assert L.pp_input_path_backward_workspace_bytes(65, 24000, 64) == 0
    tail = (2, 300, 300, 6, 9, 8, 2.0, 25.0, 250.0, 1.0, 0.0, None, None, None, None, None, 0, None)
    assert L.pp_loss(None, None, None, None, *tail) == 1
    assert L.pp_loss_list(None, None, None, None, None, None, *tail) == 1
    assert L.pp_loss_scale_grads(None, 8, None, 0, None, None, None) == 1
    assert L.pp_pfn_backward(None, 1, 9, 8, 16, 64, None, None, None, None, None, 1, 1e-5, None, None, 0, 0,
                             None, None, None, None, None, 0, None) == 1
    assert L.pp_scatter_backward(None, None, 1, 64, 8, 600, 600, None, None) == 1
    assert L.pp_aggregate_sweeps(None, 10, 5, None, 1, None, 0.001, None, None) == 1
    assert L.pp_input_path_backward(_lib.i64_array([0, 0]), 1, grid, 200, 24000, None, 64, None, None, None, None, None,
                                    1, 1e-5, 600, 600, None, None, None, None, None, None, None, None, 0, None, 0,
                                    None) == 1
    assert L.pp_assign_targets_list(None, None, None, None, None, 540000, None, None, None, None, None,
                                    _lib.i64_array([0, 0]), 1, 9, 0.6, None, None, None, None, 16, None, None, None,
                                    None, None, None, 0, None) == 1


def test_packed_batch_with_transforms_layout():
    from pp_b200.pipeline import InputPath
    off, total = InputPath._blob_layout(1000, 5, 10, F=3)
    assert list(off)[-2:] == ["xforms", "file_offsets"]
    assert off["xforms"] == (off["xforms"][0], 3 * 12 * 8) and off["file_offsets"][1] == 4 * 8
    assert off["xforms"][0] % 256 == 0 and off["file_offsets"][0] % 256 == 0 and total % 256 == 0
    assert list(InputPath._blob_layout(1000, 5, 10)[0]) == list(off)[:-2]             # F = 0: unchanged layout


def test_header_is_plain_c():
    """include/pp_b200.h is the FFI boundary: it must compile as C99 and as C++ with no torch / CUDA headers."""
    import shutil
    import subprocess
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "pp_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], check=True)
    subprocess.run(["g++", "-fsyntax-only", "-x", "c++", "-Wall", "-Werror", hdr], check=True)
    includes = [l.strip() for l in open(hdr) if l.strip().startswith("#include")]
    assert includes and all(i in ("#include <stddef.h>", "#include <stdint.h>") for i in includes), includes
