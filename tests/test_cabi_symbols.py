"""CPU-only: the C-ABI library builds, loads, and exports every symbol include/wisp_b200.h
declares; the ctypes table covers the same set; no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wisp_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wisp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("wisp_ctx_create", "wisp_com_nodes", "wisp_traj_alignment_and_averaging",
                 "wisp_cartesian_covariance", "wisp_pearson_from_covariance", "wisp_contact_map",
                 "wisp_func_pearson", "wisp_pipeline", "wisp_apsp", "wisp_get_paths",
                 "wisp_dev_gram", "wisp_dev_contact_sum", "wisp_dev_epilogue", "wisp_dev_apsp"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from wisp_mdanalysis_implementation_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), "libwisp_b200.so does not export %s" % name


def test_ctypes_table_matches_header():
    from wisp_mdanalysis_implementation_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.lib()
    assert lib.wisp_abi_version() == 1
    # layout helpers are pure host arithmetic
    assert lib.wisp_node_ld(300) == 3 * 384
    assert lib.wisp_node_ld(128) == 384
    assert lib.wisp_frames_pad(1000) == 1008
    assert lib.wisp_gram_ld(900) == 1024


def test_no_cpu_fallback_without_gpu():
    """Creating a context must fail loudly when there is no B200 (this container)."""
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.WispError):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "wisp_mdanalysis_implementation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_stage_functions_fail_loudly_without_gpu(tmp_path):
    """Every reference-facing stage function raises instead of computing anything on the CPU when
    there is no B200: nothing in the product path can quietly fall back."""
    import sys
    import numpy as np
    import torch
    from wisp_mdanalysis_implementation_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib._default_ctx.clear()
    pkg = os.path.join(ROOT, "wisp_mdanalysis_implementation_b200")
    sys.path.insert(0, pkg)
    try:
        import adjacency_matrix_functions as amf
        import functionalize_adj_matrix_functions as faf
        import network_analysis_functions as naf
        import trajectory_analysis_functions as taf
    finally:
        sys.path.remove(pkg)
    out = str(tmp_path) + os.sep
    rng = np.random.default_rng(0)
    nodes = rng.normal(size=(8, 4, 3))
    w = np.abs(rng.normal(size=(5, 5)))
    w = w + w.T
    np.fill_diagonal(w, 0.0)
    calls = [
        lambda: taf.cartesian_covariance(nodes, nodes.mean(axis=0), out),
        lambda: taf.calc_contact_map(nodes, out, 9.0),
        lambda: amf.pearson_correlation_analysis(np.eye(12), out),
        lambda: amf.linear_mutual_information_analysis(np.eye(12), out),
        lambda: faf.func_pearson_correlation(np.full((4, 4), 0.5), out),
        lambda: naf.get_paths(w, [0], [4], 2),
    ]
    for call in calls:
        with pytest.raises(_lib.WispError):
            call()
    assert not any(f.endswith(".dat") for f in os.listdir(tmp_path))      # and nothing was written
