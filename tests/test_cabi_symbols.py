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
