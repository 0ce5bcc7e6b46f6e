"""CPU: the C-ABI library builds, loads without a GPU, and exports every symbol include/mobgt.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mobgt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mobgt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("mobgt_version", "mobgt_last_error", "mobgt_apsp_edge_input", "mobgt_gen_edge_input"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_built):
    L = ctypes.CDLL(lib_built)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_binding_table_matches_header(lib_built):
    from mobgt_b200 import _C
    assert sorted(_C.SIGNATURES) == declared_symbols()
    assert _C.lib().mobgt_version() >= 100


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    from mobgt_b200 import _C
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_C.MobgtError):
        _C.require_cuda()
    import numpy as np
    from mobgt_b200 import algos
    with pytest.raises(Exception):
        algos.floyd_warshall(np.zeros((3, 3), bool))


def test_bad_arguments_return_status_not_crash(lib_built):
    from mobgt_b200 import _C
    L = _C.lib()
    rc = L.mobgt_apsp_edge_input(None, None, None, None, 1, 4, 20, 20, 0, None, None, None, None, None)
    assert rc == -6                                       # MOBGT_ERR_NULL
    assert "null" in _C.last_error()


def test_binding_arity_matches_header(lib_built):
    """Every ctypes signature in mobgt_b200/_C.py has as many arguments as the declaration in include/mobgt.h (a drifted
    binding would pass garbage in the trailing arguments instead of failing loudly)."""
    from mobgt_b200 import _C
    src = open(os.path.join(ROOT, "include", "mobgt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = dict(re.findall(r"\b(mobgt_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    assert sorted(decls) == sorted(_C.SIGNATURES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_C.SIGNATURES[name]), (name, n, len(_C.SIGNATURES[name]))
