"""ORACLE (test infrastructure): ctypes front-end of oracle/algos_oracle.c.

Mirrors the reference call surface of /root/reference/graphormer/algos.pyx
(`floyd_warshall`, `gen_edge_input`) so tests read like the reference's call
sites (wrapper.py:55-60).  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "algos_oracle.c")
LIB = os.path.join(HERE, "_build", "liboracle_algos.so")
_lib = None


def build(force=False):
    if os.path.exists(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_floyd_warshall.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.oracle_floyd_warshall.restype = None
        _lib.oracle_gen_edge_input.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.oracle_gen_edge_input.restype = ctypes.c_int
    return _lib


def floyd_warshall(adjacency_matrix):
    """algos.pyx:9-54.  bool/int [n,n] -> (M int64 [n,n], path int64 [n,n])."""
    lib = _load()
    a = np.ascontiguousarray(adjacency_matrix)
    assert a.ndim == 2 and a.shape[0] == a.shape[1]
    n = a.shape[0]
    a8 = np.ascontiguousarray(a != 0, dtype=np.uint8)
    M = np.empty((n, n), np.int64)
    path = np.empty((n, n), np.int64)
    lib.oracle_floyd_warshall(a8.ctypes.data, n, M.ctypes.data, path.ctypes.data)
    return M, path


def gen_edge_input(max_dist, path, edge_feat, hop_cap=None):
    """algos.pyx:65-96.  Returns float32 [n,n,min(max_dist,hop_cap),F] (-1 filled).

    hop_cap=None reproduces the reference shape exactly; a cap is the same as
    slicing the reference result [:, :, :hop_cap] (collator.py:323)."""
    lib = _load()
    path = np.ascontiguousarray(path, dtype=np.int64)
    ef = np.ascontiguousarray(edge_feat, dtype=np.int64)
    n = path.shape[0]
    F = ef.shape[-1]
    max_dist = int(max_dist)
    cap = max_dist if hop_cap is None else int(hop_cap)
    hops = min(max_dist, cap)
    out = np.empty((n, n, hops, F), np.float32)
    lib.oracle_gen_edge_input(max_dist, path.ctypes.data, ef.ctypes.data, n, F, cap, out.ctypes.data)
    return out
