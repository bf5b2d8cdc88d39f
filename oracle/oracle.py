"""ctypes front-end of the CPU checker (oracle/gnna_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (tests/test_boundary.py greps for it).

All arrays are numpy, C-contiguous; int32 for indices, float32 for features - the layout the
reference kernels take.  Function names follow the extension surface of the
reference (GNNAdvisor/GNNConv/GNNAdvisor.cpp:253-263).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgnna_oracle.so")

MODE_SAG, MODE_GCN, MODE_GIN = 0, 1, 2


def build(force=False):
    """Compile the checker with the Makefile beside it (gcc, OpenMP)."""
    src = os.path.join(_HERE, "gnna_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_count_parts.restype = ctypes.c_int64
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_aggregate_mt.restype = ctypes.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


I64 = ctypes.c_int64


def num_threads():
    return int(lib().oracle_num_threads())


# ---------------------------------------------------------------- build_part
def count_parts(part_size, indptr):
    indptr = _i32(indptr)
    return int(lib().oracle_count_parts(int(part_size), _p(indptr), I64(len(indptr) - 1)))


def build_part_f32(part_size, indptr):
    """What the reference returns: float32 tables, F5 + F6 included (GNNAdvisor.cpp:210-251)."""
    indptr = _i32(indptr)
    n = len(indptr) - 1
    P = count_parts(part_size, indptr)
    pp = np.zeros(P + 1, dtype=np.float32)
    pn = np.zeros(P, dtype=np.float32)
    lib().oracle_build_part_f32(int(part_size), _p(indptr), I64(n), _p(pp), _p(pn), I64(P))
    return pp, pn


def build_part(part_size, indptr, exact=False):
    """int32 tables: `compat` = reference float tables cast with .int(); `exact` = intended table."""
    indptr = _i32(indptr)
    n = len(indptr) - 1
    P = count_parts(part_size, indptr)
    pp = np.zeros(P + 1, dtype=np.int32)
    pn = np.zeros(P, dtype=np.int32)
    fn = lib().oracle_build_part_i32_exact if exact else lib().oracle_build_part_i32_compat
    fn(int(part_size), _p(indptr), I64(n), _p(pp), _p(pn), I64(P))
    return pp, pn


def degrees(indptr):
    indptr = _i32(indptr)
    out = np.empty(len(indptr) - 1, dtype=np.float32)
    lib().oracle_degrees(_p(indptr), I64(len(out)), _p(out))
    return out


# ---------------------------------------------------------------- aggregation
def aggregate(mode, X, col_idx, deg, eps, part_ptr, part2node, threads=0):
    """threads=0: the literal single-thread loop over groups; >0: the node-aligned OpenMP
    version (bitwise the same result); <0: all host threads."""
    X = _f32(X)
    n, d = X.shape
    col_idx, part_ptr, part2node = _i32(col_idx), _i32(part_ptr), _i32(part2node)
    deg = None if deg is None else _f32(deg)
    out = np.empty_like(X)
    P = len(part2node)
    if threads == 0:
        lib().oracle_aggregate(int(mode), _p(X), _p(out), _p(col_idx), _p(deg), ctypes.c_float(eps),
                               _p(part_ptr), _p(part2node), I64(n), I64(d), I64(P))
    else:
        rc = lib().oracle_aggregate_mt(int(mode), _p(X), _p(out), _p(col_idx), _p(deg), ctypes.c_float(eps),
                                       _p(part_ptr), _p(part2node), I64(n), I64(d), I64(P),
                                       ctypes.c_int(max(threads, 0)))
        if rc != 0:
            raise ValueError("part2node is not sorted; use threads=0")
    return out


def mm(A, B, trans_a=False, trans_b=False, threads=0):
    A, B = _f32(A), _f32(B)
    m = A.shape[1] if trans_a else A.shape[0]
    k = A.shape[0] if trans_a else A.shape[1]
    n = B.shape[0] if trans_b else B.shape[1]
    C = np.empty((m, n), dtype=np.float32)
    lib().oracle_mm(_p(A), int(trans_a), _p(B), int(trans_b), _p(C), I64(m), I64(k), I64(n), ctypes.c_int(threads))
    return C


# ---------------------------------------------------------------- the extension surface
def SAG(X, row_ptr, col_idx, deg, part_ptr, part2node, part_size=0, dim_worker=0, warp_per_block=0, threads=0):
    return aggregate(MODE_SAG, X, col_idx, None, 1.0, part_ptr, part2node, threads)


def forward(X, W, row_ptr, col_idx, deg, part_ptr, part2node, part_size=0, dim_worker=0, warp_per_block=0, threads=0):
    T = mm(X, W, threads=threads)
    return [aggregate(MODE_GCN, T, col_idx, deg, 1.0, part_ptr, part2node, threads)]


def backward(d_out, X, W, row_ptr, col_idx, deg, part_ptr, part2node, part_size=0, dim_worker=0, warp_per_block=0, threads=0):
    G = aggregate(MODE_GCN, d_out, col_idx, deg, 1.0, part_ptr, part2node, threads)
    return [mm(G, W, trans_b=True, threads=threads), mm(X, G, trans_a=True, threads=threads)]


def forward_gin(X, W, row_ptr, col_idx, eps, part_ptr, part2node, part_size=0, dim_worker=0, warp_per_block=0, threads=0):
    S = aggregate(MODE_GIN, X, col_idx, None, eps, part_ptr, part2node, threads)
    return [mm(S, W, threads=threads), S]


def backward_gin(d_out, x_agg, W, row_ptr, col_idx, eps, part_ptr, part2node, part_size=0, dim_worker=0, warp_per_block=0, threads=0):
    d_w = mm(x_agg, d_out, trans_a=True, threads=threads)
    Pm = mm(d_out, W, trans_b=True, threads=threads)
    return [aggregate(MODE_GIN, Pm, col_idx, None, eps, part_ptr, part2node, threads), d_w]


# ---------------------------------------------------------------- closed forms (self-check of the oracle)
def closed_form(mode, X, row_ptr, col_idx, eps=1.0):
    """A@X / diag(n) A diag(n) @ X / eps*A@X in float64 via scipy (SURVEY.md 8c closed forms)."""
    import scipy.sparse as sp
    row_ptr, col_idx = np.asarray(row_ptr), np.asarray(col_idx)
    n = len(row_ptr) - 1
    A = sp.csr_matrix((np.ones(len(col_idx)), col_idx, row_ptr), shape=(n, n))
    X64 = np.asarray(X, dtype=np.float64)
    if mode == MODE_SAG:
        return A @ X64
    if mode == MODE_GIN:
        return float(np.float32(eps)) * (A @ X64)
    nrm = degrees(row_ptr).astype(np.float64)
    return nrm[:, None] * (A @ (nrm[:, None] * X64))
