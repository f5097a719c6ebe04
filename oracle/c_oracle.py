"""ctypes loader for oracle/_build/liboracle.so (C restatement; TEST INFRASTRUCTURE ONLY).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle.so')


def build():
    subprocess.check_call(['make', '-C', _HERE, '-s'])


def _load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    f64p, f32p = C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.oracle_bpr_sgd_sequential_f64.restype = C.c_double
    lib.oracle_bpr_sgd_sequential_f64.argtypes = [f64p, f64p, C.c_int, C.c_int64, i32p, i32p, i32p,
                                                  C.c_double, C.c_double, C.c_double]
    lib.oracle_bpr_sgd_sequential_f32.restype = C.c_double
    lib.oracle_bpr_sgd_sequential_f32.argtypes = [f32p, f32p, C.c_int, C.c_int64, i32p, i32p, i32p,
                                                  C.c_float, C.c_float, C.c_float]
    for name, fp, ft in (('oracle_mf_sgd_sequential_f64', f64p, C.c_double), ('oracle_mf_sgd_sequential_f32', f32p, C.c_float)):
        fn = getattr(lib, name)
        fn.restype = C.c_double
        fn.argtypes = [C.c_int, fp, fp, C.c_int, C.c_int64, i32p, i32p, fp, ft, ft, ft, fp, fp, ft, ft]
    lib.oracle_spmm_csr_f32.restype = None
    lib.oracle_spmm_csr_f32.argtypes = [C.c_int32, i64p, i32p, f32p, f32p, f32p, C.c_int]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def bpr_sgd_sequential(P, Q, u, i, j, lr, reg_u, reg_i):
    """In place on P, Q (float64 or float32 C-contiguous); returns sum(-ln s)."""
    u = np.ascontiguousarray(u, np.int32); i = np.ascontiguousarray(i, np.int32)
    j = np.ascontiguousarray(j, np.int32)
    assert P.flags.c_contiguous and Q.flags.c_contiguous and P.dtype == Q.dtype
    if P.dtype == np.float64:
        return lib().oracle_bpr_sgd_sequential_f64(_p(P, C.c_double), _p(Q, C.c_double), P.shape[1],
                                                   len(u), _p(u, C.c_int32), _p(i, C.c_int32),
                                                   _p(j, C.c_int32), lr, reg_u, reg_i)
    assert P.dtype == np.float32
    return lib().oracle_bpr_sgd_sequential_f32(_p(P, C.c_float), _p(Q, C.c_float), P.shape[1],
                                               len(u), _p(u, C.c_int32), _p(i, C.c_int32),
                                               _p(j, C.c_int32), lr, reg_u, reg_i)


def spmm_csr(rowptr, cols, vals, X):
    rowptr = np.ascontiguousarray(rowptr, np.int64); cols = np.ascontiguousarray(cols, np.int32)
    vals = np.ascontiguousarray(vals, np.float32); X = np.ascontiguousarray(X, np.float32)
    Y = np.empty((len(rowptr) - 1, X.shape[1]), np.float32)
    lib().oracle_spmm_csr_f32(len(rowptr) - 1, _p(rowptr, C.c_int64), _p(cols, C.c_int32),
                              _p(vals, C.c_float), _p(X, C.c_float), _p(Y, C.c_float), X.shape[1])
    return Y


def mf_sgd_sequential(kind, P, Q, u, i, r, lr, reg_u=0.0, reg_i=0.0, Bu=None, Bi=None, reg_b=0.0, global_mean=0.0):
    """In place on P, Q (and Bu, Bi for kind 2 = SVD); kind 0 BasicMF, 1 PMF; returns sum(error^2)."""
    u = np.ascontiguousarray(u, np.int32); i = np.ascontiguousarray(i, np.int32)
    assert P.flags.c_contiguous and Q.flags.c_contiguous and P.dtype == Q.dtype
    ct, fn = ((C.c_double, lib().oracle_mf_sgd_sequential_f64) if P.dtype == np.float64
              else (C.c_float, lib().oracle_mf_sgd_sequential_f32))
    r = np.ascontiguousarray(r, P.dtype)
    if kind == 2:
        assert Bu.dtype == P.dtype and Bi.dtype == P.dtype and Bu.flags.c_contiguous and Bi.flags.c_contiguous
    bu = _p(Bu, ct) if kind == 2 else None
    bi = _p(Bi, ct) if kind == 2 else None
    return fn(kind, _p(P, ct), _p(Q, ct), P.shape[1], len(u), _p(u, C.c_int32), _p(i, C.c_int32), _p(r, ct),
              lr, reg_u, reg_i, bu, bi, reg_b, global_mean)
