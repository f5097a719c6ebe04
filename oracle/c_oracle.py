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
