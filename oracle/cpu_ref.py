"""ORACLE (test infrastructure only): ctypes wrapper of oracle/libcpu_ref.so (cpu_ref.cpp)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libcpu_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libcpu_ref.so is not built: run `make -C oracle`")
        L = ctypes.CDLL(path)
        dp = ctypes.POINTER(ctypes.c_double)
        L.cpu_ref_eval.restype = ctypes.c_double
        L.cpu_ref_eval.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int, dp, dp]
        L.cpu_ref_hmc.restype = ctypes.c_int64
        L.cpu_ref_hmc.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_double, ctypes.c_uint64, dp, dp, ctypes.c_int]
        L.cpu_ref_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def evaluate(model, X, Y, beta, gradient=True):
    X = np.asfortranarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    g = np.zeros_like(beta) if gradient else None
    lp = lib().cpu_ref_eval(0 if model == "linear" else 1, _p(X), _p(Y), X.shape[0], X.shape[1], _p(beta), _p(g))
    return (lp, g) if gradient else lp


def hmc(model, X, Y, init, steps, isteps, stepsize, seed=1, threads=0):
    X = np.asfortranarray(X, dtype=np.float64)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    init = np.ascontiguousarray(np.atleast_2d(init), dtype=np.float64)
    chains, M = init.shape
    final = np.empty((chains, M))
    acc = np.empty(chains)
    n = lib().cpu_ref_hmc(0 if model == "linear" else 1, _p(X), _p(Y), X.shape[0], M, _p(init), chains, steps, isteps,
                          stepsize, seed, _p(final), _p(acc), threads)
    return dict(n_leapfrogs=int(n), final=final, accept_rate=acc)


def max_threads():
    return int(lib().cpu_ref_max_threads())
