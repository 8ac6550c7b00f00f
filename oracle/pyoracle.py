"""ctypes wrapper of liboracle.so (the CPU restatement of the reference).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs --
never by anything under ivp_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ivp_b200 import _abi
from ivp_b200.types import BatchSolution, Options

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so)
            for f in ("capi.cpp", "ivp_oracle.hpp", "ivp_oracle_implicit.hpp", "problems.hpp", "test_problems.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_solve_batch.restype = C.c_int
        L.oracle_solve_batch.argtypes = [C.c_int, C.POINTER(_abi.IvpbOptions), C.c_int64, C.c_double, C.c_double,
                                         _abi.c_double_p, _abi.c_double_p, C.POINTER(_abi.IvpbOutputs), C.c_int]
        L.oracle_dense_eval.restype = C.c_int
        L.oracle_dense_eval.argtypes = [C.c_int, C.POINTER(_abi.IvpbOptions), C.c_double, C.c_double, _abi.c_double_p,
                                        _abi.c_double_p, _abi.c_double_p, C.c_int, _abi.c_double_p, _abi.c_int32_p,
                                        _abi.c_double_p]
        L.oracle_dense_eval_extrapolate.restype = C.c_int
        L.oracle_dense_eval_extrapolate.argtypes = L.oracle_dense_eval.argtypes
        L.oracle_problem_dims.argtypes = [C.c_int, _abi.c_int32_p, _abi.c_int32_p, _abi.c_int32_p]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_hardware_threads.restype = C.c_int
        _LIB = L
    return _LIB


def dims(problem: int):
    n, p, ne = C.c_int32(), C.c_int32(), C.c_int32()
    if lib().oracle_problem_dims(int(problem), C.byref(n), C.byref(p), C.byref(ne)):
        raise ValueError(f"unknown problem {problem}")
    return n.value, p.value, ne.value


def hardware_threads() -> int:
    return lib().oracle_hardware_threads()


def solve_batch(problem: int, t0: float, tf: float, y0, params, options: Options, nthreads: int = 1,
                want=None) -> BatchSolution:
    n, p, ne = dims(problem)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(-1, n))
    N = y0.shape[0]
    par = None
    if p > 0:
        par = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(N, p))
    mo = _abi.MarshalledOptions(options, n, ne)
    n_cont = options.method.coeffs_per_state() * n
    arrays, st = _abi.alloc_outputs(N, n, ne, mo.cap, int(options.max_events), want, seg_cap=mo.seg_cap, n_cont=n_cont)
    rc = lib().oracle_solve_batch(int(problem), C.byref(mo.struct), N, float(t0), float(tf), _abi.ptr(y0),
                                  _abi.ptr(par), C.byref(st), int(nthreads))
    if rc:
        raise RuntimeError(lib().oracle_last_error().decode())
    return BatchSolution(n=n, n_events=ne, **{k: arrays.get(k) for k in _abi.OUTPUT_FIELDS})


def dense_eval(problem: int, t0, tf, y0, params, options: Options, ts, extrapolate: bool = False):
    """Solve one trajectory with dense_output=true; return (ys[len(ts), n], ok[len(ts)], span|None).
    `extrapolate`: ContinuousOutput::evaluate_extrapolate instead of Solution::sol."""
    n, p, ne = dims(problem)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(n))
    par = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(p)) if p > 0 else None
    ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1))
    ys = np.zeros((ts.size, n))
    ok = np.zeros(ts.size, dtype=np.int32)
    span = np.zeros(3)
    mo = _abi.MarshalledOptions(options, n, ne)
    fn = lib().oracle_dense_eval_extrapolate if extrapolate else lib().oracle_dense_eval
    rc = fn(int(problem), C.byref(mo.struct), float(t0), float(tf), _abi.ptr(y0), _abi.ptr(par),
                                 _abi.ptr(ts), ts.size, _abi.ptr(ys), _abi.ptr(ok), _abi.ptr(span))
    if rc:
        raise RuntimeError(lib().oracle_last_error().decode())
    return ys, ok.astype(bool), ((span[0], span[1]) if span[2] else None)
