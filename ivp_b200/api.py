"""Python host binding of libivpb.so (include/ivpb.h) and the batched form of the reference's
`solve_ivp` (reference src/solve/solve_ivp.rs:99-313).

    sol = solve_ivp_batch(problem, t0, tf, Y0[N, n], params[N, p], options)   # -> BatchSolution
    sol.solution(i)                                                           # -> reference-shaped Solution

There is no CPU fallback: the first compute call loads ivp_b200/lib/libivpb.so and creates a CUDA
context; both fail loudly (ImportError / RuntimeError) when the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .types import BatchSolution, ConfigError, Method, Options, Solution, Status  # noqa: F401

_LIB_PATH = os.environ.get("IVPB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libivpb.so")
_lib = None

#: Built-in problems (include/ivpb.h `ivpb_builtin`).
PROBLEMS = {
    "decay": 0, "vdp_eps": 1, "vdp_mu": 2, "lorenz": 3, "cr3bp": 4, "bouncing_ball": 5, "robertson": 6,
    "sho": 7, "zero3": 8, "exp2": 9, "rational": 10, "cannon": 11, "linear100": 12, "medakzo64": 13, "robertson_dae": 14, "mass_linear3": 15,
    "ball_bounce": 16,
}

IVPB_FLAG_STRICT_FP = 1
IVPB_FLAG_NO_REFILL = 2
IVPB_FLAG_NO_ZEROCOPY = 4
IVPB_FLAG_FAST_FP = 8
IVPB_FLAG_NO_SORT = 16
IVPB_FLAG_SORT = 32
IVPB_FLAG_ZEROCOPY_OUT = 64
IVPB_FLAG_NO_PIPELINE = 128

#: Every symbol include/ivpb.h declares (checked by the CPU test-suite against the built library).
ABI_SYMBOLS = [
    "ivpb_create", "ivpb_destroy", "ivpb_last_error", "ivpb_device_count", "ivpb_builtin_problem",
    "ivpb_nvrtc_problem", "ivpb_solve_batch", "ivpb_solve_batch_device", "ivpb_host_alloc", "ivpb_host_free",
    "ivpb_launch_count", "ivpb_measure_fp64_peak", "ivpb_version", "ivpb_dense_eval", "ivpb_dense_span",
    "ivpb_dense_eval_extrapolate", "ivpb_dense_generation", "ivpb_last_fp_mode",
]


def load_library():
    """dlopen libivpb.so and declare the C signatures.  No CUDA call is made here."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C ivp_b200/csrc). ivp_b200 has no CPU fallback.")
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.ivpb_create.restype = C.c_int
    L.ivpb_create.argtypes = [C.POINTER(vp), _abi.c_int32_p, C.c_int]
    L.ivpb_destroy.restype = None
    L.ivpb_destroy.argtypes = [vp]
    L.ivpb_last_error.restype = C.c_char_p
    L.ivpb_last_error.argtypes = [vp]
    L.ivpb_device_count.restype = C.c_int
    L.ivpb_device_count.argtypes = [vp]
    L.ivpb_builtin_problem.restype = C.c_int
    L.ivpb_builtin_problem.argtypes = [vp, C.c_int, _abi.c_int32_p, _abi.c_int32_p, _abi.c_int32_p]
    L.ivpb_nvrtc_problem.restype = C.c_int
    L.ivpb_nvrtc_problem.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, _abi.c_int32_p]
    L.ivpb_solve_batch.restype = C.c_int
    L.ivpb_solve_batch.argtypes = [vp, C.c_int, C.POINTER(_abi.IvpbOptions), C.c_int64, C.c_double, C.c_double,
                                   vp, vp, C.POINTER(_abi.IvpbOutputs)]
    L.ivpb_solve_batch_device.restype = C.c_int
    L.ivpb_solve_batch_device.argtypes = [vp, C.c_int, C.POINTER(_abi.IvpbOptions), C.c_int64, C.c_double, C.c_double,
                                          vp, vp, C.POINTER(_abi.IvpbOutputs), vp]
    L.ivpb_host_alloc.restype = vp
    L.ivpb_host_alloc.argtypes = [C.c_size_t]
    L.ivpb_host_free.restype = None
    L.ivpb_host_free.argtypes = [vp]
    L.ivpb_launch_count.restype = C.c_uint64
    L.ivpb_launch_count.argtypes = [vp]
    L.ivpb_measure_fp64_peak.restype = C.c_int
    L.ivpb_measure_fp64_peak.argtypes = [vp, _abi.c_double_p]
    L.ivpb_last_fp_mode.restype = C.c_int
    L.ivpb_last_fp_mode.argtypes = [vp, _abi.c_int32_p]
    L.ivpb_version.restype = C.c_char_p
    L.ivpb_dense_eval.restype = C.c_int
    L.ivpb_dense_eval.argtypes = [vp, C.c_uint64, C.c_int, C.c_int64, _abi.c_int64_p, _abi.c_double_p, _abi.c_double_p, _abi.c_int32_p]
    L.ivpb_dense_generation.restype = C.c_uint64
    L.ivpb_dense_generation.argtypes = [vp]
    L.ivpb_dense_eval_extrapolate.restype = C.c_int
    L.ivpb_dense_eval_extrapolate.argtypes = L.ivpb_dense_eval.argtypes
    L.ivpb_dense_span.restype = C.c_int
    L.ivpb_dense_span.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, _abi.c_double_p, _abi.c_double_p, _abi.c_int32_p]
    _lib = L
    return L


class Problem:
    """Device form of the reference's `IVP` trait (src/ivp.rs:27-121): either a built-in problem whose
    ode/events/jac are compiled with the solver, or user CUDA C compiled through NVRTC."""

    def __init__(self, handle: int, n: int, p: int, n_events: int, name: str = "", cuda_src: Optional[str] = None,
                 has_jac: bool = False):
        self.handle, self.n, self.p, self.n_events, self.name = handle, n, p, n_events, name
        self.cuda_src, self.has_jac = cuda_src, has_jac
        # NVRTC handle per context, keyed by the Context object itself (weakly): an id() would be reused by CPython once a
        # context has been collected, handing a new context a stale handle
        self._ctx_handles = weakref.WeakKeyDictionary()

    @staticmethod
    def builtin(name_or_id) -> "Problem":
        pid = PROBLEMS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        L = load_library()
        n, p, ne = C.c_int32(), C.c_int32(), C.c_int32()
        if L.ivpb_builtin_problem(None, pid, C.byref(n), C.byref(p), C.byref(ne)):
            raise ConfigError(f"unknown built-in problem {name_or_id!r}")
        name = next((k for k, v in PROBLEMS.items() if v == pid), str(pid))
        return Problem(pid, n.value, p.value, ne.value, name)

    @staticmethod
    def from_cuda_source(src: str, n: int, p: int = 0, n_events: int = 0, has_jac: bool = False,
                         has_mass: bool = False, has_solout: bool = False) -> "Problem":
        """User problem: `src` defines `__device__ void ivp_ode(double t, const double* y, const double* p,
        double* dydt)` (+ `ivp_events`, `ivp_jac`, `ivp_mass`, `ivp_solout`), see include/ivpb.h."""
        return Problem(-1, n, p, n_events, "user", cuda_src=src, has_jac=int(bool(has_jac)) | (2 if has_mass else 0) | (4 if has_solout else 0))

    def resolve(self, ctx: "Context") -> int:
        if self.cuda_src is None and self.n > 0:
            return self.handle
        key = ctx
        if key not in self._ctx_handles:
            h = C.c_int32()
            rc = ctx.lib.ivpb_nvrtc_problem(ctx.ptr, self.cuda_src.encode() if self.cuda_src else None, self.n, self.p, self.n_events,
                                            int(self.has_jac), C.byref(h))
            ctx.check(rc)
            self._ctx_handles[key] = h.value
        return self._ctx_handles[key]


    @staticmethod
    def empty(n_events: int = 0) -> "Problem":
        """The reference's empty state vector (`y0.is_empty()`, src/solve/solve_ivp.rs:148-176): nothing to integrate."""
        return Problem(-1, 0, 0, n_events, "empty", cuda_src=None)


def builtin(name_or_id) -> Problem:
    return Problem.builtin(name_or_id)


class Context:
    """`ivpb_ctx`: the device set + streams + work queues.  `devices=None` => the current CUDA device."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        p = C.c_void_p()
        if devices is None:
            rc = self.lib.ivpb_create(C.byref(p), None, 0)
        else:
            ids = np.asarray(list(devices), dtype=np.int32)
            rc = self.lib.ivpb_create(C.byref(p), _abi.ptr(ids), len(ids))
        if rc:
            raise RuntimeError("ivpb_create failed: " + self.lib.ivpb_last_error(None).decode())
        self.ptr = p

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.ivpb_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc == 0:
            return
        msg = self.lib.ivpb_last_error(self.ptr).decode()
        if rc == 1:
            raise ConfigError(msg)
        raise RuntimeError(f"libivpb error {rc}: {msg}")

    @property
    def n_devices(self) -> int:
        return self.lib.ivpb_device_count(self.ptr)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ivpb_launch_count(self.ptr))

    def last_fp_mode(self) -> dict:
        """Which kernel build the most recent solve ran and why (ivpb_last_fp_mode)."""
        info = (C.c_int32 * 6)()
        rc = self.lib.ivpb_last_fp_mode(self.ptr, info)
        src = {0: "flag", 1: "method default", 2: "parity pilot (cached)", 3: "parity pilot"}.get(int(info[1]), "?")
        d = {"mode": "strict" if rc == 1 else "fma", "source": src}
        if int(info[1]) >= 2:
            d.update(sample=int(info[2]), status_mismatches=int(info[3]), step_count_mismatches=int(info[4]),
                     out_of_tolerance=int(info[5]))
        return d

    def last_reruns(self) -> int:
        """Debug: trajectories the last strict solve on device 0 handed to its guarded second pass (-1: none has run)."""
        self.lib.ivpb_debug_last_reruns.restype = C.c_longlong
        self.lib.ivpb_debug_last_reruns.argtypes = [C.c_void_p]
        return int(self.lib.ivpb_debug_last_reruns(self.ptr))

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        self.check(self.lib.ivpb_measure_fp64_peak(self.ptr, C.byref(v)))
        return v.value

    # -- dense output (Solution::sol on the device) ------------------------------------------------
    def dense_generation(self) -> int:
        """Identity of the retained dense log (0 = none); BatchSolution remembers it and every dense call checks it."""
        return int(self.lib.ivpb_dense_generation(self.ptr))

    def dense_eval(self, traj: np.ndarray, ts: np.ndarray, n: int, extrapolate: bool = False, generation: int = 0):
        """ivpb_dense_eval (or ivpb_dense_eval_extrapolate: ContinuousOutput::evaluate_extrapolate, cont.rs:91-150) on
        the log retained by the last dense_output solve: (y[Q, n], ok[Q])."""
        traj = np.ascontiguousarray(np.asarray(traj, dtype=np.int64).reshape(-1))
        ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1))
        if traj.size != ts.size:
            raise ValueError("traj and ts must have the same length")
        y = np.zeros((ts.size, n))
        ok = np.zeros(ts.size, dtype=np.int32)
        fn = self.lib.ivpb_dense_eval_extrapolate if extrapolate else self.lib.ivpb_dense_eval
        self.check(fn(self.ptr, int(generation), int(n), ts.size, _abi.ptr(traj), _abi.ptr(ts), _abi.ptr(y), _abi.ptr(ok)))
        return y, ok.astype(bool)

    def dense_span(self, first: int, count: int, generation: int = 0):
        t0, t1 = np.zeros(count), np.zeros(count)
        m = np.zeros(count, dtype=np.int32)
        self.check(self.lib.ivpb_dense_span(self.ptr, int(generation), int(first), int(count), _abi.ptr(t0), _abi.ptr(t1), _abi.ptr(m)))
        return t0, t1, m

    # -- host buffers ---------------------------------------------------------------------------
    def solve_host(self, problem: Problem, t0: float, tf: float, y0: np.ndarray, params: Optional[np.ndarray],
                   mo: _abi.MarshalledOptions, out_struct: _abi.IvpbOutputs):
        """Raw ivpb_solve_batch on caller-owned (ideally pinned) numpy buffers."""
        N = y0.shape[0]
        rc = self.lib.ivpb_solve_batch(self.ptr, problem.resolve(self), C.byref(mo.struct), N, float(t0), float(tf),
                                       y0.ctypes.data, params.ctypes.data if params is not None else None,
                                       C.byref(out_struct))
        self.check(rc)

    # -- device buffers -------------------------------------------------------------------------
    def solve_device(self, problem: Problem, t0: float, tf: float, N: int, d_y0: int, d_params: Optional[int],
                     mo: _abi.MarshalledOptions, d_out: dict, stream: int = 0):
        """ivpb_solve_batch_device: all pointers are raw device addresses (e.g. torch `.data_ptr()`)."""
        st = _abi.IvpbOutputs()
        for f in _abi.OUTPUT_FIELDS:
            v = d_out.get(f)
            if v:
                setattr(st, f, C.cast(C.c_void_p(int(v)), dict(_abi.IvpbOutputs._fields_)[f]))
        rc = self.lib.ivpb_solve_batch_device(self.ptr, problem.resolve(self), C.byref(mo.struct), int(N), float(t0),
                                              float(tf), C.c_void_p(int(d_y0)),
                                              C.c_void_p(int(d_params)) if d_params else None, C.byref(st),
                                              C.c_void_p(int(stream)) if stream else None)
        self.check(rc)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def solve_ivp_batch(problem, t0: float, tf: float, y0, params=None, options: Optional[Options] = None,
                    ctx: Optional[Context] = None, want: Optional[Sequence[str]] = None) -> BatchSolution:
    """Batched `solve_ivp`: trajectory i starts at `y0[i]` with parameter row `params[i]`.

    Follows reference src/solve/solve_ivp.rs:99-313 per trajectory (same options, same statuses,
    same counters); configuration errors raise `ConfigError` before any stepping, numerical failures
    are reported in `BatchSolution.status`.
    """
    if not isinstance(problem, Problem):
        problem = Problem.builtin(problem)
    options = options or Options()
    ctx = ctx or default_context()
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    if y0.ndim == 1:
        y0 = y0.reshape(1, -1)
    if y0.ndim != 2 or y0.shape[1] != problem.n:
        raise ConfigError(f"y0 must have shape [N, {problem.n}], got {y0.shape}")
    N = y0.shape[0]
    par = None
    if problem.p > 0:
        if params is None:
            raise ConfigError(f"problem {problem.name} needs params of shape [N, {problem.p}]")
        par = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        if par.ndim == 1:
            par = np.broadcast_to(par.reshape(1, -1), (N, problem.p)).copy()
        if par.shape != (N, problem.p):
            raise ConfigError(f"params must have shape [{N}, {problem.p}], got {par.shape}")
    mo = _abi.MarshalledOptions(options, problem.n, problem.n_events)
    n_cont = Method(options.method).coeffs_per_state() * problem.n
    arrays, st = _abi.alloc_outputs(N, problem.n, problem.n_events, mo.cap, int(options.max_events), want,
                                    seg_cap=mo.seg_cap, n_cont=n_cont)
    if N > 0:
        ctx.solve_host(problem, t0, tf, y0, par, mo, st)
    dense = mo.seg_cap > 0 and N > 0
    return BatchSolution(n=problem.n, n_events=problem.n_events,
                         extras={"ctx": ctx, "dense": dense, "dense_generation": ctx.dense_generation() if dense else 0},
                         **{k: arrays.get(k) for k in _abi.OUTPUT_FIELDS})


def solve_ivp(problem, x0: float, xend: float, y0, options: Optional[Options] = None, params=None,
              ctx: Optional[Context] = None) -> Solution:
    """Single-trajectory convenience with the reference's argument order (solve_ivp.rs:99-105)."""
    opts = options or Options()
    import dataclasses
    if opts.t_eval is None and opts.max_out == 0:
        opts = dataclasses.replace(opts, max_out=4096)
    if opts.dense_output and opts.max_segments == 0:
        opts = dataclasses.replace(opts, max_segments=4096)
    b = solve_ivp_batch(problem, x0, xend, np.asarray(y0, dtype=np.float64).reshape(1, -1),
                        None if params is None else np.asarray(params, dtype=np.float64).reshape(1, -1), opts, ctx)
    return b.solution(0)
