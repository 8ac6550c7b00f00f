"""SciPy-style front end: the signature and result object of the reference's Python binding
(`solve_ivp(fun, t_span, y0, method, t_eval, dense_output, events, vectorized, args, jac, jac_sparsity, **options)
-> OdeResult`, reference src/python/solve.rs:153-432, src/python/result.rs:15, src/python/solution.rs:17),
re-targeted at the batched device ABI.

Python callables cannot run on the GPU, so `fun` is a device problem instead of a function: the name of a
built-in problem ("vdp_mu", "robertson", ...), an `ivp_b200.Problem`, or CUDA C source defining `ivp_ode`
(+ `ivp_events`, `ivp_jac`) that NVRTC compiles with the solver.  `args` becomes the parameter row.  Everything
else -- method names incl. "RK45" / "Radau", options, `y` returned as (n, n_points), status 0 / 1 / -1, `sol` --
follows the reference binding.  `y0` may be 2-D ([N, n]): one OdeResult per row comes back (the batched call).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

from . import api
from .types import Direction, EventConfig, Method, Options, Status


_SOURCE_PROBLEMS: dict = {}


class OdeSolution:
    """`sol(t)`: dense output of one trajectory, evaluated on the device (src/python/solution.rs:17)."""

    def __init__(self, batch, index: int, n: int):
        self._b, self._i, self._n = batch, index, n
        span = batch.sol_span(index)
        self.t_min, self.t_max = (min(span), max(span)) if span else (np.nan, np.nan)

    def __call__(self, t):
        ts = np.atleast_1d(np.asarray(t, dtype=np.float64))
        # the reference's OdeSolution.__call__ extrapolates from the first / last step (src/python/solution.rs:41,116)
        y, ok = self._b.sol_many(np.full(ts.size, self._i), ts, extrapolate=True)
        if not ok.all():
            raise ValueError("t is outside the solution range")          # solution.rs:47-50,119-121
        return y[0] if np.ndim(t) == 0 else y.T


@dataclasses.dataclass
class OdeResult:
    """reference src/python/result.rs:15-60"""
    t: np.ndarray
    y: np.ndarray                 # (n, n_points), like SciPy
    t_events: Optional[List[np.ndarray]]
    y_events: Optional[List[np.ndarray]]
    nfev: int
    njev: int
    nlu: int
    status: int                   # -1 failed, 0 reached the end of t_span, 1 terminated by an event
    message: str
    success: bool
    sol: Optional[OdeSolution] = None


def _event_configs(events, n_events: int):
    """`events`: None, one spec or a sequence; each spec carries SciPy's `terminal` / `direction` attributes
    (solve.rs:245-287).  The event FUNCTIONS are the problem's own `ivp_events`."""
    if events is None:
        # no event handling requested: keep the compiled-in event functions inert (non-terminal, All)
        return [EventConfig(Direction.All, None)] * n_events if n_events else None, False
    specs = list(events) if isinstance(events, (list, tuple)) else [events]
    if len(specs) != n_events:
        raise ValueError(f"{len(specs)} event specs given, the problem defines {n_events} event functions")
    cfgs = []
    for s in specs:
        term = getattr(s, "terminal", False)
        d = float(getattr(s, "direction", 0.0))
        count = None if not term else (1 if term is True else int(term))
        cfgs.append(EventConfig(Direction.Positive if d > 0 else Direction.Negative if d < 0 else Direction.All, count))
    return cfgs, True


def solve_ivp(fun, t_span, y0, method=None, t_eval=None, dense_output=False, events=None, vectorized=False,
              args=None, jac=None, jac_sparsity=None, **options):
    del vectorized, jac_sparsity          # accepted for signature compatibility (solve.rs:169)
    if callable(fun) and not isinstance(fun, api.Problem):
        raise TypeError("ivp_b200 integrates on the GPU: pass a built-in problem name, an ivp_b200.Problem, or CUDA C "
                        "source defining ivp_ode -- a Python callable cannot run on the device")
    y0 = np.asarray(y0, dtype=np.float64)
    single = y0.ndim == 1
    Y0 = np.atleast_2d(y0)
    if isinstance(fun, api.Problem):
        problem = fun
    elif isinstance(fun, str) and "ivp_ode" in fun:
        p = 0 if args is None else len(np.atleast_1d(args))
        # event functions live in the source (`ivp_events` fills g[0..n_events)): their number is the number of
        # event specs given, or `n_events=` when no specs are passed
        ne = int(options.pop("n_events", 0)) or (len(events) if isinstance(events, (list, tuple)) else int(events is not None))
        ne = ne if "ivp_events" in fun else 0
        key = (fun, Y0.shape[1], p, ne)
        if key not in _SOURCE_PROBLEMS:          # one NVRTC module cache per distinct source: repeated calls reuse it
            _SOURCE_PROBLEMS[key] = api.Problem.from_cuda_source(fun, n=Y0.shape[1], p=p, n_events=ne, has_jac="ivp_jac" in fun,
                                                                   has_mass="ivp_mass" in fun)
        problem = _SOURCE_PROBLEMS[key]
    else:
        problem = api.Problem.builtin(fun)
    t0, tf = float(t_span[0]), float(t_span[1])
    known = {"rtol", "atol", "max_step", "min_step", "first_step", "max_steps",          # solve.rs:290-340
             "mass_storage", "nind1", "nind2", "nind3",                                  # Options of the Rust crate (RADAU)
             "jac_sparsity"}                                                             # solve.rs: sparse finite differences
    unknown = set(options) - known - {"max_events", "max_out", "max_segments", "strict_fp", "fast_fp"}
    if unknown:
        raise TypeError(f"unknown options: {sorted(unknown)}")
    cfgs, has_events = _event_configs(events, problem.n_events)
    m = Method.from_str(method) if isinstance(method, str) else (Method.DOPRI5 if method is None else Method(method))
    opts = Options(method=m, rtol=options.get("rtol", 1e-3), atol=options.get("atol", 1e-6),
                   max_steps=options.get("max_steps"), t_eval=None if t_eval is None else np.asarray(t_eval, dtype=np.float64),
                   first_step=options.get("first_step"), max_step=options.get("max_step"), min_step=options.get("min_step"),
                   dense_output=bool(dense_output), event_config=cfgs, max_events=int(options.get("max_events", 64)),
                   max_out=int(options.get("max_out", 4096)), max_segments=int(options.get("max_segments", 4096)) if dense_output else 0,
                   mass_storage=options.get("mass_storage", "Identity"), nind1=options.get("nind1"),
                   nind2=options.get("nind2"), nind3=options.get("nind3"), jac_sparsity=options.get("jac_sparsity"),
                   jac_mode=1 if jac in (True, "analytic") or (isinstance(jac, str) and "ivp_jac" in jac) else 0,
                   flags=(api.IVPB_FLAG_STRICT_FP if options.get("strict_fp") else 0) |
                         (api.IVPB_FLAG_FAST_FP if options.get("fast_fp") else 0))
    params = None
    if problem.p > 0:
        if args is None:
            raise ValueError(f"the problem takes {problem.p} parameters: pass them as args=(...)")
        params = np.broadcast_to(np.atleast_2d(np.asarray(args, dtype=np.float64)), (Y0.shape[0], problem.p)).copy()
    b = api.solve_ivp_batch(problem, t0, tf, Y0, params, opts)
    out = []
    for i in range(len(b)):
        s = b.solution(i)
        status = 0 if s.status == Status.Success else (1 if s.status == Status.UserInterrupt else -1)     # solve.rs:404-408
        out.append(OdeResult(t=s.t, y=s.y.T.copy(), t_events=s.t_events if has_events else None,
                             y_events=s.y_events if has_events else None, nfev=s.nfev, njev=s.njev, nlu=s.nlu,
                             status=status, message=s.status.name, success=status >= 0,
                             sol=OdeSolution(b, i, problem.n) if dense_output else None))
    return out[0] if single else out
