"""Host-side mirror of the reference crate's public types for the batched path.

Names, defaults and meanings follow the Rust crate `ivp` v0.5.1 so a user of the reference
finds the same vocabulary:

* `Method`      reference src/solve/options.rs:14-73
* `Status`      reference src/status.rs:4-26
* `Direction`, `EventConfig`   reference src/solve/event.rs:5-77
* `Options` (+ `Options.builder()`)   reference src/solve/options.rs:75-123
* `Solution`    reference src/solve/solution.rs:7-20

The Rust toolchain is absent from this image, so this Python layer (over ctypes -> the C ABI in
include/ivpb.h) is the executable host mirror; rust/ holds the equivalent crate sources.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


class Method(enum.IntEnum):
    """options.rs:14-27 (declaration order == C ABI code)."""
    RK23 = 0
    DOPRI5 = 1
    DOP853 = 2
    RK4 = 3
    RADAU = 4
    BDF = 5

    @classmethod
    def from_str(cls, s: str) -> "Method":
        """`impl From<&str> for Method` (options.rs:61-73): unknown names fall back to DOPRI5."""
        return {
            "RK23": cls.RK23, "DOPRI5": cls.DOPRI5, "RK45": cls.DOPRI5, "DOP853": cls.DOP853, "RK4": cls.RK4,
            "RADAU": cls.RADAU, "RADAU5": cls.RADAU, "BDF": cls.BDF, "BDF15": cls.BDF,
        }.get(s.upper(), cls.DOPRI5)

    def coeffs_per_state(self) -> int:
        """options.rs:34-43."""
        return {Method.RK4: 4, Method.RK23: 4, Method.DOPRI5: 5, Method.DOP853: 8, Method.RADAU: 4, Method.BDF: 7}[self]


class Status(enum.IntEnum):
    """status.rs:4-19."""
    Success = 0
    UserInterrupt = 1
    NeedLargerNMax = 2
    StepSizeTooSmall = 3
    ProbablyStiff = 4
    SingularMatrix = 5
    PoorConvergence = 6

    def is_success(self) -> bool:
        """status.rs:22-25."""
        return self in (Status.Success, Status.UserInterrupt)


class Direction(enum.IntEnum):
    """event.rs:55-77 (`From<i32>`: >0 Positive, <0 Negative, 0 All)."""
    All = 0
    Positive = 1
    Negative = -1


@dataclass
class EventConfig:
    """event.rs:5-53."""
    direction: Direction = Direction.All
    terminal_count: Optional[int] = None

    def terminal(self):
        self.terminal_count = 1
        return self

    def all(self):
        self.direction = Direction.All
        return self

    def positive(self):
        self.direction = Direction.Positive
        return self

    def negative(self):
        self.direction = Direction.Negative
        return self


class ConfigError(ValueError):
    """`Error::Config` (reference src/error.rs:18-60): invalid solver parameters, detected before stepping."""


class InterpolationError(RuntimeError):
    """`Error::Interpolation` (src/error.rs): dense output disabled or t outside the covered span."""


@dataclass
class Options:
    """`Options` (options.rs:75-123).  Field names and defaults are the reference's; the trailing block
    holds what the reference takes from the `IVP` trait or does not need for one trajectory."""
    method: Method = Method.DOPRI5
    rtol: object = 1e-3          # scalar or length-n sequence (`Tolerance`)
    atol: object = 1e-6
    max_steps: Optional[int] = None
    t_eval: Optional[Sequence[float]] = None
    first_step: Optional[float] = None
    max_step: Optional[float] = None
    min_step: Optional[float] = None
    dense_output: bool = False
    jac_storage: str = "Full"        # accepted for API parity; device Jacobians are dense
    mass_storage: str = "Identity"
    nind1: Optional[int] = None
    nind2: Optional[int] = None
    nind3: Optional[int] = None
    # --- batch-ABI extras (include/ivpb.h) ---
    event_config: Optional[List[EventConfig]] = None   # None => the problem's IVP::event_config defaults
    max_events: int = 8          # capacity of t_events/y_events per event function per trajectory
    max_out: int = 0             # step-mode capacity of Solution.t/.y per trajectory (0 = final state only)
    jac_mode: int = 0            # 0 finite differences (ivp.rs:67-107), 1 analytic
    jac_sparsity: object = None  # (n, n) structure of the Jacobian: finite differences per column GROUP (src/python/sparsity.rs)
    flags: int = 0
    user_solout: bool = False    # the problem's own SolOut hook replaces DefaultSolOut (src/solout.rs:55-63; include/ivpb.h)
    max_segments: int = 0        # dense_output: interpolant segments kept per trajectory (one per accepted step)

    def __post_init__(self):
        if isinstance(self.method, str):
            self.method = Method.from_str(self.method)
        self.method = Method(self.method)

    @staticmethod
    def builder() -> "_OptionsBuilder":
        """`Options::builder()` (bon-generated in the reference)."""
        return _OptionsBuilder()


class _OptionsBuilder:
    def __init__(self):
        self._kw = {}

    def __getattr__(self, name):
        if name.startswith("_") or name not in Options.__dataclass_fields__:
            raise AttributeError(name)

        def setter(value):
            self._kw[name] = value
            return self
        return setter

    def build(self) -> Options:
        return Options(**self._kw)


@dataclass
class Solution:
    """`Solution` (solution.rs:7-20) for one trajectory of a batch."""
    t: np.ndarray
    y: np.ndarray                 # [len(t), n]
    t_events: List[np.ndarray]
    y_events: List[np.ndarray]    # per event: [k, n]
    nfev: int
    njev: int
    nlu: int
    nstep: int
    naccpt: int
    nrejct: int
    status: Status
    truncated: bool = False       # batch-only: more samples/events than the preallocated capacity

    def iter(self):
        """solution.rs:73-78."""
        return zip(self.t, self.y)


@dataclass
class BatchSolution:
    """Struct-of-arrays result of `solve_ivp_batch`; `solution(i)` / `solutions()` give per-trajectory
    `Solution`s.  Array layout == include/ivpb.h `ivpb_outputs`."""
    n: int
    n_events: int
    status: np.ndarray
    counters: np.ndarray
    t_final: np.ndarray
    y_final: np.ndarray
    h_next: Optional[np.ndarray] = None
    n_out: Optional[np.ndarray] = None
    t_out: Optional[np.ndarray] = None
    y_out: Optional[np.ndarray] = None
    ev_count: Optional[np.ndarray] = None
    ev_t: Optional[np.ndarray] = None
    ev_y: Optional[np.ndarray] = None
    n_seg: Optional[np.ndarray] = None      # dense_output segment copies (only when asked for via `want`)
    seg_x: Optional[np.ndarray] = None
    seg_cont: Optional[np.ndarray] = None
    extras: dict = field(default_factory=dict)

    # ---- Solution::sol / sol_many / sol_span (solution.rs:25-72), evaluated on the device ----
    def sol_many(self, traj, ts, extrapolate: bool = False):
        """Dense output of trajectory `traj[q]` at `ts[q]`; returns (y[Q, n], ok[Q]).  `extrapolate`: the rule of
        ContinuousOutput::evaluate_extrapolate (cont.rs:91-150) for times outside the stored steps."""
        return self.extras["ctx"].dense_eval(np.asarray(traj), np.asarray(ts), self.n, extrapolate,
                                             generation=self.extras.get("dense_generation", 0))

    def sol(self, i: int, t: float):
        """Solution::sol for trajectory i: InterpolationError semantics of solution.rs:25-44."""
        span = self.sol_span(i)
        if span is None:
            raise InterpolationError("dense output not enabled")
        lo, hi = min(span), max(span)
        if t < lo or t > hi:
            raise InterpolationError(f"t={t} outside the dense output span {span}")
        y, ok = self.sol_many([i], [t])
        if not ok[0]:
            raise InterpolationError(f"t={t} outside the dense output span {span}")
        return y[0]

    def sol_span(self, i: int):
        if "ctx" not in self.extras or not self.extras.get("dense"):
            return None
        t0, t1, m = self.extras["ctx"].dense_span(i, 1, generation=self.extras.get("dense_generation", 0))
        return (float(t0[0]), float(t1[0])) if m[0] > 0 else None

    def __len__(self):
        return int(self.status.shape[0])

    @property
    def nfev(self): return self.counters[:, 0]
    @property
    def njev(self): return self.counters[:, 1]
    @property
    def nlu(self): return self.counters[:, 2]
    @property
    def nstep(self): return self.counters[:, 3]
    @property
    def naccpt(self): return self.counters[:, 4]
    @property
    def nrejct(self): return self.counters[:, 5]

    def solution(self, i: int) -> Solution:
        trunc = False
        if self.t_out is not None and self.n_out is not None:
            cap = self.t_out.shape[1]
            m = int(self.n_out[i])
            trunc |= m > cap
            m = min(m, cap)
            t = self.t_out[i, :m].copy()
            y = self.y_out[i, :m].copy() if self.y_out is not None else np.zeros((m, self.n))
        else:
            t = np.array([self.t_final[i]])
            y = self.y_final[i][None, :].copy()
        te, ye = [], []
        for e in range(self.n_events):
            k = int(self.ev_count[i, e]) if self.ev_count is not None else 0
            cap = self.ev_t.shape[2] if self.ev_t is not None else 0
            trunc |= k > cap
            k = min(k, cap)
            te.append(self.ev_t[i, e, :k].copy() if self.ev_t is not None else np.zeros(0))
            ye.append(self.ev_y[i, e, :k].copy() if self.ev_y is not None else np.zeros((0, self.n)))
        c = self.counters[i]
        return Solution(t=t, y=y, t_events=te, y_events=ye, nfev=int(c[0]), njev=int(c[1]), nlu=int(c[2]),
                        nstep=int(c[3]), naccpt=int(c[4]), nrejct=int(c[5]), status=Status(int(self.status[i])),
                        truncated=trunc)

    def solutions(self):
        return [self.solution(i) for i in range(len(self))]
