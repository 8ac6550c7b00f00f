"""ctypes mirror of include/ivpb.h (struct layouts + helpers to marshal numpy arrays).

Pure marshalling: no numerics here.  The structs are shared by the product binding
(`ivp_b200._lib`) and by the test-side oracle wrapper (`oracle/pyoracle.py`), which
exposes the same output layout so arrays can be compared element for element.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)
c_int64_p = C.POINTER(C.c_int64)


class IvpbOptions(C.Structure):
    _fields_ = [
        ("method", C.c_int32),
        ("n_rtol", C.c_int32),
        ("n_atol", C.c_int32),
        ("rtol", c_double_p),
        ("atol", c_double_p),
        ("has_first_step", C.c_int32),
        ("has_max_step", C.c_int32),
        ("has_min_step", C.c_int32),
        ("has_max_steps", C.c_int32),
        ("first_step", C.c_double),
        ("max_step", C.c_double),
        ("min_step", C.c_double),
        ("max_steps", C.c_uint64),
        ("has_t_eval", C.c_int32),
        ("n_t_eval", C.c_int32),
        ("t_eval", c_double_p),
        ("dense_output", C.c_int32),
        ("n_event_cfg", C.c_int32),
        ("ev_direction", c_int32_p),
        ("ev_terminal_count", c_int64_p),
        ("max_events", C.c_int32),
        ("max_out", C.c_int32),
        ("jac_mode", C.c_int32),
        ("flags", C.c_int32),
        ("max_segments", C.c_int32),
        ("mass_storage", C.c_int32),
        ("user_solout", C.c_int32),
        ("nind1", C.c_int32),
        ("nind2", C.c_int32),
        ("nind3", C.c_int32),
        ("has_jac_sparsity", C.c_int32),
        ("jac_sparsity_colptr", c_int32_p),
        ("jac_sparsity_rows", c_int32_p),
    ]


class IvpbOutputs(C.Structure):
    _fields_ = [
        ("status", c_int32_p),
        ("counters", c_uint32_p),
        ("t_final", c_double_p),
        ("y_final", c_double_p),
        ("h_next", c_double_p),
        ("n_out", c_int32_p),
        ("t_out", c_double_p),
        ("y_out", c_double_p),
        ("ev_count", c_int32_p),
        ("ev_t", c_double_p),
        ("ev_y", c_double_p),
        ("n_seg", c_int32_p),
        ("seg_x", c_double_p),
        ("seg_cont", c_double_p),
    ]


OUTPUT_FIELDS = [f for f, _ in IvpbOutputs._fields_]
_OUT_DTYPES = {
    "status": np.int32, "counters": np.uint32, "t_final": np.float64, "y_final": np.float64,
    "h_next": np.float64, "n_out": np.int32, "t_out": np.float64, "y_out": np.float64,
    "ev_count": np.int32, "ev_t": np.float64, "ev_y": np.float64,
    "n_seg": np.int32, "seg_x": np.float64, "seg_cont": np.float64,
}
_PTR_TYPES = {np.dtype(np.int32): c_int32_p, np.dtype(np.uint32): c_uint32_p,
              np.dtype(np.float64): c_double_p, np.dtype(np.int64): c_int64_p}


def ptr(a: np.ndarray | None):
    """numpy array -> typed ctypes pointer (NULL for None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_PTR_TYPES[a.dtype])


def out_cap(opt_has_t_eval: bool, n_t_eval: int, max_out: int) -> int:
    """Capacity of t_out/y_out rows: include/ivpb.h (`out_cap`)."""
    return n_t_eval + 1 if opt_has_t_eval else max_out


def output_shapes(N: int, n: int, n_events: int, cap: int, max_events: int, seg_cap: int = 0, n_cont: int = 0) -> dict:
    return {
        "n_seg": (N,), "seg_x": (N, seg_cap, 2), "seg_cont": (N, seg_cap, n_cont),
        "status": (N,), "counters": (N, 6), "t_final": (N,), "y_final": (N, n), "h_next": (N,),
        "n_out": (N,), "t_out": (N, cap), "y_out": (N, cap, n),
        "ev_count": (N, n_events), "ev_t": (N, n_events, max_events), "ev_y": (N, n_events, max_events, n),
    }


def alloc_outputs(N, n, n_events, cap, max_events, want=None, seg_cap=0, n_cont=0):
    """Allocate host output arrays (zero-filled) for the requested fields; returns (dict, IvpbOutputs).
    The dense-output segment copies (n_seg, seg_x, seg_cont) are only allocated when named in `want`."""
    shapes = output_shapes(N, n, n_events, cap, max_events, seg_cap, n_cont)
    arrays = {}
    for f in OUTPUT_FIELDS:
        if want is not None and f not in want:
            continue
        shp = shapes[f]
        if f in ("t_out", "y_out") and cap == 0:
            continue
        if f in ("ev_count", "ev_t", "ev_y") and n_events == 0:
            continue
        if f in ("n_seg", "seg_x", "seg_cont") and (seg_cap == 0 or want is None or f not in want):
            continue
        arrays[f] = np.zeros(shp, dtype=_OUT_DTYPES[f])
    st = IvpbOutputs()
    for f in OUTPUT_FIELDS:
        setattr(st, f, ptr(arrays.get(f)))
    return arrays, st


def sparsity_to_csc(sp, n: int):
    """`jac_sparsity` as the reference's Python front end accepts it (a dense array-like of shape (n, n) whose non-zeros
    mark the structure, or a scipy sparse matrix; src/python/sparsity.rs:29-84) -> (colptr[n + 1], rows[nnz]) int32."""
    if hasattr(sp, "toarray"):
        sp = sp.toarray()
    a = np.asarray(sp)
    if a.ndim != 2 or a.shape != (n, n):
        raise ValueError(f"jac_sparsity must have shape ({n}, {n}), got {a.shape}")
    nz = a != 0
    colptr = np.zeros(n + 1, dtype=np.int32)
    colptr[1:] = np.cumsum(nz.sum(axis=0))
    rows = np.ascontiguousarray(np.nonzero(nz.T)[1].astype(np.int32))      # column by column, rows ascending
    if rows.size == 0:
        rows = np.zeros(1, dtype=np.int32)
    return colptr, rows


class MarshalledOptions:
    """Owns the numpy buffers an IvpbOptions struct points into."""

    def __init__(self, opts, n: int, n_events: int):
        o = IvpbOptions()
        self.struct = o
        o.method = int(opts.method)
        self.rtol = np.atleast_1d(np.asarray(opts.rtol, dtype=np.float64)).copy()
        self.atol = np.atleast_1d(np.asarray(opts.atol, dtype=np.float64)).copy()
        for name, a in (("rtol", self.rtol), ("atol", self.atol)):
            if a.ndim != 1 or a.size not in (1, n):
                # reference: Tolerance::Vector length mismatch panics (src/methods/mod.rs:156-161)
                raise ValueError(f"{name} must be a scalar or have length n={n}, got {a.size}")
        o.n_rtol, o.n_atol = self.rtol.size, self.atol.size
        o.rtol, o.atol = ptr(self.rtol), ptr(self.atol)
        for fld in ("first_step", "max_step", "min_step"):
            v = getattr(opts, fld)
            setattr(o, "has_" + fld, int(v is not None))
            setattr(o, fld, float(v) if v is not None else 0.0)
        o.has_max_steps = int(opts.max_steps is not None)
        o.max_steps = int(opts.max_steps) if opts.max_steps is not None else 0
        self.t_eval = None
        o.has_t_eval = int(opts.t_eval is not None)
        if opts.t_eval is not None:
            self.t_eval = np.ascontiguousarray(np.asarray(opts.t_eval, dtype=np.float64).reshape(-1))
            o.n_t_eval = self.t_eval.size
            o.t_eval = ptr(self.t_eval) if self.t_eval.size else None
        o.dense_output = int(bool(opts.dense_output))
        self.ev_dir = self.ev_term = None
        if opts.event_config is not None:
            cfgs = list(opts.event_config)
            if len(cfgs) != n_events:
                raise ValueError(f"event_config has {len(cfgs)} entries, problem has {n_events} events")
            self.ev_dir = np.array([int(c.direction) for c in cfgs], dtype=np.int32)
            self.ev_term = np.array([-1 if c.terminal_count is None else int(c.terminal_count) for c in cfgs],
                                    dtype=np.int64)
            o.n_event_cfg = n_events
            o.ev_direction, o.ev_terminal_count = ptr(self.ev_dir), ptr(self.ev_term)
        o.max_events = int(opts.max_events)
        o.max_out = int(opts.max_out)
        o.jac_mode = int(opts.jac_mode)
        o.flags = int(opts.flags)
        o.max_segments = int(getattr(opts, "max_segments", 0))
        ms = str(getattr(opts, "mass_storage", "Identity")).lower()
        if ms not in ("identity", "full"):
            raise ValueError("mass_storage must be 'Identity' or 'Full' (banded mass matrices are not supported)")
        o.mass_storage = int(ms == "full")
        o.user_solout = int(bool(getattr(opts, "user_solout", False)))
        for fld in ("nind1", "nind2", "nind3"):
            v = getattr(opts, fld, None)
            setattr(o, fld, -1 if v is None else int(v))
        # jac_sparsity (src/python/sparsity.rs:13-102): dense 0/1 array or anything with .toarray() -> compressed columns
        self.sp_colptr = self.sp_rows = None
        sp = getattr(opts, "jac_sparsity", None)
        if sp is not None:
            self.sp_colptr, self.sp_rows = sparsity_to_csc(sp, n)
            o.has_jac_sparsity = 1
            o.jac_sparsity_colptr, o.jac_sparsity_rows = ptr(self.sp_colptr), ptr(self.sp_rows)

    @property
    def seg_cap(self) -> int:
        return int(self.struct.max_segments) if self.struct.dense_output else 0

    @property
    def cap(self) -> int:
        return out_cap(bool(self.struct.has_t_eval), int(self.struct.n_t_eval), int(self.struct.max_out))
