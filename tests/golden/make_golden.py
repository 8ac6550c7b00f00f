#!/usr/bin/env python
"""Regenerates tests/golden/oracle_cases.npz.

Provenance: the reference (Rust crate `ivp`) cannot be built or run in this image (no rustc / cargo / maturin),
so these vectors come from the CPU ORACLE (oracle/, the line-by-line restatement of the reference), run on
x86-64 glibc 2.39.  They pin (a) the oracle against accidental drift (tests/test_oracle_pins.py loads them on
the CPU) and (b) the strict CUDA build against fixed bits (tests/test_gpu_parity.py).  The reference's own
literal golden numbers (tests/test_ivp.py:152-170 cannon event, tests/ivp.rs pi/2 crossings, ...) are asserted
directly in tests/test_oracle_pins.py with their file:line.

usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ivp_b200 import Method, Options, synth          # noqa: E402
from ivp_b200.api import PROBLEMS                      # noqa: E402
from oracle import pyoracle                            # noqa: E402

#: name -> (workload, N, t_end override or None, Options kwargs)
CASES = {
    "vdp_dop853": ("vdp", 64, None, dict(method=Method.DOP853, rtol=1e-8, atol=1e-8)),
    "vdp_dopri5_teval": ("vdp", 48, 20.0, dict(method=Method.DOPRI5, rtol=1e-6, atol=1e-9, t_eval=np.linspace(0.0, 20.0, 9))),
    "vdp_rk23": ("vdp", 32, 10.0, dict(method=Method.RK23, rtol=1e-5, atol=1e-8)),
    "lorenz_rk4": ("lorenz", 32, 2.0, dict(method=Method.RK4, first_step=0.005)),
    "cr3bp_dop853_teval": ("cr3bp", 24, None, dict(method=Method.DOP853, rtol=1e-10, atol=1e-12, t_eval="span101")),
    "ball_dopri5_events": ("ball", 64, None, dict(method=Method.DOPRI5, rtol=1e-8, atol=1e-10, max_events=2)),
    "robertson_radau": ("robertson", 32, None, dict(method=Method.RADAU, rtol=1e-6, atol=1e-6)),
    "robertson_bdf_jac": ("robertson", 32, None, dict(method=Method.BDF, rtol=1e-6, atol=1e-6, jac_mode=1)),
    "vdpstiff_radau": ("vdp_stiff", 24, None, dict(method=Method.RADAU, rtol=1e-4, atol=1e-6)),
    "vdpstiff_bdf": ("vdp_stiff", 24, None, dict(method=Method.BDF, rtol=1e-4, atol=1e-6)),
    "medakzo_bdf": ("medakzo", 6, 7.0, dict(method=Method.BDF, rtol=1e-5, atol=1e-7)),
    # RADAU with a mass matrix (+ index-2 scaling), the SolOut hook with ModifiedSolution
    "robertson_dae_radau_mass": ("robertson_dae", 24, 1e6, dict(method=Method.RADAU, rtol=1e-6, atol=1e-10, mass_storage="Full",
                                                               t_eval=np.array([1e-2, 1.0, 1e2, 1e4, 1e6]))),
    "robertson_dae_radau_nind2": ("robertson_dae", 16, 1e3, dict(method=Method.RADAU, rtol=1e-6, atol=1e-10, mass_storage="Full", nind2=1)),
    "mass_linear3_radau_jac": ("mass_linear3", 24, None, dict(method=Method.RADAU, rtol=1e-8, atol=1e-11, mass_storage="Full", jac_mode=1)),
    "ball_bounce_dop853_hook": ("ball_bounce", 32, None, dict(method=Method.DOP853, rtol=1e-8, atol=1e-10, user_solout=True, max_out=48)),
    "ball_bounce_radau_hook": ("ball_bounce", 16, None, dict(method=Method.RADAU, rtol=1e-8, atol=1e-10, user_solout=True, max_out=48)),
    "ball_bounce_bdf_hook": ("ball_bounce", 16, None, dict(method=Method.BDF, rtol=1e-8, atol=1e-10, user_solout=True, max_out=48)),
}
FIELDS = ("status", "counters", "t_final", "y_final", "h_next", "n_out", "t_out", "y_out", "ev_count", "ev_t", "ev_y")


def case_inputs(name):
    wl, N, t_end, kw = CASES[name]
    prob, y0, par, t0, tf = synth.ensemble(wl, N)
    tf = tf if t_end is None else t_end
    kw = dict(kw)
    if isinstance(kw.get("t_eval"), str):
        kw["t_eval"] = np.linspace(t0, tf, 101)
    return prob, y0, par, t0, tf, Options(**kw)


def main():
    out = {}
    for name in CASES:
        prob, y0, par, t0, tf, opts = case_inputs(name)
        o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=os.cpu_count())
        for f in FIELDS:
            v = getattr(o, f)
            if v is not None:
                out[f"{name}/{f}"] = v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_cases.npz"), **out)
    print("wrote", len(out), "arrays for", len(CASES), "cases")


if __name__ == "__main__":
    main()
