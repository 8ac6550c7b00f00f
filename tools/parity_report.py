"""Parity report (run on the GPU box): GPU vs CPU oracle on every BASELINE workload, for the DEFAULT call (no flag: the
parity pilot picks the build for the explicit methods, RADAU / BDF run strict) and for the two builds forced by flag.
Writes a markdown table to stdout; committed as profiles/parity_r2.md."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS, IVPB_FLAG_STRICT_FP, IVPB_FLAG_FAST_FP
from oracle import pyoracle

CASES = [  # workload, method, rtol, atol, N, extra
    ("vdp", Method.DOP853, 1e-8, 1e-8, 65536, {}),          # north star
    ("vdp", Method.DOPRI5, 1e-6, 1e-9, 32768, {}),          # configs[0] tolerances
    ("vdp", Method.RK23, 1e-5, 1e-8, 16384, {}),
    ("decay", Method.DOPRI5, 1e-6, 1e-9, 65536, {}),        # configs[1]
    ("lorenz", Method.DOPRI5, 1e-6, 1e-9, 16384, {}),
    ("lorenz", Method.RK4, 1e-6, 1e-9, 16384, {"first_step": 0.01}),
    ("cr3bp", Method.DOP853, 1e-10, 1e-12, 4096, {"t_eval": 101}),   # configs[2]
    ("ball", Method.DOPRI5, 1e-8, 1e-10, 65536, {}),        # configs[3]
    ("robertson", Method.RADAU, 1e-6, 1e-6, 16384, {}),     # configs[4]
    ("robertson", Method.BDF, 1e-6, 1e-6, 16384, {}),
    ("vdp_stiff", Method.RADAU, 1e-4, 1e-6, 8192, {}),
    ("vdp_stiff", Method.BDF, 1e-4, 1e-6, 8192, {}),
    ("medakzo", Method.RADAU, 1e-5, 1e-7, 256, {}),         # n = 64, warp-cooperative LU
    ("linear100", Method.DOPRI5, 1e-6, 1e-8, 2048, {}),     # n = 100, warp per trajectory
]

print("| workload | method | N | build | status equal | step counts equal | all 6 counters equal | inside max(10 rtol |y|, 10 atol) | bit-identical y_final | samples/events equal |")
print("|---|---|---|---|---|---|---|---|---|---|")
for wl, m, rtol, atol, N, extra in CASES:
    prob, y0, par, t0, tf = synth.ensemble(wl, N)
    kw = dict(extra)
    if "t_eval" in kw:
        kw["t_eval"] = np.linspace(t0, tf, kw["t_eval"])
    for flags, name in ((0, "default"), (IVPB_FLAG_FAST_FP, "fma (flag)"), (IVPB_FLAG_STRICT_FP, "strict (flag)")):
        opts = Options(method=m, rtol=rtol, atol=atol, flags=flags, max_events=2, **kw)
        g = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts)
        if flags == 0:
            fm = ib.api.default_context().last_fp_mode()
            name = f"default -> {fm['mode']} ({fm['source']}" + (f": {fm['out_of_tolerance']} of {fm['sample']} sampled out of tolerance, {fm['step_count_mismatches']} count mismatches" if 'sample' in fm else "") + ")"
        o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=os.cpu_count())
        steps = (g.naccpt == o.naccpt) & (g.nrejct == o.nrejct) & (g.nstep == o.nstep)
        allc = (g.counters == o.counters).all(axis=1)
        tol = np.abs(g.y_final - o.y_final) <= np.maximum(10 * rtol * np.abs(o.y_final), 10 * atol)
        bits = (g.y_final.view(np.uint64) == o.y_final.view(np.uint64)).all(axis=1)
        extra_eq = "-"
        if g.y_out is not None:
            tol_s = np.abs(g.y_out - o.y_out) <= np.maximum(10 * rtol * np.abs(o.y_out), 10 * atol)
            extra_eq = (f"n_out {np.array_equal(g.n_out, o.n_out)}, samples inside tolerance {tol_s.all(axis=(1, 2)).mean():.5f}, "
                        f"y_out bits {np.mean(g.y_out.view(np.uint64) == o.y_out.view(np.uint64)):.4f}")
        if g.ev_t is not None:
            extra_eq = f"ev_count {np.array_equal(g.ev_count, o.ev_count)}, ev_t bits {np.mean(g.ev_t.view(np.uint64) == o.ev_t.view(np.uint64)):.4f}"
        print(f"| {wl} | {m.name} | {N} | {name} | {np.array_equal(g.status, o.status)} | {steps.mean():.5f} | {allc.mean():.5f} | "
              f"{tol.all(axis=1).mean():.5f} | {bits.mean():.5f} | {extra_eq} |", flush=True)
