import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS, IVPB_FLAG_STRICT_FP
from oracle import pyoracle
N = 2048
prob, y0, par, t0, tf = synth.ensemble("cr3bp", N)
for frac in (0.25, 0.5, 0.9, 1.0):
    for fl in (0, IVPB_FLAG_STRICT_FP):
        te = np.linspace(t0, tf * frac, 101)
        opts = Options(method=Method.DOP853, rtol=1e-10, atol=1e-12, t_eval=te, flags=fl)
        g = ib.solve_ivp_batch(prob, t0, tf * frac, y0, par, opts)
        o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf * frac, y0, par, opts, nthreads=os.cpu_count())
        same = (g.naccpt == o.naccpt) & (g.nrejct == o.nrejct)
        d = np.abs(g.y_out - o.y_out).max(axis=(1, 2))
        rel = (np.abs(g.y_out - o.y_out) / np.maximum(10e-10 * np.abs(o.y_out), 10e-12)).max(axis=(1, 2))
        print(f"span {frac:4.2f}T flags={fl} parity={same.mean():.4f} max|dy| pct50/90/99/100 = "
              f"{np.percentile(d,50):.2e} {np.percentile(d,90):.2e} {np.percentile(d,99):.2e} {d.max():.2e}; tol_viol={np.mean(rel>1):.4f} "
              f"status_eq={np.array_equal(g.status,o.status)} naccpt mean {g.naccpt.mean():.0f}")
