"""Tiny run through every kernel family (for compute-sanitizer memcheck / racecheck on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import IVPB_FLAG_FAST_FP, IVPB_FLAG_STRICT_FP

def run(wl, N, tf=None, **kw):
    prob, y0, par, t0, tfd = synth.ensemble(wl, N)
    g = ib.solve_ivp_batch(prob, t0, tf if tf is not None else tfd, y0, par, Options(**kw))
    print(wl, kw.get("method").name, "status", np.bincount(g.status, minlength=3)[:3], flush=True)
    return g

run("vdp", 96, 10.0, method=Method.DOP853, rtol=1e-8, atol=1e-8, t_eval=np.linspace(0, 10, 7))
run("vdp", 96, 10.0, method=Method.DOPRI5, rtol=1e-6, atol=1e-9, flags=IVPB_FLAG_STRICT_FP)
run("ball", 96, None, method=Method.DOPRI5, rtol=1e-8, atol=1e-10, max_events=2)
g = run("vdp", 64, 5.0, method=Method.RK23, rtol=1e-5, atol=1e-8, dense_output=True, max_segments=256, max_out=256)
print("dense", g.sol_many([0, 63], [1.0, 4.0])[1])
run("linear100", 10, 2.0, method=Method.DOPRI5, rtol=1e-6, atol=1e-8, t_eval=np.linspace(0, 2, 3))
run("medakzo", 6, 0.2, method=Method.DOP853, rtol=1e-6, atol=1e-8, flags=IVPB_FLAG_STRICT_FP)
run("robertson", 64, 1e3, method=Method.RADAU, rtol=1e-6, atol=1e-6)
run("robertson", 64, 1e3, method=Method.BDF, rtol=1e-6, atol=1e-6, flags=IVPB_FLAG_FAST_FP)
run("cr3bp", 64, 1.0, method=Method.RADAU, rtol=1e-6, atol=1e-8)
run("cr3bp", 64, 1.0, method=Method.BDF, rtol=1e-6, atol=1e-8)
run("medakzo", 5, 0.5, method=Method.RADAU, rtol=1e-5, atol=1e-7)
run("medakzo", 5, 0.5, method=Method.BDF, rtol=1e-5, atol=1e-7, flags=IVPB_FLAG_FAST_FP)
run("robertson_dae", 64, 1e3, method=Method.RADAU, rtol=1e-6, atol=1e-10, mass_storage="Full")          # mass matrix
run("mass_linear3", 64, None, method=Method.RADAU, rtol=1e-8, atol=1e-11, mass_storage="Full", jac_mode=1)
for m in (Method.DOPRI5, Method.RADAU, Method.BDF):                                                       # SolOut hooks
    kw = dict(first_step=1e-3) if m == Method.RK4 else dict(rtol=1e-8, atol=1e-10)
    g = run("ball_bounce", 64, None, method=m, user_solout=True, max_out=48, **kw)
    print("  bounces", int(g.n_out.min()) - 1, "..", int(g.n_out.max()) - 1)
run("robertson", 5000, 1e3, method=Method.BDF, rtol=1e-6, atol=1e-6)                                      # locality order (N >= 4096)
print("sanity ok")
