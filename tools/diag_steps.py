"""Diagnostic: where do GPU and oracle step sequences diverge? (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS, IVPB_FLAG_STRICT_FP
from oracle import pyoracle

N = 8192
prob, y0, par, t0, tf = synth.ensemble("vdp", N)
opts = Options(method=Method.DOPRI5, rtol=1e-6, atol=1e-9, max_out=1024)
g = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts)
o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=os.cpu_count())
bad = np.where((g.naccpt != o.naccpt) | (g.nrejct != o.nrejct))[0]
print("mismatching:", len(bad), bad[:20])
for i in bad[:6]:
    tg, to = g.t_out[i], o.t_out[i]
    m = min(g.n_out[i], o.n_out[i])
    d = np.abs(tg[:m] - to[:m])
    first = np.argmax(d > 1e-9) if np.any(d > 1e-9) else -1
    print(f"traj {i}: y0={y0[i]} acc g/o {g.naccpt[i]}/{o.naccpt[i]} rej {g.nrejct[i]}/{o.nrejct[i]} first_div_step={first}")
    lo = max(0, first - 6)
    for j in range(lo, min(m, first + 3)):
        print(f"    step {j}: t_g={tg[j]:.17g} t_o={to[j]:.17g} diff={tg[j]-to[j]:.3g}  y_g={g.y_out[i,j]} dy={g.y_out[i,j]-o.y_out[i,j]}")
# growth of |t_g - t_o| along good trajectories
good = np.where((g.naccpt == o.naccpt) & (g.nrejct == o.nrejct))[0][:2000]
mx = np.array([np.abs(g.t_out[i, :g.n_out[i]] - o.t_out[i, :g.n_out[i]]).max() for i in good])
print("good trajectories: max |dt| percentiles", np.percentile(mx, [50, 90, 99, 100]))
