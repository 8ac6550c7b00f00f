import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth, api
from ivp_b200.api import IVPB_FLAG_NO_REFILL
prob, y0, par, t0, tf = synth.ensemble("vdp", 3000)
for k in range(1, 12):
    o = dict(method=Method.DOP853, rtol=1e-8, atol=1e-8, max_steps=k)
    a = ib.solve_ivp_batch(prob, t0, 100.0, y0, par, Options(**o))
    b = ib.solve_ivp_batch(prob, t0, 100.0, y0, par, Options(flags=IVPB_FLAG_NO_REFILL, **o))
    bad = np.where((a.h_next != b.h_next) | (a.y_final != b.y_final).any(axis=1))[0]
    print("max_steps", k, "differing", bad[:8], [ (a.h_next[i], b.h_next[i], a.t_final[i]-b.t_final[i], (a.y_final[i]-b.y_final[i]).tolist(), a.counters[i].tolist()) for i in bad[:2]])
