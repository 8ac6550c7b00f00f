"""Diagnostic: GPU-vs-oracle parity statistics per workload (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS, IVPB_FLAG_STRICT_FP
from oracle import pyoracle

def stats(wl, method, rtol, atol, N, flags=0, **kw):
    prob, y0, par, t0, tf = synth.ensemble(wl, N)
    opts = Options(method=method, rtol=rtol, atol=atol, flags=flags, **kw)
    g = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts)
    o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=os.cpu_count())
    same = (g.naccpt == o.naccpt) & (g.nrejct == o.nrejct)
    err = np.abs(g.y_final - o.y_final) / np.maximum(10 * rtol * np.abs(o.y_final), 10 * atol)
    worst = err.max(axis=1)
    print(f"{wl:8s} {method.name:7s} flags={flags} N={N} status_eq={np.array_equal(g.status,o.status)} "
          f"count_parity={same.mean():.5f} tol_viol={np.mean(worst>1):.5f} "
          f"max_err_ratio(all)={worst.max():.3g} max_err_ratio(same)={worst[same].max():.3g} "
          f"max|dt_final|={np.abs(g.t_final-o.t_final).max():.3g}")
    if g.ev_t is not None:
        d = np.abs(g.ev_t - o.ev_t)
        print("   events: count_eq", np.array_equal(g.ev_count, o.ev_count), "max|dt_ev|", d.max(), "on same:", d[same].max(),
              "frac>1e-10:", np.mean(d > 1e-10))
    return g, o

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    for fl in (0, IVPB_FLAG_STRICT_FP):
        stats("vdp", Method.DOP853, 1e-8, 1e-8, N, fl)
        stats("vdp", Method.DOPRI5, 1e-6, 1e-9, N, fl)
        stats("vdp", Method.RK23, 1e-5, 1e-8, N, fl)
        stats("decay", Method.DOPRI5, 1e-6, 1e-9, N, fl)
        stats("ball", Method.DOPRI5, 1e-8, 1e-10, N, fl)
        stats("cr3bp", Method.DOP853, 1e-10, 1e-12, min(N, 2048), fl)
        stats("lorenz", Method.DOPRI5, 1e-6, 1e-9, min(N, 2048), fl)
