"""Diagnostic: GPU-vs-oracle parity statistics of the implicit path (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth
from ivp_b200.api import PROBLEMS, IVPB_FLAG_STRICT_FP
from oracle import pyoracle


def stats(wl, method, rtol, atol, N, flags=0, **kw):
    prob, y0, par, t0, tf = synth.ensemble(wl, N)
    opts = Options(method=method, rtol=rtol, atol=atol, flags=flags, **kw)
    t = time.time()
    g = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts)
    tg = time.time() - t
    t = time.time()
    o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=os.cpu_count())
    to = time.time() - t
    same = (g.naccpt == o.naccpt) & (g.nrejct == o.nrejct) & (g.nstep == o.nstep)
    allc = (g.counters == o.counters).all(axis=1)
    err = np.abs(g.y_final - o.y_final) / np.maximum(10 * rtol * np.abs(o.y_final), 10 * atol)
    worst = err.max(axis=1)
    print(f"{wl:10s} {method.name:6s} flags={flags} kw={kw} N={N} status_eq={np.array_equal(g.status, o.status)} "
          f"step_parity={same.mean():.5f} all_counters={allc.mean():.5f} tol_viol={np.mean(worst > 1):.5f} "
          f"max_err_ratio={worst.max():.3g} (same: {worst[same].max() if same.any() else float('nan'):.3g}) "
          f"status={np.bincount(g.status, minlength=7)} gpu {tg:.3f}s cpu {to:.3f}s", flush=True)
    return g, o


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for fl in (IVPB_FLAG_STRICT_FP, 0):
        for jm in (0, 1):
            for m in (Method.RADAU, Method.BDF):
                stats("robertson", m, 1e-6, 1e-6, N, fl, jac_mode=jm)
                stats("vdp_stiff", m, 1e-4, 1e-6, N, fl, jac_mode=jm)
        for m in (Method.RADAU, Method.BDF):
            stats("vdp", m, 1e-6, 1e-8, N, fl)
            stats("decay", m, 1e-6, 1e-9, N, fl)
            stats("cr3bp", m, 1e-6, 1e-8, min(N, 512), fl)
