#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
IVPB_LIB=ivp_b200/lib/libivpb_gdbg.so timeout 100 python - > $O/r2z4_guard_debug.log 2>&1 <<PY
import numpy as np, ivp_b200 as ib
from ivp_b200 import Method, Options, synth
for wl, m in (("robertson", Method.BDF),):
    prob, y0, par, t0, tf = synth.ensemble(wl, 256)
    g = ib.solve_ivp_batch(prob, t0, tf, y0, par, Options(method=m, rtol=1e-6, atol=1e-6))
    print(wl, m, "status", np.unique(g.status), "reruns", ib.api.default_context().last_reruns(), flush=True)
PY
grep -v "div<0>.*a=0 \|div<0> a=-0 " $O/r2z4_guard_debug.log | head -30; wc -l $O/r2z4_guard_debug.log
