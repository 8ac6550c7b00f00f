import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ivp_b200 as ib
from ivp_b200 import Method, Options, synth, api
from ivp_b200.api import IVPB_FLAG_NO_REFILL, IVPB_FLAG_STRICT_FP
USER_VDP = r"""
__device__ void ivp_ode(double t, const double* y, const double* p, double* dydt) {
  dydt[0] = y[1];
  dydt[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
}
"""
prob, y0, par, t0, tf = synth.ensemble("vdp", 3000)
user = api.Problem.from_cuda_source(USER_VDP, n=2, p=1)
for m, kw in ((Method.DOP853, dict(rtol=1e-8, atol=1e-8)), (Method.DOPRI5, dict(rtol=1e-6, atol=1e-9)), (Method.RK23, dict(rtol=1e-5, atol=1e-8)), (Method.RK4, dict(first_step=0.01))):
    for tf_ in (0.5, 100.0):
        for fl in (0, IVPB_FLAG_STRICT_FP):
            a = ib.solve_ivp_batch(prob, t0, tf_, y0, par, Options(method=m, flags=fl, **kw))
            b = ib.solve_ivp_batch(prob, t0, tf_, y0, par, Options(method=m, flags=fl | IVPB_FLAG_NO_REFILL, **kw))
            c = ib.solve_ivp_batch(user, t0, tf_, y0, par, Options(method=m, flags=fl, **kw))
            print(m.name, "tf", tf_, "flags", fl, "static-vs-queue: ndiff_y", int((a.y_final != b.y_final).any(axis=1).sum()), "max", np.abs(a.y_final - b.y_final).max(),
                  "counters diff", int((a.counters != b.counters).any(axis=1).sum()), "h_next diff", int((a.h_next != b.h_next).sum()),
                  "| nvrtc-vs-builtin ndiff_y", int((a.y_final != c.y_final).any(axis=1).sum()), "max", np.abs(a.y_final - c.y_final).max(), "counters", int((a.counters != c.counters).any(axis=1).sum()))
