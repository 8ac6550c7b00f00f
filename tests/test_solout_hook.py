"""the solvers' SolOut callback slot with a problem-supplied hook (reference src/solout.rs:55-78) --
ControlFlag::Interrupt / ModifiedSolution as handled by every explicit solver (dop853.rs:246-268,596-624,
dopri5.rs:250-264,414-434, rk23.rs:172-188,258-285, rk4.rs:124-140,196-216) and by RADAU / BDF (radau.rs:336-356,712-740,
bdf.rs:243-274,516-545) -- running on the device, so that
"bounce and continue" (reference examples/bouncing_ball.py:14-36, a host loop of solve_ivp restarts) needs no host
round trip.  Options.user_solout = 1 is the reference's low-level `Method::solve(.., Some(&mut solout))` call.
"""
import numpy as np
import pytest

import ivp_b200 as ib
from ivp_b200 import Direction, EventConfig, Method, Options, Status, synth
from ivp_b200.api import IVPB_FLAG_STRICT_FP, PROBLEMS

EXPLICIT = [Method.RK23, Method.DOPRI5, Method.DOP853, Method.RK4]
# radau.rs:347-351,731-735 (f0 re-evaluated) and bdf.rs:255-271,525-541 (difference table restarted at order 1, new Jacobian)
HOOKED = EXPLICIT + [Method.RADAU, Method.BDF]


def opts_for(method, **kw):
    base = dict(first_step=1e-3) if method == Method.RK4 else dict(rtol=1e-8, atol=1e-10)
    return Options(method=method, user_solout=True, max_out=64, **base, **kw)


def host_restart_chain(oracle, y0, g, drag, restitution, tf, method):
    """reference examples/bouncing_ball.py:14-36 with the oracle as solve_ivp: terminal event, restart from the host."""
    t_curr, state, bounces = 0.0, np.array(y0, dtype=float), []
    kw = dict(first_step=1e-3) if method == Method.RK4 else dict(rtol=1e-8, atol=1e-10)
    for _ in range(64):
        s = oracle.solve_batch(PROBLEMS["bouncing_ball"], t_curr, tf, [state], [[g, drag]],
                               Options(method=method, event_config=[EventConfig(Direction.Negative, 1)], **kw)).solution(0)
        if len(s.t_events[0]) == 0:
            break
        t_curr = float(s.t_events[0][0])
        bounces.append(t_curr)
        v = -restitution * float(s.y_events[0][0][1])
        if abs(v) < 0.1:
            break
        state = np.array([0.0, v])
    return np.array(bounces)


@pytest.mark.parametrize("method", HOOKED)
def test_oracle_bounce_hook_matches_the_host_restart_chain(oracle, method):
    """The in-solver bounce (SolOut + ModifiedSolution) and the reference example's host loop of terminal events find the
    same impacts (to integration accuracy: the restart re-runs hinit, the hook carries the step size over)."""
    y0, g, drag, e = [10.0, 5.0], 9.81, 0.02, 0.75              # the example's numbers
    o = oracle.solve_batch(PROBLEMS["ball_bounce"], 0.0, 15.0, [y0], [[g, drag, e]], opts_for(method))
    m = int(o.n_out[0])
    hook_t = o.t_out[0, 1:m]                                     # sample 0 is the start point
    chain_t = host_restart_chain(oracle, y0, g, drag, e, 15.0, method)
    assert o.status[0] == Status.UserInterrupt                   # the hook stops once |v| < 0.1 after an impact
    assert hook_t.size == chain_t.size >= 10
    np.testing.assert_allclose(hook_t, chain_t, rtol=0, atol=2e-3 if method == Method.RK4 else 2e-6)
    assert np.all(o.y_out[0, 1:m, 0] == 0.0) and np.all(o.y_out[0, 1:m, 1] > 0.0)     # emitted on the ground, moving up
    assert o.t_final[0] == hook_t[-1] and abs(o.y_final[0, 1]) < 0.1
    # every ModifiedSolution costs one extra RHS evaluation (e.g. dop853.rs:613-617)
    plain = oracle.solve_batch(PROBLEMS["ball_bounce"], 0.0, 1.0, [y0], [[g, drag, e]], opts_for(method))
    assert plain.status[0] == Status.Success and plain.n_out[0] == 1 and abs(plain.t_final[0] - 1.0) < 1e-12


def test_oracle_hook_config_errors(oracle):
    with pytest.raises(RuntimeError, match="no SolOut hook"):
        oracle.solve_batch(PROBLEMS["sho"], 0.0, 1.0, [[1.0, 0.0]], None, Options(user_solout=True))


# ---- the CUDA path -----------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("method", HOOKED)
def test_bounce_hook_strict_bit_exact_and_fma_inside_tolerance(oracle, method):
    prob, y0, par, t0, tf = synth.ensemble("ball_bounce", 4000)
    o = oracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts_for(method), nthreads=8)
    g = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts_for(method, flags=IVPB_FLAG_STRICT_FP))
    for f in ("status", "counters", "t_final", "y_final", "h_next", "n_out", "t_out", "y_out"):
        a, b = getattr(g, f), getattr(o, f)
        if f in ("t_out", "y_out"):                              # slots past n_out are not initialised on the device
            valid = np.arange(a.shape[1])[None, :] < np.minimum(g.n_out, a.shape[1])[:, None]
            a, b = a[valid], b[valid]
        assert np.array_equal(a, b), f
    assert set(np.unique(g.status)) <= {int(Status.Success), int(Status.UserInterrupt)} and (g.n_out > 3).all()
    f = ib.solve_ivp_batch(prob, t0, tf, y0, par, opts_for(method))          # default (FMA) build
    # same number of bounces, except where the stop rule |v| < 0.1 is decided by the last bits of an impact velocity
    same = f.n_out == o.n_out
    assert np.array_equal(f.status, o.status) and same.mean() >= 0.99
    m = (np.arange(f.t_out.shape[1])[None, :] < np.minimum(f.n_out, f.t_out.shape[1])[:, None]) & same[:, None]
    np.testing.assert_allclose(f.t_out[m], o.t_out[m], rtol=1e-6, atol=1e-6)
    # (no step-count claim for the FMA build here: every impact time is the end of a 60-step bisection, so a last-bit
    #  difference restarts the following arc from a slightly different point and the step sequences decouple -- measured
    #  52 % equal counts for DOPRI5; the strict build above is bit-identical)


USER_PRINTER = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = y[1]; d[1] = -p[0] * y[0]; }
// src/solout.rs:31-54 (the `Printer` example of the SolOut docs): equidistant output through the step interpolant
template <class Interp, class Emit>
__device__ int ivp_solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit) {
  const double dx = 0.25;
  if (!dense.valid()) { emit(x, y); state[0] = 1.0; return 0; }          // state[0]: index of the next output point
  double yi[2];
  while (state[0] * dx <= x) {
    dense.eval(state[0] * dx, yi);
    emit(state[0] * dx, yi);
    state[0] += 1.0;
  }
  if (y[0] < -0.5) return 1;                                             // Interrupt once the oscillator swings past -0.5
  return 0;
}
"""


@pytest.mark.gpu
@pytest.mark.parametrize("method", [Method.DOPRI5, Method.DOP853, Method.RK23])
def test_user_printer_hook_matches_t_eval_sampling(method):
    """A user SolOut in CUDA C (NVRTC): the docs' equidistant `Printer`, checked against DefaultSolOut's t_eval samples of
    the same solve, plus ControlFlag::Interrupt."""
    from ivp_b200 import api
    N = 600
    y0 = np.stack([1.0 + 0.001 * np.arange(N), np.zeros(N)], axis=1)
    par = np.full((N, 1), 1.0)
    src_plain = USER_PRINTER.split("// src/solout.rs")[0]
    hook = api.Problem.from_cuda_source(USER_PRINTER, n=2, p=1, has_solout=True)
    plain = api.Problem.from_cuda_source(src_plain, n=2, p=1)
    g = ib.solve_ivp_batch(hook, 0.0, 10.0, y0, par, Options(method=method, rtol=1e-9, atol=1e-12, user_solout=True, max_out=64))
    assert np.all(g.status == Status.UserInterrupt)                 # A cos(t) < -0.5 seen at the end of a step
    assert np.all(g.y_final[:, 0] < -0.5) and np.all(g.t_final > np.arccos(-0.5 / y0[:, 0])) and np.all(g.t_final < 3.5)
    te = 0.25 * np.arange(64)
    r = ib.solve_ivp_batch(plain, 0.0, 10.0, y0, par, Options(method=method, rtol=1e-9, atol=1e-12, t_eval=te))
    for i in (0, N // 2, N - 1):
        m = int(g.n_out[i])
        assert m >= 8 and np.array_equal(g.t_out[i, :m], te[:m])
        np.testing.assert_allclose(g.y_out[i, :m], r.y_out[i, :m], rtol=0, atol=1e-12)
        np.testing.assert_allclose(g.y_out[i, :m, 0], y0[i, 0] * np.cos(te[:m]), atol=1e-7)


@pytest.mark.gpu
def test_user_solout_config_errors():
    y0 = np.array([[1.0, 0.0]])
    with pytest.raises(ib.ConfigError, match="no SolOut hook"):
        ib.solve_ivp_batch("sho", 0.0, 1.0, y0, None, Options(user_solout=True))
    par = np.array([[9.81, 0.0, 0.5]])
    with pytest.raises(ib.ConfigError, match="DefaultSolOut"):
        ib.solve_ivp_batch("ball_bounce", 0.0, 1.0, y0, par, Options(user_solout=True, t_eval=[0.5]))
    # without user_solout the same problem is an ordinary ODE under DefaultSolOut: the ball falls through the floor
    g = ib.solve_ivp_batch("ball_bounce", 0.0, 3.0, np.array([[10.0, 0.0]]), par, Options(rtol=1e-8, atol=1e-10))
    assert g.status[0] == Status.Success and g.y_final[0, 0] < 0.0


# ---- hooks in the warp-per-trajectory kernels (n > 32 explicit, n > 8 RADAU / BDF) -------------------------------------
# The device twin of oracle/test_problems.hpp DecayKick40 (problem 110): 40 decaying components; the hook bisects the point
# where y[0] falls below 0.5 on the step interpolant, restarts there with the whole state doubled, records it, and stops at
# the third kick.  `dense.buffer()` provides the n doubles eval() fills -- shared by the warp in these kernels.
KICK40_SRC = r"""
__device__ double ivp_ode_i(double t, const double* y, const double* p, int i) { return -(0.5 + 0.05 * (double)i) * y[i]; }
template <class Interp, class Emit>
__device__ int ivp_solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit) {
  if (!dense.valid()) { emit(x, y); return 0; }
  if (!(y[0] < 0.5)) return 0;
  double lo = xold, hi = x;
  double* yi = dense.buffer();
  for (int it = 0; it < 40; ++it) {
    const double mid = 0.5 * (lo + hi);
    dense.eval(mid, yi);
    if (yi[0] < 0.5) hi = mid; else lo = mid;
  }
  dense.eval(hi, yi);
  x = hi;
  for (int i = 0; i < 40; ++i) y[i] = 2.0 * yi[i];
  state[0] += 1.0;
  emit(x, y);
  if (state[0] >= 3.0) return 1;
  return 2;
}
"""


def kick40_y0(N):
    rng = np.random.default_rng(40)
    y0 = rng.uniform(0.8, 1.2, (N, 40))
    y0[:, 0] = np.linspace(0.9, 1.1, N)
    return y0


@pytest.mark.parametrize("method", HOOKED)
def test_oracle_kick40_hook(oracle, method):
    """The oracle side of the warp-kernel hook test: three kicks, each where y[0] crosses 0.5 (y[0] = y0 e^{-t/2})."""
    y0 = kick40_y0(3)
    o = oracle.solve_batch(110, 0.0, 20.0, y0, None, opts_for(method))
    assert np.all(o.status == Status.UserInterrupt) and np.all(o.n_out == 4)
    t1 = 2.0 * np.log(y0[:, 0] / 0.5)                                 # first crossing; afterwards y[0] restarts from 1.0
    tol = 5e-3 if method == Method.RK4 else 1e-5
    np.testing.assert_allclose(o.t_out[:, 1], t1, rtol=tol)
    np.testing.assert_allclose(o.t_out[:, 3] - o.t_out[:, 2], 2.0 * np.log(2.0), rtol=tol)
    np.testing.assert_allclose(o.y_out[:, 1:4, 0], 1.0, rtol=tol)
    assert np.array_equal(o.t_final, o.t_out[:, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("method", HOOKED)
def test_hook_in_warp_per_trajectory_kernels(oracle, method):
    """Options.user_solout on a problem with n = 40: explicit methods run one trajectory per warp (n > 32), RADAU / BDF the
    warp-cooperative kernels (n > 8).  The hook sees the full state in shared memory, evaluates the step interpolant into
    dense.buffer(), moves (x, y) and returns ModifiedSolution / Interrupt; strict build bit-identical to the oracle."""
    from ivp_b200 import api
    N = 37
    y0 = kick40_y0(N)
    user = api.Problem.from_cuda_source(KICK40_SRC, n=40, has_solout=True)
    opts = opts_for(method, flags=IVPB_FLAG_STRICT_FP)
    g = ib.solve_ivp_batch(user, 0.0, 20.0, y0, None, opts)
    o = oracle.solve_batch(110, 0.0, 20.0, y0, None, opts)
    assert np.array_equal(g.status, o.status) and np.all(g.status == Status.UserInterrupt)
    assert np.array_equal(g.n_out, o.n_out) and np.array_equal(g.counters, o.counters)
    assert np.array_equal(g.t_out, o.t_out) and np.array_equal(g.y_out, o.y_out)
    assert np.array_equal(g.t_final, o.t_final) and np.array_equal(g.y_final, o.y_final)


@pytest.mark.parametrize("method", HOOKED)
def test_nvrtc_compiles_the_hook_for_the_warp_kernels(method):
    """No GPU needed: the K_USER instances of erk_warp_kernel / implicit_warp_kernel compile for sm_100a (strict build)."""
    import ctypes
    from ivp_b200 import api
    lib = api.load_library()
    lib.ivpb_debug_nvrtc_compile.restype = ctypes.c_longlong
    log = ctypes.create_string_buffer(1 << 16)
    size = lib.ivpb_debug_nvrtc_compile(KICK40_SRC.encode(), 40, 0, 0, 4, int(method), 4, 1, log, 1 << 16)
    assert size > 10000, log.value.decode()
