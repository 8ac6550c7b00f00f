"""The reference's Python test-suite for the solve path, restated against this repo.

Every test below restates one test of reference tests/test_basic_integration.py, test_t_eval.py, test_events.py,
test_step_control.py, test_edge_cases.py or test_args.py (cited per test) through `ivp_b200.scipy_api.solve_ivp`,
the front end with the reference binding's signature (src/python/solve.rs:153-432).  The reference's Python
callables become device problems: CUDA C sources compiled by NVRTC with the solver (`SRC_*` below), with the very
same right-hand sides restated for the CPU oracle in oracle/test_problems.hpp.

Each test runs on two back ends:
  * `oracle` (CPU, not marked gpu): `api.solve_ivp_batch` is replaced by the oracle, so the SAME front-end code and
    the SAME assertions pin the oracle against the reference's own expectations;
  * `gpu` (marked gpu): the CUDA path through the C ABI, additionally compared with the oracle run of the same call.
"""
import numpy as np
import pytest

from ivp_b200 import Method, Status, api, scipy_api

# ---- the reference tests' callables as device code ------------------------------------------------------------------
RATIONAL_ODE = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) {
  d[0] = y[1] / t;
  d[1] = y[1] * (y[0] + 2.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
}
__device__ void ivp_jac(double t, const double* y, const double* p, double* J) {
  J[0] = 0.0; J[1] = 1.0 / t;
  J[2] = -2.0 * y[1] * y[1] / (t * (y[0] - 1.0) * (y[0] - 1.0));
  J[3] = (y[0] + 4.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
}
"""
# tests/test_helpers.py:23-25,34-40 + tests/test_events.py:13-17,103-104 (third event threshold = p[0])
SRC_RATIONAL_EV = RATIONAL_ODE + """
__device__ void ivp_events(double t, const double* y, const double* p, double* g) {
  g[0] = y[0] - pow(y[1], 0.7);
  g[1] = pow(y[1], 0.6) - y[0];
  g[2] = t - p[0];
}
"""
# tests/test_args.py:11-35
SRC_SYS3 = """
__device__ void ivp_ode(double t, const double* w, const double* p, double* d) {
  d[0] = -p[0] * w[1]; d[1] = p[0] * w[0]; d[2] = p[1] * w[2] * (1.0 - w[2]);
}
__device__ void ivp_events(double t, const double* w, const double* p, double* g) { g[0] = w[0]; g[1] = w[1]; g[2] = w[2] - p[2]; }
__device__ void ivp_jac(double t, const double* w, const double* p, double* J) {
  J[0] = 0.0;  J[1] = -p[0]; J[2] = 0.0;
  J[3] = p[0]; J[4] = 0.0;   J[5] = 0.0;
  J[6] = 0.0;  J[7] = 0.0;   J[8] = p[1] * (1.0 - 2.0 * w[2]);
}
"""
SRC_SCALE1 = "__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = p[0] * y[0]; }"
SRC_SCALE2 = "__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = p[0] * y[0]; d[1] = p[0] * y[1]; }"
# tests/test_edge_cases.py:76-88
SRC_RADIAL = """
__device__ void ivp_ode(double t, const double* s, const double* p, double* d) {
  const double r = exp(t);
  const double V = -11.0 / r + 10.0 * r / (0.05 + r * r);
  d[0] = r * s[1];
  d[1] = -2.0 * r * ((-0.2 - V) * s[0] + 1.0 / r * s[1]);
}
"""
# tests/test_edge_cases.py:104-110
SRC_CONST = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) {
  d[0] = 1.73307544e-02; d[1] = 6.49376470e-06; d[2] = 0.0; d[3] = 0.0;
}
"""
# tests/test_helpers.py:11-16
SRC_LINEAR = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = -y[0] - 5.0 * y[1]; d[1] = y[0] + y[1]; }
__device__ void ivp_jac(double t, const double* y, const double* p, double* J) { J[0] = -1.0; J[1] = -5.0; J[2] = 1.0; J[3] = 1.0; }
"""
ORACLE_ID = {SRC_LINEAR: 106, SRC_RATIONAL_EV: 100, SRC_SYS3: 101, SRC_SCALE1: 102, SRC_SCALE2: 103, SRC_RADIAL: 104, SRC_CONST: 105}

METHODS = ["RK23", "RK45", "DOP853", "Radau", "BDF"]


def sol_rational(t):                                      # tests/test_helpers.py:50-51
    t = np.asarray(t, dtype=float)
    return np.asarray((t / (t + 10), 10 * t / (t + 10) ** 2))


def compute_error(y, y_true, rtol, atol):                 # tests/test_helpers.py:124-126
    e = (y - y_true) / (atol + rtol * np.abs(y_true))
    return np.linalg.norm(e, axis=0) / np.sqrt(e.shape[0])


class Ev:
    """An event spec: carries SciPy's `terminal` / `direction` function attributes."""

    def __init__(self, terminal=False, direction=0):
        self.terminal, self.direction = terminal, direction


# ---- back ends --------------------------------------------------------------------------------------------------------
class _OracleDense:
    """Stands in for the device context behind BatchSolution.sol_many / sol_span on the oracle back end."""

    def __init__(self, oracle, pid, t0, tf, Y0, params, opts):
        self.a = (oracle, pid, t0, tf, Y0, params, opts)

    def _one(self, i, ts, extrapolate=False):
        oracle, pid, t0, tf, Y0, params, opts = self.a
        return oracle.dense_eval(pid, t0, tf, Y0[i], None if params is None else params[i], opts, ts, extrapolate)

    def dense_eval(self, traj, ts, n, extrapolate=False, generation=0):
        traj, ts = np.atleast_1d(traj), np.atleast_1d(np.asarray(ts, dtype=float))
        y, ok = np.zeros((ts.size, n)), np.zeros(ts.size, dtype=bool)
        for i in np.unique(traj):
            m = traj == i
            y[m], ok[m], _ = self._one(int(i), ts[m], extrapolate)
        return y, ok

    def dense_span(self, first, count, generation=0):
        t0, t1, m = np.zeros(count), np.zeros(count), np.zeros(count, dtype=np.int32)
        for k in range(count):
            _, _, span = self._one(first + k, [])
            if span is not None:
                t0[k], t1[k], m[k] = span[0], span[1], 1
        return t0, t1, m


def _oracle_batch(oracle):
    def solve(problem, t0, tf, y0, params=None, options=None, ctx=None, want=None):
        pid = ORACLE_ID[problem.cuda_src] if problem.cuda_src is not None else problem.handle
        b = oracle.solve_batch(pid, t0, tf, y0, params, options)
        b.extras.update(ctx=_OracleDense(oracle, pid, t0, tf, np.atleast_2d(y0), params, options),
                        dense=bool(options.dense_output))
        return b
    return solve


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def solve_ivp(request, oracle, monkeypatch):
    """`solve_ivp` with the reference binding's signature on the chosen back end."""
    if request.param == "oracle":
        monkeypatch.setattr(api.Problem, "builtin", staticmethod(_cpu_builtin(oracle)))
        monkeypatch.setattr(scipy_api.api, "solve_ivp_batch", _oracle_batch(oracle))
        return scipy_api.solve_ivp
    dev_batch, ora_batch = api.solve_ivp_batch, _oracle_batch(oracle)

    def both(problem, t0, tf, y0, params=None, options=None, ctx=None, want=None):
        """The device solve, cross-checked against the oracle on the same call (NVRTC problems are compiled apart
        from the oracle, so the check is the north star's value tolerance, not bits)."""
        g = dev_batch(problem, t0, tf, y0, params, options, ctx, want)
        o = ora_batch(problem, t0, tf, y0, params, options)
        assert np.array_equal(g.status, o.status)
        rt = float(np.max(np.atleast_1d(options.rtol)))
        at = float(np.max(np.atleast_1d(options.atol)))
        if rt >= 1e-13:          # (test_array_rtol asks for rtol = 1e-16 on one component: a pure stress setting)
            np.testing.assert_allclose(g.y_final, o.y_final, rtol=100 * rt, atol=100 * at)
        if g.ev_count is not None:
            assert np.array_equal(g.ev_count, o.ev_count)
        return g
    monkeypatch.setattr(scipy_api.api, "solve_ivp_batch", both)
    return scipy_api.solve_ivp


def _cpu_builtin(oracle):
    """Problem.builtin without libivpb (dims from the oracle), for the CPU back end."""
    def builtin(name_or_id):
        pid = api.PROBLEMS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        n, p, ne = oracle.dims(pid)
        return api.Problem(pid, n, p, ne, str(name_or_id))
    return builtin


def rational(solve_ivp, t_span, events=None, p=7.4, **kw):
    """fun_rational with the three event functions compiled in; `events` = specs for the ones a test uses."""
    if events is None:
        return solve_ivp(SRC_RATIONAL_EV, t_span, [1 / 3, 2 / 9], args=(p,), n_events=3, **kw)
    return solve_ivp(SRC_RATIONAL_EV, t_span, [1 / 3, 2 / 9], args=(p,), events=list(events), **kw)


# ---- tests/test_basic_integration.py:12-154 ---------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("t_span", [[5, 9], [5, 1]])
def test_integration_rational(solve_ivp, method, t_span):
    rtol, atol = 1e-3, 1e-6
    res = rational(solve_ivp, t_span, rtol=rtol, atol=atol, method=method, dense_output=True)
    assert res.t[0] == t_span[0] and res.t[-1] == t_span[1]
    assert res.success and res.status == 0
    assert res.nfev > 0 and (res.njev > 0) == (method in ("Radau", "BDF")) and (res.nlu > 0) == (method in ("Radau", "BDF"))
    e = compute_error(res.y, sol_rational(res.t), rtol, atol)
    assert np.all(e < 5)
    # dense output at the step midpoints (tests/test_ivp.py:201-214 does the same with its tolerance of 5)
    tc = (res.t[:-1] + res.t[1:]) / 2
    e = compute_error(res.sol(tc), sol_rational(tc), rtol, atol)
    assert np.all(e < 5)
    assert res.sol.t_min == min(t_span) and res.sol.t_max == max(t_span)


def test_integration_analytic_jacobian(solve_ivp):
    """tests/test_ivp.py:172-241 passes jac_rational to the implicit methods."""
    for method in ("Radau", "BDF"):
        res = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, method=method, jac=True)
        assert res.success and res.njev > 0
        assert np.all(compute_error(res.y, sol_rational(res.t), 1e-3, 1e-6) < 5)


@pytest.mark.parametrize("method", ["Radau", "BDF"])
@pytest.mark.parametrize("jac", [True, None])
def test_integration_const_jac(solve_ivp, method, jac):           # tests/test_ivp.py:272-317, tests/test_stiff.py:14-94
    rtol, atol, t_span = 1e-3, 1e-6, [0, 2]
    sol_linear = lambda t: np.vstack((-5 * np.sin(2 * t), 2 * np.cos(2 * t) + np.sin(2 * t)))   # test_helpers.py:19-21
    res = solve_ivp(SRC_LINEAR, t_span, [0, 2], rtol=rtol, atol=atol, method=method, dense_output=True, jac=jac)
    assert res.t[0] == t_span[0] and res.t_events is None and res.y_events is None
    assert res.success and res.status == 0
    assert res.nfev < 100
    assert np.all(compute_error(res.y, sol_linear(res.t), rtol, atol) < 10)
    tc = np.linspace(*t_span)
    e = compute_error(res.sol(tc), sol_linear(tc), rtol, atol)
    assert np.all(e < (60 if method == "BDF" else 15))            # the reference's own relaxed BDF bound
    np.testing.assert_allclose(res.sol(res.t), res.y, rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_integration_stiff(solve_ivp, method):                    # tests/test_ivp.py:319-343, tests/test_stiff.py:97-143
    res = solve_ivp("robertson", [0, 1e8], [1e4, 0, 0], args=(0.04, 1e4, 3e7), rtol=1e-6, atol=1e-6, method=method)
    assert res.success and res.nfev < 5000 and res.njev < 200
    np.testing.assert_allclose(res.y[:, -1].sum(), 1e4, rtol=1e-4)   # (mass conservation; stronger than the reference)


# ---- tests/test_t_eval.py:10-160 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("t_span,t_eval,check", [
    ([5, 9], np.linspace(5, 9, 10), True),                       # test_t_eval_forward
    ([5, 1], np.linspace(5, 1, 10), False),                      # test_t_eval_backward
    ([5, 9], [5, 5.01, 7, 8, 8.01, 9], True),                    # test_t_eval_irregular_forward
    ([5, 1], [5, 4.99, 3, 1.5, 1.1, 1.01, 1], False),            # test_t_eval_irregular_backward
    ([5, 9], [5.01, 7, 8, 8.01], True),                          # test_t_eval_interior_forward
    ([5, 1], [4.99, 3, 1.5, 1.1, 1.01], False),                  # test_t_eval_interior_backward
])
def test_t_eval(solve_ivp, t_span, t_eval, check):
    rtol, atol = 1e-3, 1e-6
    res = rational(solve_ivp, t_span, rtol=rtol, atol=atol, t_eval=t_eval)
    assert np.array_equal(res.t, np.asarray(t_eval, dtype=float))
    assert res.success and res.status == 0
    if check:
        assert np.all(compute_error(res.y, sol_rational(res.t), rtol, atol) < 5)


def test_t_eval_dense_output(solve_ivp):                          # tests/test_t_eval.py:110-131
    t_eval = np.linspace(5, 9, 10)
    res = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, t_eval=t_eval)
    res_d = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, t_eval=t_eval, dense_output=True)
    assert np.array_equal(res.t, t_eval) and res.success and res.status == 0
    assert np.array_equal(res.t, res_d.t) and np.array_equal(res.y, res_d.y)
    assert res_d.success and res_d.status == 0


@pytest.mark.parametrize("method", METHODS)
def test_t_eval_early_event(solve_ivp, method):                   # tests/test_t_eval.py:134-160
    res = rational(solve_ivp, [5, 9], events=[Ev(), Ev(), Ev(terminal=True)], p=7.0, rtol=1e-3, atol=1e-6,
                   method=method, t_eval=np.linspace(7.5, 9, 16), jac=True)
    assert res.success and res.status == 1
    assert res.t_events[2].size == 1
    np.testing.assert_allclose(res.t_events[2][0], 7, rtol=1e-10, atol=1e-10)
    # solout.rs:315-324: a terminal event appends its own point although no t_eval sample was reached
    assert res.t.size == 1 and res.t[0] == res.t_events[2][0]


# ---- tests/test_events.py:10-162 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
def test_events(solve_ivp, method):                               # test_events_RK23 ... test_events_BDF
    res = rational(solve_ivp, [5, 8], events=[Ev(), Ev(), Ev()], p=100.0, method=method)
    assert res.status == 0
    assert len(res.t_events[0]) == 1 and len(res.t_events[1]) == 1 and len(res.t_events[2]) == 0
    assert 5.3 < res.t_events[0][0] < 5.7
    assert 7.3 < res.t_events[1][0] < 7.7
    assert res.y_events[0].shape == (1, 2) and res.y_events[1].shape == (1, 2)
    # the event functions vanish at the located states (tests/test_ivp.py:386-392 checks this with 1e-6... here 1e-9)
    y = res.y_events[0][0]
    assert abs(y[0] - y[1] ** 0.7) < 1e-9
    y = res.y_events[1][0]
    assert abs(y[1] ** 0.6 - y[0]) < 1e-9


@pytest.mark.parametrize("method", METHODS)
def test_events_directions_and_terminal(solve_ivp, method):       # tests/test_ivp.py:345-420
    ev1 = lambda y: y[0] - y[1] ** 0.7
    ev2 = lambda y: y[1] ** 0.6 - y[0]
    res = rational(solve_ivp, [5, 8], events=[Ev(direction=1), Ev(direction=1), Ev()], p=100.0, method=method)
    assert res.status == 0 and len(res.t_events[0]) == 1 and len(res.t_events[1]) == 0
    assert 5.3 < res.t_events[0][0] < 5.7 and res.y_events[0].shape == (1, 2)
    assert np.isclose(ev1(res.y_events[0][0]), 0, atol=1e-5)
    res = rational(solve_ivp, [5, 8], events=[Ev(direction=-1), Ev(direction=-1), Ev()], p=100.0, method=method)
    assert res.status == 0 and len(res.t_events[0]) == 0 and len(res.t_events[1]) == 1
    assert 7.3 < res.t_events[1][0] < 7.7 and res.y_events[1].shape == (1, 2)
    assert np.isclose(ev2(res.y_events[1][0]), 0, atol=1e-5)
    res = rational(solve_ivp, [5, 8], events=[Ev(), Ev(), Ev(terminal=True)], p=7.4, method=method, dense_output=True)
    assert res.status == 1
    assert len(res.t_events[0]) == 1 and len(res.t_events[1]) == 0 and len(res.t_events[2]) == 1
    assert 5.3 < res.t_events[0][0] < 5.7 and 7.3 < res.t_events[2][0] < 7.5
    assert res.y_events[0].shape == (1, 2) and res.y_events[2].shape == (1, 2)
    assert np.isclose(ev1(res.y_events[0][0]), 0, atol=1e-5) and np.isclose(res.t_events[2][0] - 7.4, 0, atol=1e-5)
    # tests/test_ivp.py:429-438: termination by an event does not break the interpolants; y_event matches the solution
    assert res.t[-1] == res.t_events[2][0]
    tc = np.linspace(res.t[0], res.t[-1])
    assert np.all(compute_error(res.sol(tc), sol_rational(tc), 1e-3, 1e-6) < 5)
    assert np.allclose(sol_rational(res.t_events[0][0]), res.y_events[0][0], rtol=1e-3, atol=1e-6)
    # tests/test_ivp.py:440-492: backward direction (crossings seen with the opposite sign of time)
    back = dict(method=method, p=100.0)
    res = solve_ivp(SRC_RATIONAL_EV, [8, 5], [4 / 9, 20 / 81], args=(100.0,), events=[Ev(), Ev(), Ev()], method=method)
    assert res.status == 0 and len(res.t_events[0]) == 1 and len(res.t_events[1]) == 1
    assert 5.3 < res.t_events[0][0] < 5.7 and 7.3 < res.t_events[1][0] < 7.7
    assert np.isclose(ev1(res.y_events[0][0]), 0, atol=1e-5) and np.isclose(ev2(res.y_events[1][0]), 0, atol=1e-5)
    res = solve_ivp(SRC_RATIONAL_EV, [8, 5], [4 / 9, 20 / 81], args=(100.0,), events=[Ev(direction=-1), Ev(direction=-1), Ev()],
                    method=method)
    assert res.status == 0 and len(res.t_events[0]) == 1 and len(res.t_events[1]) == 0
    res = solve_ivp(SRC_RATIONAL_EV, [8, 5], [4 / 9, 20 / 81], args=(100.0,), events=[Ev(direction=1), Ev(direction=1), Ev()],
                    method=method)
    assert res.status == 0 and len(res.t_events[0]) == 0 and len(res.t_events[1]) == 1
    del back


def test_terminal_event(solve_ivp):                               # tests/test_events.py:99-113
    res = rational(solve_ivp, [5, 8], events=[Ev(), Ev(), Ev(terminal=True)], p=7.4, method="RK45", dense_output=True)
    assert res.status == 1
    assert len(res.t_events[2]) == 1 and 7.3 < res.t_events[2][0] < 7.5
    assert res.t[-1] == res.t_events[2][0]                        # the terminal point closes the solution
    assert len(res.t_events[0]) == 1 and len(res.t_events[1]) == 0   # event 2 (t ~ 7.5) lies after the stop


def test_event_directions(solve_ivp):                             # tests/test_events.py:116-144
    res = rational(solve_ivp, [5, 8], events=[Ev(direction=1), Ev(), Ev()], p=100.0, method="RK45")
    assert res.status == 0 and len(res.t_events[0]) == 1 and 5.3 < res.t_events[0][0] < 5.7
    res = rational(solve_ivp, [5, 8], events=[Ev(direction=-1), Ev(), Ev()], p=100.0, method="RK45")
    assert res.status == 0 and len(res.t_events[0]) == 0
    # tests/test_ivp.py:421-437: the second event only counts downward crossings of y1^0.6 - y0 ... it rises
    res = rational(solve_ivp, [5, 8], events=[Ev(direction=1), Ev(direction=-1), Ev()], p=100.0, method="RK45")
    assert len(res.t_events[0]) == 1 and len(res.t_events[1]) == 1


def test_duplicate_timestamps(solve_ivp):                         # tests/test_events.py:147-162 (t_span = [0, inf])
    sol = solve_ivp("cannon", [0, np.inf], [0, 0.01], max_step=0.05 * 0.001 / 9.80665,
                    events=Ev(terminal=True, direction=-1), dense_output=True)
    np.testing.assert_allclose(sol.sol(0.01), np.asarray([-0.00039033, -0.08806632]), rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(sol.t_events[0], np.asarray([0.00203943]), rtol=1e-5, atol=1e-8)
    assert sol.success and sol.status == 1


# ---- tests/test_step_control.py:9-177 ----------------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("t_span", [[5, 9], [5, 1]])
def test_max_step(solve_ivp, method, t_span):                     # test_max_step_forward / _backward
    rtol, atol = 1e-3, 1e-6
    res = rational(solve_ivp, t_span, rtol=rtol, max_step=0.5, atol=atol, method=method, dense_output=True)
    assert res.t[0] == t_span[0] and res.t[-1] == t_span[-1]
    assert np.all(np.abs(np.diff(res.t)) <= 0.5 + 1e-15)
    assert res.success and res.status == 0
    if t_span[1] > t_span[0]:
        assert np.all(compute_error(res.y, sol_rational(res.t), rtol, atol) < 5)


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("t_span", [[5, 9], [5, 1]])
def test_first_step(solve_ivp, method, t_span):                   # test_first_step_forward / _backward
    res = rational(solve_ivp, t_span, rtol=1e-3, max_step=0.5, atol=1e-6, method=method, dense_output=True, first_step=0.1)
    assert res.t[0] == t_span[0] and res.t[-1] == t_span[-1]
    np.testing.assert_allclose(0.1, np.abs(res.t[1] - 5))
    assert res.success and res.status == 0


@pytest.mark.parametrize("method", METHODS)
def test_max_steps(solve_ivp, method):                            # test_max_steps_parameter / _large_value
    res = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, method=method, max_steps=1)
    assert not res.success and res.status == -1 and "NeedLargerNMax" in res.message
    res = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, method=method, max_steps=1_000_000)
    assert res.success and res.status == 0


@pytest.mark.parametrize("method", METHODS)
def test_default_max_steps_is_unlimited(solve_ivp, method):       # tests/test_step_control.py:133-162
    res = solve_ivp(SRC_SCALE1, [0, 100000], [1.0], args=(-0.001,), method=method, rtol=1e-8, atol=1e-10, max_out=400000)
    assert res.success and res.status == 0, res.message
    assert res.t[-1] == 100000
    # (stronger than the reference: the decay is integrated correctly over the 100 time constants)
    np.testing.assert_allclose(res.y[0, -1], np.exp(-100.0), rtol=0, atol=1e-8)


@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_min_step(solve_ivp, method):                             # tests/test_step_control.py:165-177
    res = rational(solve_ivp, [5, 9], rtol=1e-3, atol=1e-6, method=method, min_step=1e-10)
    assert res.success and res.status == 0, res.message


# ---- tests/test_edge_cases.py:9-121 ------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
def test_no_integration(solve_ivp, method):                       # tests/test_edge_cases.py:9-16
    sol = solve_ivp(SRC_SCALE2, [4, 4], [2, 3], args=(-1.0,), method=method, dense_output=True)
    assert np.array_equal(sol.sol(4), [2, 3])
    assert np.array_equal(sol.sol([4, 5, 6]), [[2, 2, 2], [3, 3, 3]])    # the constant segment, extrapolated


@pytest.mark.parametrize("method", METHODS)
def test_integration_zero_rhs(solve_ivp, method):                 # tests/test_edge_cases.py:35-42
    res = solve_ivp("zero3", [0, 10], np.ones(3), method=method)
    assert res.success and res.status == 0
    np.testing.assert_allclose(res.y, 1.0, rtol=1e-15)


@pytest.mark.parametrize("method", METHODS)
def test_zero_interval(solve_ivp, method):                        # tests/test_edge_cases.py:45-53
    res = solve_ivp(SRC_SCALE1, (0.0, 0.0), np.array([1.0]), args=(2.0,), method=method)
    assert res.success
    np.testing.assert_allclose(res.y[0, -1], 1.0)


@pytest.mark.parametrize("method", METHODS)
def test_tbound_respected_small_interval(solve_ivp, method):      # tests/test_edge_cases.py:56-67 (gh-17341)
    SMALL = 1e-4
    res = solve_ivp(SRC_SCALE1, (0.0, SMALL), np.array([1.0]), args=(2.0,), method=method)
    assert res.success
    assert np.all(res.t <= SMALL) and res.t[-1] == SMALL
    np.testing.assert_allclose(res.y[0, -1], np.exp(2 * SMALL), rtol=1e-6)


@pytest.mark.parametrize("method", METHODS)
def test_tbound_respected_larger_interval(solve_ivp, method):     # tests/test_edge_cases.py:70-98 (gh-8848)
    res = solve_ivp(SRC_RADIAL, (-17, 2), np.array([1.0, -11.0]), max_step=0.03, t_eval=None, atol=1e-8, rtol=1e-5,
                    method=method)
    assert res.success
    assert res.t[0] == -17 and res.t[-1] == 2 and np.all(res.t <= 2)


@pytest.mark.parametrize("method", METHODS)
def test_tbound_respected_oscillator(solve_ivp, method):          # tests/test_edge_cases.py:101-121 (gh-9198)
    init = np.array([134.08298555, 138.82348612, 100.0, 0.0])
    res = solve_ivp(SRC_CONST, (100.0, 200.0), init, dense_output=True, max_step=100.0, method=method)
    assert res.success and np.all(res.t <= 200.0)
    np.testing.assert_allclose(res.y[:, -1], init + 100.0 * np.array([1.73307544e-02, 6.49376470e-06, 0, 0]), rtol=1e-12)


# ---- tests/test_args.py:8-95 -------------------------------------------------------------------------------------------
def test_args_with_events(solve_ivp):                             # tests/test_args.py:8-66
    omega, k, tfinal, zfinal = 2, 4, 5, 0.99
    z0 = np.exp(-k * tfinal) / ((1 - zfinal) / zfinal + np.exp(-k * tfinal))
    sol = solve_ivp(SRC_SYS3, [0, 2 * tfinal], [0, -1, z0], events=[Ev(direction=-1), Ev(direction=1), Ev(terminal=True)],
                    dense_output=True, args=(omega, k, zfinal), method="Radau", jac=True, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(sol.t_events[0], [0.5 * np.pi, 1.5 * np.pi], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(sol.t_events[1], [0.25 * np.pi, 1.25 * np.pi], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(sol.t_events[2], [tfinal], rtol=1e-5, atol=1e-5)
    t = np.linspace(0, sol.t_events[2][0], 250)
    w = sol.sol(t)
    np.testing.assert_allclose(w[0], np.sin(omega * t), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(w[1], -np.cos(omega * t), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(w[2], 1 / (((1 - z0) / z0) * np.exp(-k * t) + 1), rtol=1e-4, atol=1e-6)


def test_args_single_value(solve_ivp):                            # tests/test_args.py:69-76
    sol = solve_ivp(SRC_SCALE1, (0, 0.1), [1], args=(-1,))
    np.testing.assert_allclose(sol.y[0, -1], np.exp(-0.1), rtol=1e-3)      # default rtol = 1e-3


def test_array_rtol(solve_ivp):                                   # tests/test_args.py:79-95 (gh-15482)
    sol = solve_ivp("exp2", (0, 1), [1.0, 1.0], rtol=[1e-1, 1e-1])
    err1 = np.abs(np.linalg.norm(sol.y[:, -1] - np.exp(1)))
    sol = solve_ivp("exp2", (0, 1), [1.0, 1.0], rtol=[1e-1, 1e-16])
    err2 = np.abs(np.linalg.norm(sol.y[:, -1] - np.exp(1)))
    assert err2 < err1
