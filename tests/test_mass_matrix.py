"""RADAU with a constant mass matrix, M y' = f(t, y) (Options.mass_storage = Full + IVP::mass,
reference src/methods/radau.rs:283,358-359,375-386,525-539,626-634) and index-2 / index-3 error scaling
(Options.nind1..3, radau.rs:210-245,434-445), as solve_ivp wires them (src/solve/solve_ivp.rs:246-258).

The reference's own tests never set a mass matrix, so the oracle is pinned here against analytic solutions
(M y' = A y  =>  y = expm(M^-1 A t) y0) and against the identity-mass formulation of the same physics (Robertson as
an index-1 DAE vs Robertson as an ODE); the CUDA path is then compared with the oracle bit for bit (strict build).
"""
import numpy as np
import pytest

import ivp_b200 as ib
from ivp_b200 import Method, Options, Status
from ivp_b200.api import IVPB_FLAG_FAST_FP, PROBLEMS

M3 = np.array([[2.0, 0.5, 0.0], [0.25, 1.5, -0.5], [0.0, 0.75, 3.0]])
A3 = np.array([[-2.0, 1.0, 0.0], [1.0, -2.0, 1.0], [0.0, 1.0, -2.0]])
M4 = np.array([[(2.0 + 0.5 * i) if i == j else 0.25 / (1.0 + abs(i - j)) * (-1.0 if (i + j) % 2 else 1.0)
                for j in range(4)] for i in range(4)])
A4 = -2.0 * np.eye(4) + np.eye(4, k=1) + np.eye(4, k=-1)
RATES = [0.04, 1e4, 3e7]

SRC_MASS4 = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) {
  d[0] = p[0] * (-2.0 * y[0] + y[1]);
  d[1] = p[0] * (y[0] - 2.0 * y[1] + y[2]);
  d[2] = p[0] * (y[1] - 2.0 * y[2] + y[3]);
  d[3] = p[0] * (y[2] - 2.0 * y[3]);
}
__device__ void ivp_mass(const double* p, double* M) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      M[i * 4 + j] = (i == j) ? 2.0 + 0.5 * (double)i : 0.25 / (1.0 + (double)(i > j ? i - j : j - i)) * ((i + j) % 2 ? -1.0 : 1.0);
}
"""


def expm_sol(M, A, k, t, y0):
    sl = pytest.importorskip("scipy.linalg")
    return sl.expm(np.linalg.solve(M, k * A) * t) @ y0


def mass_ensemble(N, n, seed=7):
    rng = np.random.default_rng(seed)
    return rng.uniform(-2.0, 2.0, (N, n)), rng.uniform(0.5, 40.0, (N, 1))


def robertson_ensemble(N, seed=11):
    rng = np.random.default_rng(seed)
    x = 1.0 - rng.uniform(0.0, 0.2, N)
    y0 = np.stack([x, np.zeros(N), 1.0 - x], axis=1)          # consistent initial values: x + y + z = 1
    par = np.tile(RATES, (N, 1)) * (1.0 + rng.uniform(-0.1, 0.1, (N, 3)))
    return y0, par


# ---- the oracle against analytic / identity-mass answers (CPU) ---------------------------------------------------------
@pytest.mark.parametrize("jac_mode", [0, 1])
def test_oracle_full_mass_linear_matches_matrix_exponential(oracle, jac_mode):
    y0, par = mass_ensemble(6, 3)
    opts = Options(method=Method.RADAU, rtol=1e-9, atol=1e-12, mass_storage="Full", jac_mode=jac_mode)
    o = oracle.solve_batch(PROBLEMS["mass_linear3"], 0.0, 1.5, y0, par, opts)
    assert np.all(o.status == Status.Success)
    for i in range(6):
        np.testing.assert_allclose(o.y_final[i], expm_sol(M3, A3, par[i, 0], 1.5, y0[i]), rtol=1e-6, atol=1e-9)
    y0, par = mass_ensemble(4, 4)
    o = oracle.solve_batch(107, 0.0, 1.0, y0, par, Options(method=Method.RADAU, rtol=1e-9, atol=1e-12, mass_storage="Full"))
    for i in range(4):
        np.testing.assert_allclose(o.y_final[i], expm_sol(M4, A4, par[i, 0], 1.0, y0[i]), rtol=1e-6, atol=1e-9)


def test_oracle_robertson_dae_matches_ode_form(oracle):
    y0, par = robertson_ensemble(8)
    te = np.array([1e-3, 1.0, 1e2, 1e4, 1e6])
    od = Options(method=Method.RADAU, rtol=1e-8, atol=1e-12, mass_storage="Full", t_eval=te)
    oo = Options(method=Method.RADAU, rtol=1e-8, atol=1e-12, t_eval=te)
    d = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e6, y0, par, od)
    o = oracle.solve_batch(PROBLEMS["robertson"], 0.0, 1e6, y0, par, oo)
    assert np.all(d.status == Status.Success) and np.all(d.n_out == te.size)
    np.testing.assert_allclose(d.y_out, o.y_out, rtol=1e-5, atol=1e-10)
    assert np.abs(d.y_out[:, :te.size].sum(axis=2) - 1.0).max() < 1e-10      # the algebraic row holds at every sample
    assert np.all(d.nlu > 0) and np.all(d.njev > 0)


def test_oracle_dae_partition_rules(oracle):
    """radau.rs:210-245: omitted nind1 is inferred, explicit sums must equal n; index-2 scaling changes the error norm."""
    y0, par = robertson_ensemble(2)
    base = dict(method=Method.RADAU, rtol=1e-6, atol=1e-10, mass_storage="Full")
    a = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, Options(**base))
    b = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, Options(nind1=3, **base))
    assert np.array_equal(a.counters, b.counters) and np.array_equal(a.y_final, b.y_final)
    c = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, Options(nind2=1, **base))
    d = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, Options(nind1=2, nind2=1, **base))
    assert np.array_equal(c.counters, d.counters) and not np.array_equal(a.counters, c.counters)
    np.testing.assert_allclose(c.y_final, a.y_final, rtol=1e-4, atol=1e-9)
    for bad in (dict(nind1=1, nind2=1), dict(nind2=2, nind3=2), dict(nind1=4)):
        with pytest.raises(RuntimeError, match="DAE partition"):
            oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, Options(**bad, **base))
    with pytest.raises(RuntimeError, match="no mass matrix"):
        oracle.solve_batch(PROBLEMS["robertson"], 0.0, 1e3, y0, par, Options(**base))


# ---- the CUDA path against the oracle (GPU) ----------------------------------------------------------------------------
def exact(g, o, fields=("status", "counters", "t_final", "y_final", "h_next")):
    for f in fields:
        a, b = getattr(g, f), getattr(o, f)
        assert np.array_equal(a, b, equal_nan=(a.dtype.kind == "f")), f


@pytest.mark.gpu
@pytest.mark.parametrize("jac_mode", [0, 1])
@pytest.mark.parametrize("wl", ["robertson_dae", "mass_linear3"])
def test_radau_mass_matrix_strict_bit_exact(oracle, wl, jac_mode):
    N = 3000
    if wl == "robertson_dae":
        (y0, par), tf, te = robertson_ensemble(N), 1e5, np.array([1e-2, 1.0, 1e2, 1e5])
    else:
        (y0, par), tf, te = mass_ensemble(N, 3), 2.0, np.linspace(0.0, 2.0, 9)
    for extra in ({}, {"t_eval": te}, {"dense_output": True, "max_segments": 600, "max_out": 600}):
        opts = Options(method=Method.RADAU, rtol=1e-7, atol=1e-11, mass_storage="Full", jac_mode=jac_mode, **extra)
        g = ib.solve_ivp_batch(wl, 0.0, tf, y0, par, opts)
        o = oracle.solve_batch(PROBLEMS[wl], 0.0, tf, y0, par, opts, nthreads=8)
        assert np.all(g.status == Status.Success)
        exact(g, o)
        if extra:
            assert np.array_equal(g.n_out, o.n_out) and np.array_equal(g.t_out, o.t_out) and np.array_equal(g.y_out, o.y_out)
    if wl == "mass_linear3":
        for i in (0, N - 1):
            np.testing.assert_allclose(g.y_final[i], expm_sol(M3, A3, par[i, 0], tf, y0[i]), rtol=1e-5, atol=1e-8)
    else:
        assert np.abs(g.y_final.sum(axis=1) - 1.0).max() < 1e-9
    # the FMA build stays inside the north-star tolerance of the oracle
    opts = Options(method=Method.RADAU, rtol=1e-7, atol=1e-11, mass_storage="Full", jac_mode=jac_mode, flags=IVPB_FLAG_FAST_FP)
    f = ib.solve_ivp_batch(wl, 0.0, tf, y0, par, opts)
    assert np.array_equal(f.status, o.status)
    tol = np.maximum(10 * 1e-7 * np.abs(o.y_final), 10 * 1e-11)
    assert (np.abs(f.y_final - o.y_final) <= tol).all(axis=1).mean() >= 0.99


@pytest.mark.gpu
def test_radau_dae_index_scaling_and_config_errors(oracle):
    y0, par = robertson_ensemble(500)
    for nind in (dict(nind2=1), dict(nind1=1, nind2=1, nind3=1), dict(nind3=1)):
        opts = Options(method=Method.RADAU, rtol=1e-6, atol=1e-10, mass_storage="Full", **nind)
        g = ib.solve_ivp_batch("robertson_dae", 0.0, 1e3, y0, par, opts)
        o = oracle.solve_batch(PROBLEMS["robertson_dae"], 0.0, 1e3, y0, par, opts, nthreads=8)
        exact(g, o)
    base = dict(method=Method.RADAU, mass_storage="Full")
    for bad in (dict(nind1=1, nind2=1), dict(nind2=2, nind3=2), dict(nind1=4)):        # ConfigError::InvalidDAEPartition
        with pytest.raises(ib.ConfigError, match="DAE partition"):
            ib.solve_ivp_batch("robertson_dae", 0.0, 1.0, y0, par, Options(**bad, **base))
    with pytest.raises(ib.ConfigError, match="no mass matrix"):
        ib.solve_ivp_batch("robertson", 0.0, 1.0, y0, par, Options(**base))
    with pytest.raises(ib.ConfigError, match="mass_storage = Full"):
        ib.solve_ivp_batch("robertson_dae", 0.0, 1.0, y0, par, Options(method=Method.RADAU))
    # every other method ignores the mass matrix, like the reference (solve_ivp.rs passes it to RADAU only)
    g = ib.solve_ivp_batch("mass_linear3", 0.0, 0.5, *mass_ensemble(64, 3), Options(method=Method.BDF, rtol=1e-6, atol=1e-9))
    o = oracle.solve_batch(PROBLEMS["mass_linear3"], 0.0, 0.5, *mass_ensemble(64, 3), Options(method=Method.BDF, rtol=1e-6, atol=1e-9))
    exact(g, o)


@pytest.mark.gpu
def test_radau_mass_matrix_user_problem_shared_memory_matrices(oracle):
    """n = 4: J, E1, E2 and M live in shared memory ([element][thread]); the problem comes in as CUDA C with `ivp_mass`."""
    from ivp_b200 import api, scipy_api
    y0, par = mass_ensemble(700, 4)
    user = api.Problem.from_cuda_source(SRC_MASS4, n=4, p=1, has_mass=True)
    opts = Options(method=Method.RADAU, rtol=1e-8, atol=1e-11, mass_storage="Full", t_eval=np.linspace(0.0, 1.0, 5))
    g = ib.solve_ivp_batch(user, 0.0, 1.0, y0, par, opts)
    o = oracle.solve_batch(107, 0.0, 1.0, y0, par, opts, nthreads=8)
    assert np.all(g.status == Status.Success) and np.array_equal(g.status, o.status)
    assert (g.counters == o.counters).all(axis=1).mean() >= 0.99          # separate compilations: not a bit-level claim
    np.testing.assert_allclose(g.y_out, o.y_out, rtol=1e-6, atol=1e-9)
    for i in (0, 699):
        np.testing.assert_allclose(g.y_final[i], expm_sol(M4, A4, par[i, 0], 1.0, y0[i]), rtol=1e-6, atol=1e-9)
    # through the SciPy-style front end: mass_storage / nind are plain options there
    r = scipy_api.solve_ivp(SRC_MASS4, (0.0, 1.0), y0[0], method="Radau", args=(par[0, 0],), rtol=1e-8, atol=1e-11,
                            mass_storage="Full")
    np.testing.assert_allclose(r.y[:, -1], expm_sol(M4, A4, par[0, 0], 1.0, y0[0]), rtol=1e-6, atol=1e-9)


# ---- n > 8: the warp-per-trajectory RADAU kernel (mass matrix next to the Jacobian in the warp's global slot) ----------
M12 = np.array([[(2.0 + 0.25 * i) if i == j else 0.25 / (1.0 + abs(i - j)) * (-1.0 if (i + j) % 2 else 1.0)
                 for j in range(12)] for i in range(12)])
A12 = -2.0 * np.eye(12) + np.eye(12, k=1) + np.eye(12, k=-1)

SRC_MASS12 = """
__device__ void ivp_ode(double t, const double* y, const double* p, double* d) {
  for (int i = 0; i < 12; ++i) {
    const double lo = i > 0 ? y[i - 1] : 0.0, hi = i < 11 ? y[i + 1] : 0.0;
    d[i] = p[0] * (lo - 2.0 * y[i] + hi);
  }
}
__device__ void ivp_mass(const double* p, double* M) {
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j)
      M[i * 12 + j] = (i == j) ? 2.0 + 0.25 * (double)i : 0.25 / (1.0 + (double)(i > j ? i - j : j - i)) * ((i + j) % 2 ? -1.0 : 1.0);
}
"""


def test_oracle_mass_linear12_matches_matrix_exponential(oracle):
    y0, par = mass_ensemble(4, 12)
    o = oracle.solve_batch(109, 0.0, 1.0, y0, par, Options(method=Method.RADAU, rtol=1e-9, atol=1e-12, mass_storage="Full"))
    assert np.all(o.status == Status.Success)
    for i in range(4):
        np.testing.assert_allclose(o.y_final[i], expm_sol(M12, A12, par[i, 0], 1.0, y0[i]), rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
def test_radau_mass_matrix_warp_cooperative_kernel(oracle):
    """n = 12 > 8: RADAU runs one trajectory per warp; M y' = f with a full constant mass matrix and the index-2 / index-3
    error scaling (radau.rs:283,358-386,433-445,525-539,626-634) against the oracle's restatement (problem 109)."""
    from ivp_b200 import api
    from ivp_b200.api import IVPB_FLAG_STRICT_FP
    y0, par = mass_ensemble(300, 12)
    user = api.Problem.from_cuda_source(SRC_MASS12, n=12, p=1, has_mass=True)
    te = np.linspace(0.0, 1.0, 5)
    for extra in ({}, dict(nind2=2), dict(nind1=8, nind2=2, nind3=2)):
        opts = Options(method=Method.RADAU, rtol=1e-8, atol=1e-11, mass_storage="Full", t_eval=te, flags=IVPB_FLAG_STRICT_FP, **extra)
        g = ib.solve_ivp_batch(user, 0.0, 1.0, y0, par, opts)
        o = oracle.solve_batch(109, 0.0, 1.0, y0, par, opts, nthreads=8)
        assert np.all(g.status == Status.Success) and np.array_equal(g.status, o.status)
        assert (g.counters == o.counters).all(axis=1).mean() >= 0.99          # separate compilations of the RHS
        np.testing.assert_allclose(g.y_out, o.y_out, rtol=1e-6, atol=1e-9)
        if not extra:
            for i in (0, 299):
                np.testing.assert_allclose(g.y_final[i], expm_sol(M12, A12, par[i, 0], 1.0, y0[i]), rtol=1e-6, atol=1e-9)
    with pytest.raises(ib.ConfigError, match="mass_storage = Full"):
        ib.solve_ivp_batch(user, 0.0, 1.0, y0, par, Options(method=Method.RADAU))
