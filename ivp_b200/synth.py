"""Synthetic ensembles for the BASELINE.json configurations.

Counter-based generator so host, C++ and CUDA can produce identical bits:
    u(i, j) = (splitmix64(seed + i * n_fields + j) >> 11) * 2**-53,  seed = 0x5EED1234
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EED1234
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(N: int, n_fields: int, seed: int = SEED, offset: int = 0) -> np.ndarray:
    """[N, n_fields] doubles in [0, 1); `offset` = global index of row 0 (for sharding across ranks)."""
    i = (np.arange(N, dtype=np.uint64) + np.uint64(offset))[:, None]
    j = np.arange(n_fields, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        ctr = np.uint64(seed) + i * np.uint64(n_fields) + j
    return (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def ensemble(name: str, N: int, offset: int = 0):
    """Return (problem_name, y0[N,n], params[N,p] or None, t0, tf) for a named workload."""
    if name == "vdp":            # north star: VdP mu=1, y0 ~ U(-3,3)^2
        u = uniform(N, 2, offset=offset)
        return "vdp_mu", -3.0 + 6.0 * u, np.ones((N, 1)), 0.0, 100.0
    if name == "decay":          # cfg 2: y0 ~ U(1,10), k ~ U(0.1,1)
        u = uniform(N, 2, offset=offset)
        return "decay", 1.0 + 9.0 * u[:, :1], 0.1 + 0.9 * u[:, 1:2], 0.0, 10.0
    if name == "lorenz":         # cfg 2: y0 = 1 + U(-.5,.5)^3, sigma=10, rho ~ U(20,35), beta=8/3
        u = uniform(N, 4, offset=offset)
        par = np.stack([np.full(N, 10.0), 20.0 + 15.0 * u[:, 3], np.full(N, 8.0 / 3.0)], axis=1)
        return "lorenz", 1.0 + (u[:, :3] - 0.5), par, 0.0, 10.0
    if name == "cr3bp":          # cfg 3: perturbed Arenstorf orbit
        u = uniform(N, 4, offset=offset)
        y0 = np.zeros((N, 6))
        y0[:, 0] = 0.994 * (1.0 + 1e-3 * (2.0 * u[:, 0] - 1.0))
        y0[:, 4] = -2.00158510637908252240537862224 * (1.0 + 1e-3 * (2.0 * u[:, 1] - 1.0))
        y0[:, 2] = 1e-3 * (2.0 * u[:, 2] - 1.0)
        y0[:, 5] = 1e-3 * (2.0 * u[:, 3] - 1.0)
        return "cr3bp", y0, np.full((N, 1), 0.012277471), 0.0, 17.0652165601579625588917206249
    if name == "ball":           # cfg 4: h0 ~ U(1,20), v0 ~ U(-5,10), g=9.81, drag ~ U(0,0.05)
        u = uniform(N, 3, offset=offset)
        y0 = np.stack([1.0 + 19.0 * u[:, 0], -5.0 + 15.0 * u[:, 1]], axis=1)
        par = np.stack([np.full(N, 9.81), 0.05 * u[:, 2]], axis=1)
        return "bouncing_ball", y0, par, 0.0, 10.0
    if name == "robertson":      # cfg 5
        u = uniform(N, 1, offset=offset)
        y0 = np.zeros((N, 3))
        y0[:, 0] = 1e4 * (1.0 + 0.1 * (2.0 * u[:, 0] - 1.0))
        par = np.tile(np.array([[0.04, 1e4, 3e7]]), (N, 1))
        return "robertson", y0, par, 0.0, 1e8
    if name == "ball_bounce":    # cfg 4's ball, bouncing inside the solve (restitution ~ U(0.6, 0.9))
        u = uniform(N, 4, offset=offset)
        y0 = np.stack([1.0 + 19.0 * u[:, 0], -5.0 + 15.0 * u[:, 1]], axis=1)
        par = np.stack([np.full(N, 9.81), 0.05 * u[:, 2], 0.6 + 0.3 * u[:, 3]], axis=1)
        return "ball_bounce", y0, par, 0.0, 15.0
    if name == "mass_linear3":   # M y' = k A y with a full constant M, y0 ~ U(-2, 2)^3, k ~ U(0.5, 40)
        u = uniform(N, 4, offset=offset)
        return "mass_linear3", 4.0 * u[:, :3] - 2.0, 0.5 + 39.5 * u[:, 3:4], 0.0, 2.0
    if name == "robertson_dae":  # Robertson as an index-1 DAE (mass matrix diag(1, 1, 0)), x + y + z = 1
        u = uniform(N, 1, offset=offset)
        y0 = np.zeros((N, 3))
        y0[:, 0] = 1.0 - 0.2 * u[:, 0]
        y0[:, 2] = 1.0 - y0[:, 0]
        par = np.tile(np.array([[0.04, 1e4, 3e7]]), (N, 1))
        return "robertson_dae", y0, par, 0.0, 1e8
    if name == "vdp_stiff":      # cfg 5: mu = 1000
        u = uniform(N, 2, offset=offset)
        y0 = np.stack([2.0 + (u[:, 0] - 0.5), u[:, 1] - 0.5], axis=1)
        return "vdp_mu", y0, np.full((N, 1), 1000.0), 0.0, 3000.0
    if name == "linear100":      # benches/benchmark.py:137-146 (y' = -y, N = 100), y0 ~ U(0.5, 1.5)
        u = uniform(N, 100, offset=offset)
        return "linear100", 0.5 + u, None, 0.0, 10.0
    if name == "medakzo":        # tests/test_ivp.py:244-269 on 32 grid points: u = 0, v = v0 ~ U(0.9, 1.1)
        u = uniform(N, 1, offset=offset)
        y0 = np.zeros((N, 64))
        y0[:, 1::2] = 0.9 + 0.2 * u
        return "medakzo64", y0, None, 0.0, 20.0
    raise KeyError(name)


def medakzo_sparsity(n_grid: int) -> np.ndarray:
    """Structure of the MEDAKZO Jacobian on `n_grid` points (2 n_grid states) as a dense 0/1 array -- the pattern of the
    reference's tests/test_helpers.py:81-112 (`medazko_sparsity`), for Options.jac_sparsity."""
    n = 2 * n_grid
    S = np.zeros((n, n), dtype=np.int8)
    i = np.arange(n_grid) * 2            # u rows: f[i] reads y[i - 2], y[i], y[i + 1], y[i + 2]
    S[i[1:], i[1:] - 2] = 1
    S[i, i] = 1
    S[i, i + 1] = 1
    S[i[:-1], i[:-1] + 2] = 1
    j = i + 1                            # v rows: f[j] reads y[j - 1], y[j]
    S[j, j] = 1
    S[j, j - 1] = 1
    return S


def medakzo_cuda_source(n_grid: int) -> str:
    """MEDAKZO (reference tests/test_helpers.py:54-80, `fun_medazko`) on `n_grid` points as a user problem in CUDA C:
    the per-component RHS `ivp_ode_i` the warp-per-trajectory kernels (n > 32) ask for.  n = 2 n_grid states."""
    return f"""
#define NG {int(n_grid)}
__device__ __forceinline__ double medakzo_z(double t, const double* y, int m) {{   // hstack((phi, 0, y, y[-2]))
  if (m == 0) return t <= 5.0 ? 2.0 : 0.0;
  if (m == 1) return 0.0;
  if (m == 2 * NG + 2) return y[2 * NG - 2];
  return y[m - 2];
}}
__device__ double ivp_ode_i(double t, const double* y, const double* p, int i) {{
  const double k = 100.0, c = 4.0, d = 1.0 / (double)NG;
  const int j = i / 2 + 1;
  const double u = medakzo_z(t, y, 2 * j), v = medakzo_z(t, y, 2 * j + 1);
  if (i & 1) return -k * v * u;
  const double w = (double)j * d - 1.0;
  const double alpha = 2.0 * ((w * w) * w) / (c * c), beta = ((w * w) * (w * w)) / (c * c);
  const double zp = medakzo_z(t, y, 2 * j + 2), zm = medakzo_z(t, y, 2 * j - 2);
  return alpha * (zp - zm) / (2.0 * d) + beta * (zm - 2.0 * u + zp) / (d * d) - k * u * v;
}}
"""
