"""CPU checks of the arguments behind the work-removal shortcuts of the RADAU / BDF kernels (ivpb_implicit.cuh): they must
not change a single bit of what the reference computes, and these are the facts that claim rests on.  The kernels
themselves are compared with the oracle bit for bit in the GPU suite (test_stiff_ensembles_strict_bit_exact, ...)."""
import math

import numpy as np


def compute_r(order, factor):
    """reference src/methods/bdf.rs:694-713, operation for operation (numpy float64 = IEEE double)."""
    size = order + 1
    m = np.zeros((size, size))
    m[0, :] = 1.0
    for i in range(1, size):
        for j in range(1, size):
            m[i, j] = (np.float64(i) - 1.0 - np.float64(factor) * np.float64(j)) / np.float64(i)
    r = np.zeros((size, size))
    r[0, :] = m[0, :]
    for i in range(1, size):
        for j in range(size):
            r[i, j] = r[i - 1, j] * m[i, j]
    return r


def matmul_ref(a, b):
    """reference src/methods/bdf.rs:715-731: zero entries of `a` skipped, sums in ascending k."""
    rows, cols, inner = a.shape[0], b.shape[1], b.shape[0]
    res = np.zeros((rows, cols))
    for i in range(rows):
        for k in range(inner):
            c = a[i, k]
            if c == 0.0:
                continue
            for j in range(cols):
                res[i, j] = res[i, j] + c * b[k, j]
    return res


def matmul_triangular(a, b):
    """what BdfTraj::change_d evaluates: only k <= j, no zero test on `a`."""
    rows, cols = a.shape[0], b.shape[1]
    res = np.zeros((rows, cols))
    for i in range(rows):
        for j in range(cols):
            acc = np.float64(0.0)
            for k in range(j + 1):
                acc = acc + a[i, k] * b[k, j]
            res[i, j] = acc
    return res


def test_u_is_upper_triangular_and_triangular_product_is_bit_identical():
    rng = np.random.default_rng(7)
    factors = np.concatenate([[0.2, 0.5, 0.9, 1.0 / 3.0, 2.0, 10.0, 0.25, 4.0], rng.uniform(0.2, 10.0, 400)])
    for order in range(1, 6):
        u = compute_r(order, 1.0)
        for m in range(order + 1):
            for row in range(m):
                assert u[m, row] == 0.0, (order, m, row)          # (+-)0 below the diagonal
        for f in factors:
            r = compute_r(order, f)
            full, tri = matmul_ref(r, u), matmul_triangular(r, u)
            # bit for bit, the sign of a zero included
            assert np.array_equal(full.view(np.uint64), tri.view(np.uint64)), (order, f)


def test_power_by_products_decides_like_powf_outside_the_margin():
    """bdf.rs:408 (rate^remaining, remaining = 1..3) and radau.rs:572 (theta^rem, rem = 0..4): the comparison made with
    repeated products equals the one made with libm's pow whenever the product-based estimate is more than 1e-12 away
    from the bound, and the two estimates never differ by more than 1e-15."""
    rng = np.random.default_rng(11)
    n = 400000
    rate = np.concatenate([rng.uniform(0.0, 1.0, n), 10.0 ** rng.uniform(-12, 0, n)])
    rate = rate[(rate > 0.0) & (rate < 1.0)]
    dy = 10.0 ** rng.uniform(-14, 2, rate.size)
    tol = 10.0 ** rng.uniform(-10, -1, rate.size)
    worst = 0.0
    for rem in (1, 2, 3, 4):
        pa = rate.copy()
        for _ in range(rem - 1):
            pa = pa * rate
        exact = np.power(rate, float(rem)) / (1.0 - rate) * dy
        approx = pa / (1.0 - rate) * dy
        rel = np.abs(approx - exact) / exact
        worst = max(worst, float(rel.max()))
        decided_true = approx > tol * (1.0 + 1e-12)
        decided_false = approx < tol * (1.0 - 1e-12)
        assert np.all(exact[decided_true] > tol[decided_true])
        assert not np.any(exact[decided_false] > tol[decided_false])
        # adversarial: bounds placed right at the estimate -- inside the margin neither branch may claim a decision
        near = exact * (1.0 + rng.uniform(-5e-13, 5e-13, exact.size))
        undecided = ~(approx > near * (1.0 + 1e-12)) & ~(approx < near * (1.0 - 1e-12))
        assert undecided.all()
    assert worst < 1e-15, worst
    assert math.pow(0.37, 1.0) == 0.37 and math.pow(0.37, 0.0) == 1.0      # rem = 1 / 0 are exact in libm as well


def test_zero_residual_exit_matches_the_reference_flow():
    """bdf.rs:398-421 with an all-zero right-hand side: the solve returns zeros, dy_norm == 0, no rate test fires, the
    increments change nothing and the loop leaves converged -- the exit BdfTraj::step takes at once."""
    y_new, delta = np.array([1.5, -2.0, 0.0]), np.array([1e-3, 0.0, -4e-7])
    dy = np.zeros(3)                                   # LU solve of a zero vector
    scale = np.array([1e-6, 2e-6, 1e-6])
    dy_norm = math.sqrt(float(np.sum((dy / scale) ** 2)) / 3.0)
    assert dy_norm == 0.0
    rate = dy_norm / 4.6e-31
    assert not (rate >= 1.0) and not (rate ** 3 / (1.0 - rate) * dy_norm > 1e-3)
    assert np.array_equal(y_new + dy, y_new) and np.array_equal(delta + dy, delta)
