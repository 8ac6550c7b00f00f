// test_problems.hpp -- oracle-only problems (ids >= 100) restating the callables of the reference's Python test
// suite, so that tests/test_reference_pytests.py can run each of those tests against the oracle on the CPU and
// against the CUDA path (the same right-hand sides as NVRTC user problems) on the GPU.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "problems.hpp"

namespace oracle {

// fun_rational (tests/test_helpers.py:23-25) + the event functions of tests/test_events.py:13-17,103-104
// (event_rational_1, event_rational_2, event_rational_3 with its threshold in p[0]: 7.4 there, 7 in
// tests/test_t_eval.py:139-140)
struct RationalEv : ProblemBase<RationalEv, 2, 1, 3> {
  void ode(double t, const double* y, double* d) const {
    d[0] = y[1] / t;
    d[1] = y[1] * (y[0] + 2.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
  void events(double t, const double* y, double* g) const {
    g[0] = y[0] - std::pow(y[1], 0.7);
    g[1] = std::pow(y[1], 0.6) - y[0];
    g[2] = t - p[0];
  }
  static constexpr bool HAS_JAC = true;                // tests/test_helpers.py:34-40
  void jac(double t, const double* y, double* J) const {
    J[0] = 0.0; J[1] = 1.0 / t;
    J[2] = -2.0 * y[1] * y[1] / (t * (y[0] - 1.0) * (y[0] - 1.0));
    J[3] = (y[0] + 4.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
};

// sys3 + sys3_jac + its three events (tests/test_args.py:11-35); p = (omega, k, zfinal)
struct Sys3 : ProblemBase<Sys3, 3, 3, 3> {
  void ode(double, const double* w, double* d) const {
    d[0] = -p[0] * w[1];
    d[1] = p[0] * w[0];
    d[2] = p[1] * w[2] * (1.0 - w[2]);
  }
  void events(double, const double* w, double* g) const { g[0] = w[0]; g[1] = w[1]; g[2] = w[2] - p[2]; }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* w, double* J) const {
    J[0] = 0.0;  J[1] = -p[0]; J[2] = 0.0;
    J[3] = p[0]; J[4] = 0.0;   J[5] = 0.0;
    J[6] = 0.0;  J[7] = 0.0;   J[8] = p[1] * (1.0 - 2.0 * w[2]);
  }
};

// y' = a y for each component (tests/test_args.py:73-78 fun_with_arg; test_edge_cases.py:14,53 with a = -1, 2;
// test_step_control.py:144-145 with a = -0.001); one and two components
struct Scale1 : ProblemBase<Scale1, 1, 1, 0> {
  void ode(double, const double* y, double* d) const { d[0] = p[0] * y[0]; }
};
struct Scale2 : ProblemBase<Scale2, 2, 1, 0> {
  void ode(double, const double* y, double* d) const { d[0] = p[0] * y[0]; d[1] = p[0] * y[1]; }
};

// tests/test_edge_cases.py:76-88 (gh-8848): radial Schroedinger-like system in r = exp(t)
struct Radial : ProblemBase<Radial, 2, 0, 0> {
  void ode(double t, const double* s, double* d) const {
    const double r = std::exp(t);
    const double V = -11.0 / r + 10.0 * r / (0.05 + r * r);
    d[0] = r * s[1];
    d[1] = -2.0 * r * ((-0.2 - V) * s[0] + 1.0 / r * s[1]);
  }
};

// tests/test_edge_cases.py:104-110 (gh-9198): constant rates
struct ConstRates : ProblemBase<ConstRates, 4, 0, 0> {
  void ode(double, const double*, double* d) const {
    d[0] = 1.73307544e-02; d[1] = 6.49376470e-06; d[2] = 0.0; d[3] = 0.0;
  }
};

// fun_linear + jac_linear (tests/test_helpers.py:11-16), tests/test_ivp.py:272-317, tests/test_stiff.py:14-94
struct Linear2 : ProblemBase<Linear2, 2, 0, 0> {
  void ode(double, const double* y, double* d) const { d[0] = -y[0] - 5.0 * y[1]; d[1] = y[0] + y[1]; }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double*, double* J) const { J[0] = -1.0; J[1] = -5.0; J[2] = 1.0; J[3] = 1.0; }
};

// M y' = p0 A y, n = 4, tridiagonal A and a full constant M (the shared-memory matrix variant of the device kernels)
struct MassLinear4 : ProblemBase<MassLinear4, 4, 1, 0> {
  void ode(double, const double* y, double* d) const {
    d[0] = p[0] * (-2.0 * y[0] + y[1]);
    d[1] = p[0] * (y[0] - 2.0 * y[1] + y[2]);
    d[2] = p[0] * (y[1] - 2.0 * y[2] + y[3]);
    d[3] = p[0] * (y[2] - 2.0 * y[3]);
  }
  static constexpr bool HAS_MASS = true;
  void mass(double* M) const {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) M[i * 4 + j] = (i == j) ? 2.0 + 0.5 * (double)i : 0.25 / (1.0 + (double)(i > j ? i - j : j - i)) * ((i + j) % 2 ? -1.0 : 1.0);
  }
};

// M y' = p0 A y, n = 12 (the warp-cooperative device kernels, n > 8): tridiagonal A, full constant M
struct MassLinear12 : ProblemBase<MassLinear12, 12, 1, 0> {
  void ode(double, const double* y, double* d) const {
    for (int i = 0; i < 12; ++i) {
      const double lo = i > 0 ? y[i - 1] : 0.0, hi = i < 11 ? y[i + 1] : 0.0;
      d[i] = p[0] * (lo - 2.0 * y[i] + hi);
    }
  }
  static constexpr bool HAS_MASS = true;
  void mass(double* M) const {
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) M[i * 12 + j] = (i == j) ? 2.0 + 0.25 * (double)i : 0.25 / (1.0 + (double)(i > j ? i - j : j - i)) * ((i + j) % 2 ? -1.0 : 1.0);
  }
};

// fun_medazko at the reference's own size (tests/test_helpers.py:54-80 with n = 200 grid points => 400 states), for the golden
// numbers of tests/test_ivp.py:245-269 and tests/test_stiff.py:147-183.  (The device kernels stop at n = 84 / 118; this pins
// the ORACLE's RADAU / BDF / LU / finite-difference Jacobian against the reference's published MEDAKZO values.)
struct Medakzo400 : ProblemBase<Medakzo400, 400, 0, 0> {
  static constexpr int NG = 200;
  static double z(double t, const double* y, int m) {
    if (m == 0) return t <= 5.0 ? 2.0 : 0.0;
    if (m == 1) return 0.0;
    if (m == 2 * NG + 2) return y[2 * NG - 2];
    return y[m - 2];
  }
  void ode(double t, const double* y, double* f) const {
    const double k = 100.0, c = 4.0, d = 1.0 / (double)NG;
    for (int i = 0; i < 2 * NG; ++i) {
      const int j = i / 2 + 1;
      const double u = z(t, y, 2 * j), v = z(t, y, 2 * j + 1);
      if (i & 1) { f[i] = -k * v * u; continue; }
      const double w = (double)j * d - 1.0;
      const double alpha = 2.0 * ((w * w) * w) / (c * c), beta = ((w * w) * (w * w)) / (c * c);
      const double zp = z(t, y, 2 * j + 2), zm = z(t, y, 2 * j - 2);
      f[i] = alpha * (zp - zm) / (2.0 * d) + beta * (zm - 2.0 * u + zp) / (d * d) - k * u * v;
    }
  }
};

// Test infrastructure for SolOut hooks in the warp-per-trajectory kernels (n > 32 explicit, n > 8 RADAU / BDF): 40 decaying
// components y_i' = -(0.5 + 0.05 i) y_i; the hook finds the point where y[0] falls below 0.5 on the step interpolant
// (bisection), restarts there with the whole state doubled (ModifiedSolution), records the point, and stops the
// integration at the third kick (Interrupt).  tests/test_solout_hook.py holds the same hook in CUDA C.
struct DecayKick40 : ProblemBase<DecayKick40, 40, 0, 0> {
  void ode(double, const double* y, double* d) const {
    for (int i = 0; i < 40; ++i) d[i] = -(0.5 + 0.05 * (double)i) * y[i];
  }
  static constexpr bool HAS_SOLOUT = true;
  template <class Interp, class Emit>
  int solout(double xold, double& x, double* y, double* state, const Interp& dense, Emit& emit) const {
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
};

}  // namespace oracle
