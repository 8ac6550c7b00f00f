// problems.hpp -- the built-in problems as CPU functors for the oracle.  TEST INFRASTRUCTURE ONLY.
// Each functor is the `IVP` trait impl (reference src/ivp.rs:27-121) of the cited reference program,
// with the struct fields read from the trajectory's parameter row `p`.
#pragma once
#include "ivp_oracle.hpp"

namespace oracle {

template <class D, int N_, int P_, int NEV_>
struct ProblemBase {
  static constexpr int N = N_, P = P_, NEV = NEV_;
  const double* p = nullptr;               // parameter row of this trajectory
  const EventConfig* ev_cfg = nullptr;     // override of event_config (from ivpb_options), or null
  int jac_mode = 0;                        // 0: default finite differences (ivp.rs:67-107); 1: analytic override
  const Sparsity* sparsity = nullptr;      // jac_sparsity of the Python front end (src/python/sparsity.rs), or null
  int n_events() const { return NEV; }
  void events(double, const double*, double*) const {}
  EventConfig default_event_config(int) const { return EventConfig(); }
  EventConfig event_config(int i) const {
    return ev_cfg ? ev_cfg[i] : static_cast<const D*>(this)->default_event_config(i);
  }
  static constexpr bool HAS_JAC = false;
  void jac(double, const double*, double*) const {}
  static constexpr bool HAS_SOLOUT = false;   // the problem's own SolOut (src/solout.rs:55-63), see UserSolOut
  static constexpr int NSTATE = 4;
  static constexpr bool HAS_MASS = false;  // IVP::mass (src/ivp.rs:109-120): constant, row-major n x n
  void mass(double*) const {}
};

struct Decay : ProblemBase<Decay, 1, 1, 0> {          // examples/exponential_decay.rs:10-12
  void ode(double, const double* y, double* d) const { d[0] = -p[0] * y[0]; }
};
struct VdpEps : ProblemBase<VdpEps, 2, 1, 0> {        // examples/van_der_pol.rs:10-13
  void ode(double, const double* y, double* d) const {
    d[0] = y[1];
    d[1] = ((1.0 - y[0] * y[0]) * y[1] - y[0]) / p[0];
  }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* y, double* J) const {
    J[0] = 0.0; J[1] = 1.0;
    J[2] = (-2.0 * y[0] * y[1] - 1.0) / p[0];
    J[3] = (1.0 - y[0] * y[0]) / p[0];
  }
};
struct VdpMu : ProblemBase<VdpMu, 2, 1, 0> {          // benches/benchmark.py:22-27
  void ode(double, const double* y, double* d) const {
    d[0] = y[1];
    d[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
  }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* y, double* J) const {
    J[0] = 0.0; J[1] = 1.0;
    J[2] = -2.0 * p[0] * y[0] * y[1] - 1.0;
    J[3] = p[0] * (1.0 - y[0] * y[0]);
  }
};
struct Lorenz : ProblemBase<Lorenz, 3, 3, 0> {        // benches/benchmark.py:30-37
  void ode(double, const double* y, double* d) const {
    d[0] = p[0] * (y[1] - y[0]);
    d[1] = y[0] * (p[1] - y[2]) - y[1];
    d[2] = y[0] * y[1] - p[2] * y[2];
  }
};
struct Cr3bp : ProblemBase<Cr3bp, 6, 1, 0> {          // examples/cr3bp.rs:24-35
  void ode(double, const double* s, double* d) const {
    const double mu = p[0];
    const double x = s[0], y = s[1], z = s[2], vx = s[3], vy = s[4], vz = s[5];
    const double r1 = std::sqrt(sq(x + mu) + sq(y) + sq(z));
    const double r2 = std::sqrt(sq(x - 1.0 + mu) + sq(y) + sq(z));
    auto cube = [](double r) { return (r * r) * r; };   // powi(3)
    d[0] = vx; d[1] = vy; d[2] = vz;
    d[3] = x + 2.0 * vy - (1.0 - mu) * (x + mu) / cube(r1) - mu * (x - 1.0 + mu) / cube(r2);
    d[4] = y - 2.0 * vx - (1.0 - mu) * y / cube(r1) - mu * y / cube(r2);
    d[5] = -(1.0 - mu) * z / cube(r1) - mu * z / cube(r2);
  }
};
struct Ball : ProblemBase<Ball, 2, 2, 1> {            // examples/bouncing_ball.rs:11-31
  void ode(double, const double* s, double* d) const {
    const double vy = s[1];
    d[0] = vy;
    d[1] = -p[0] - p[1] * vy * std::fabs(vy);
  }
  void events(double, const double* s, double* g) const { g[0] = s[0]; }
  EventConfig default_event_config(int) const { EventConfig c; c.terminal_count = 1; c.direction = Direction::Negative; return c; }
};
// The bouncing ball WITHOUT host round trips (reference examples/bouncing_ball.py:14-36 restarts solve_ivp after every
// terminal event; src/solout.rs:18-29 offers the in-solver alternative): a SolOut that finds the impact inside the step
// with the interpolant, moves (x, y) to the impact with the velocity reversed and damped, and returns ModifiedSolution.
// p = (g, drag, restitution).  state[0] = bounces so far, state[1] = time of the last one.  Every impact is emitted.
struct BallBounce : ProblemBase<BallBounce, 2, 3, 0> {
  void ode(double, const double* s, double* d) const {
    const double vy = s[1];
    d[0] = vy;
    d[1] = -p[0] - p[1] * vy * std::fabs(vy);
  }
  static constexpr bool HAS_SOLOUT = true;
  template <class Interp, class Emit>
  int solout(double xold, double& x, double* y, double* state, const Interp& dense, Emit& emit) const {
    if (!dense.valid()) { emit(x, y); return 0; }             // initial call: record the start
    if (!(y[0] < 0.0)) return 0;                               // still above the ground at the end of the step
    // impact inside (xold, x]: bisection on the step interpolant, 60 halvings (deterministic, direction-free)
    double lo = xold, hi = x, yi[2];
    for (int it = 0; it < 60; ++it) {
      const double mid = 0.5 * (lo + hi);
      dense.eval(mid, yi);
      if (yi[0] < 0.0) hi = mid; else lo = mid;
    }
    dense.eval(hi, yi);
    x = hi;
    y[0] = 0.0;
    y[1] = -p[2] * yi[1];
    state[0] += 1.0; state[1] = x;
    emit(x, y);
    if (std::fabs(y[1]) < 0.1) return 1;                       // bouncing_ball.py:33-34: too slow to matter -> stop
    return 2;
  }
};
struct Robertson : ProblemBase<Robertson, 3, 3, 0> {  // tests/test_stiff.py:104-110
  void ode(double, const double* s, double* d) const {
    const double x = s[0], y = s[1], z = s[2];
    d[0] = -p[0] * x + p[1] * y * z;
    d[1] = p[0] * x - p[1] * y * z - p[2] * y * y;
    d[2] = p[2] * y * y;
  }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* s, double* J) const {
    const double y = s[1], z = s[2];
    J[0] = -p[0];  J[1] = p[1] * z;                       J[2] = p[1] * y;
    J[3] = p[0];   J[4] = -p[1] * z - 2.0 * p[2] * y;     J[5] = -p[1] * y;
    J[6] = 0.0;    J[7] = 2.0 * p[2] * y;                 J[8] = 0.0;
  }
};
struct Sho : ProblemBase<Sho, 2, 0, 1> {              // tests/common.rs:3-9; events tests/ivp.rs:151-220
  void ode(double, const double* y, double* d) const { d[0] = y[1]; d[1] = -y[0]; }
  void events(double, const double* y, double* g) const { g[0] = y[0]; }
};
struct Zero3 : ProblemBase<Zero3, 3, 0, 0> {          // tests/ivp.rs:11-18
  void ode(double, const double*, double* d) const { d[0] = 0.0; d[1] = 0.0; d[2] = 0.0; }
};
struct Exp2 : ProblemBase<Exp2, 2, 0, 0> {            // tests/ivp.rs:291-297
  void ode(double, const double* y, double* d) const { d[0] = y[0]; d[1] = y[1]; }
};
struct Rational : ProblemBase<Rational, 2, 0, 0> {    // tests/test_helpers.py:23-25
  void ode(double t, const double* y, double* d) const {
    d[0] = y[1] / t;
    d[1] = y[1] * (y[0] + 2.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
  static constexpr bool HAS_JAC = true;                // tests/test_helpers.py:34-40
  void jac(double t, const double* y, double* J) const {
    J[0] = 0.0; J[1] = 1.0 / t;
    J[2] = -2.0 * y[1] * y[1] / (t * (y[0] - 1.0) * (y[0] - 1.0));
    J[3] = (y[0] + 4.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
};
struct Cannon : ProblemBase<Cannon, 2, 0, 1> {        // tests/test_ivp.py:153-160
  void ode(double, const double* y, double* d) const { d[0] = y[1]; d[1] = -9.80665; }
  void events(double, const double* y, double* g) const { g[0] = y[0]; }
  EventConfig default_event_config(int) const { EventConfig c; c.terminal_count = 1; c.direction = Direction::Negative; return c; }
};

// Robertson as an index-1 DAE (Hairer & Wanner II, the classic RADAU5 mass-matrix demonstration): the third
// equation becomes the conservation law.  Exercises radau.rs:375-386 (E1/E2 with M), :525-539 (M F in the Newton
// right-hand side) and :626-634 (M in the error estimate).
struct RobertsonDae : ProblemBase<RobertsonDae, 3, 3, 0> {
  void ode(double, const double* s, double* d) const {
    const double x = s[0], y = s[1], z = s[2];
    d[0] = -p[0] * x + p[1] * y * z;
    d[1] = p[0] * x - p[1] * y * z - p[2] * y * y;
    d[2] = x + y + z - 1.0;
  }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* s, double* J) const {
    const double y = s[1], z = s[2];
    J[0] = -p[0];  J[1] = p[1] * z;                       J[2] = p[1] * y;
    J[3] = p[0];   J[4] = -p[1] * z - 2.0 * p[2] * y;     J[5] = -p[1] * y;
    J[6] = 1.0;    J[7] = 1.0;                            J[8] = 1.0;
  }
  static constexpr bool HAS_MASS = true;
  void mass(double* M) const { for (int k = 0; k < 9; ++k) M[k] = 0.0; M[0] = 1.0; M[4] = 1.0; }
};
// M y' = p0 * A y with a full, invertible, non-symmetric constant M: every entry of the mass loops is exercised.
struct MassLinear3 : ProblemBase<MassLinear3, 3, 1, 0> {
  void ode(double, const double* y, double* d) const {
    d[0] = p[0] * (-2.0 * y[0] + 1.0 * y[1]);
    d[1] = p[0] * (1.0 * y[0] - 2.0 * y[1] + 1.0 * y[2]);
    d[2] = p[0] * (1.0 * y[1] - 2.0 * y[2]);
  }
  static constexpr bool HAS_JAC = true;
  void jac(double, const double*, double* J) const {
    J[0] = -2.0 * p[0]; J[1] = p[0];        J[2] = 0.0;
    J[3] = p[0];        J[4] = -2.0 * p[0]; J[5] = p[0];
    J[6] = 0.0;         J[7] = p[0];        J[8] = -2.0 * p[0];
  }
  static constexpr bool HAS_MASS = true;
  void mass(double* M) const {
    M[0] = 2.0; M[1] = 0.5;  M[2] = 0.0;
    M[3] = 0.25; M[4] = 1.5; M[5] = -0.5;
    M[6] = 0.0; M[7] = 0.75; M[8] = 3.0;
  }
};

struct Linear100 : ProblemBase<Linear100, 100, 0, 0> {  // benches/benchmark.py:39-41,137-146
  void ode(double, const double* y, double* d) const { for (int i = 0; i < 100; ++i) d[i] = -y[i]; }
};

struct Medakzo64 : ProblemBase<Medakzo64, 64, 0, 0> {   // tests/test_ivp.py:77-101 (fun_medazko), 32 grid points
  static double z(double t, const double* y, int m) {
    if (m == 0) return t <= 5.0 ? 2.0 : 0.0;
    if (m == 1) return 0.0;
    if (m == 2 * 32 + 2) return y[2 * 32 - 2];
    return y[m - 2];
  }
  void ode(double t, const double* y, double* f) const {
    const int NG = 32;
    const double k = 100.0, c = 4.0, d = 1.0 / (double)NG;
    for (int i = 0; i < 64; ++i) {
      const int j = i / 2 + 1;
      const double u = z(t, y, 2 * j), v = z(t, y, 2 * j + 1);
      if (i & 1) { f[i] = -k * v * u; continue; }
      const double w = (double)j * d - 1.0;
      const double alpha = 2.0 * ((w * w) * w) / (c * c), beta = ((w * w) * (w * w)) / (c * c);
      const double zp = z(t, y, 2 * j + 2), zm = z(t, y, 2 * j - 2);
      f[i] = alpha * (zp - zm) / (2.0 * d) + beta * (zm - 2.0 * u + zp) / (d * d) - k * u * v;
    }
  }
  // analytic Jacobian (row-major 64 x 64; the reference's own MEDAKZO test differentiates numerically, tests/test_ivp.py:244-269):
  // test infrastructure for jac_mode = 1 on the warp-cooperative kernels -- the device problem uses the same expressions
  static constexpr bool HAS_JAC = true;
  void jac(double, const double* y, double* J) const {
    const int NG = 32;
    const double k = 100.0, c = 4.0, d = 1.0 / (double)NG;
    for (int q = 0; q < 64 * 64; ++q) J[q] = 0.0;
    for (int j = 1; j <= NG; ++j) {
      const int iu = 2 * (j - 1), iv = iu + 1;
      const double u = y[iu], v = y[iv];
      const double w = (double)j * d - 1.0;
      const double alpha = 2.0 * ((w * w) * w) / (c * c), beta = ((w * w) * (w * w)) / (c * c);
      const double a1 = alpha / (2.0 * d), b1 = beta / (d * d);
      double duu = -2.0 * b1 - k * v;                  // d f_u_j / d u_j
      if (j < NG) J[iu * 64 + iu + 2] = a1 + b1;       // d f_u_j / d u_{j+1}
      else duu = duu + (a1 + b1);                      // boundary: u_{N+1} = u_N
      if (j > 1) J[iu * 64 + iu - 2] = b1 - a1;        // d f_u_j / d u_{j-1} (j = 1: z(0) = phi(t), not a state)
      J[iu * 64 + iu] = duu;
      J[iu * 64 + iv] = -k * u;                        // d f_u_j / d v_j
      J[iv * 64 + iu] = -k * v;                        // d f_v_j / d u_j
      J[iv * 64 + iv] = -k * u;                        // d f_v_j / d v_j
    }
  }
};

}  // namespace oracle
