// ivpb_problems.cuh -- built-in problems as __device__ code compiled together with the solvers.
// A problem is the device form of the reference's `IVP` trait (src/ivp.rs:27-121): ode / events /
// jac, with the struct fields of the reference programs read from the trajectory's parameter row `p`.
#pragma once
#include "ivpb_problems_min.cuh"
#include "ivpb_fastmath.cuh"
#include "ivpb_exact.cuh"

namespace ivpb {

struct PDecay : ProblemDefaults<1, 1, 0> {      // reference examples/exponential_decay.rs:10-12
  IVPB_DEV void ode(double, const double* y, const double* p, double* d) { d[0] = -p[0] * y[0]; }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double*, const double* p, double* J) { J[0] = -p[0]; }
};

struct PVdpEps : ProblemDefaults<2, 1, 0> {     // reference examples/van_der_pol.rs:10-13
  IVPB_DEV void ode(double, const double* y, const double* p, double* d) {
    d[0] = y[1];
    d[1] = ((1.0 - y[0] * y[0]) * y[1] - y[0]) / p[0];
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* y, const double* p, double* J) {
    J[0] = 0.0; J[1] = 1.0;
    J[2] = (-2.0 * y[0] * y[1] - 1.0) / p[0];
    J[3] = (1.0 - y[0] * y[0]) / p[0];
  }
};

struct PVdpMu : ProblemDefaults<2, 1, 0> {      // reference benches/benchmark.py:22-27
  IVPB_DEV void ode(double, const double* y, const double* p, double* d) {
    d[0] = y[1];
    d[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* y, const double* p, double* J) {
    J[0] = 0.0; J[1] = 1.0;
    J[2] = -2.0 * p[0] * y[0] * y[1] - 1.0;
    J[3] = p[0] * (1.0 - y[0] * y[0]);
  }
};

struct PLorenz : ProblemDefaults<3, 3, 0> {     // reference benches/benchmark.py:30-37
  IVPB_DEV void ode(double, const double* y, const double* p, double* d) {
    d[0] = p[0] * (y[1] - y[0]);
    d[1] = y[0] * (p[1] - y[2]) - y[1];
    d[2] = y[0] * y[1] - p[2] * y[2];
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* y, const double* p, double* J) {
    J[0] = -p[0];        J[1] = p[0];   J[2] = 0.0;
    J[3] = p[1] - y[2];  J[4] = -1.0;   J[5] = -y[0];
    J[6] = y[1];         J[7] = y[0];   J[8] = -p[2];
  }
};

// Robertson as an index-1 DAE, M = diag(1, 1, 0) (the third row is the conservation law x + y + z = 1): the classic
// mass-matrix demonstration for RADAU5.  Needs Options.mass_storage = Full.
struct PRobertsonDae : ProblemDefaults<3, 3, 0> {
  IVPB_DEV void ode(double, const double* s, const double* p, double* d) {
    const double x = s[0], y = s[1], z = s[2];
    d[0] = -p[0] * x + p[1] * y * z;
    d[1] = p[0] * x - p[1] * y * z - p[2] * y * y;
    d[2] = x + y + z - 1.0;
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* s, const double* p, double* J) {
    const double y = s[1], z = s[2];
    J[0] = -p[0];  J[1] = p[1] * z;                       J[2] = p[1] * y;
    J[3] = p[0];   J[4] = -p[1] * z - 2.0 * p[2] * y;     J[5] = -p[1] * y;
    J[6] = 1.0;    J[7] = 1.0;                            J[8] = 1.0;
  }
  static constexpr bool HAS_MASS = true;
  IVPB_DEV void mass(const double*, double* M) {
#pragma unroll
    for (int k = 0; k < 9; ++k) M[k] = 0.0;
    M[0] = 1.0; M[4] = 1.0;
  }
};

// M y' = p0 A y with a full, invertible, non-symmetric constant M (every entry of the mass loops is exercised)
struct PMassLinear3 : ProblemDefaults<3, 1, 0> {
  IVPB_DEV void ode(double, const double* y, const double* p, double* d) {
    d[0] = p[0] * (-2.0 * y[0] + 1.0 * y[1]);
    d[1] = p[0] * (1.0 * y[0] - 2.0 * y[1] + 1.0 * y[2]);
    d[2] = p[0] * (1.0 * y[1] - 2.0 * y[2]);
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double*, const double* p, double* J) {
    J[0] = -2.0 * p[0]; J[1] = p[0];        J[2] = 0.0;
    J[3] = p[0];        J[4] = -2.0 * p[0]; J[5] = p[0];
    J[6] = 0.0;         J[7] = p[0];        J[8] = -2.0 * p[0];
  }
  static constexpr bool HAS_MASS = true;
  IVPB_DEV void mass(const double*, double* M) {
    M[0] = 2.0;  M[1] = 0.5;  M[2] = 0.0;
    M[3] = 0.25; M[4] = 1.5;  M[5] = -0.5;
    M[6] = 0.0;  M[7] = 0.75; M[8] = 3.0;
  }
};

// The bouncing ball without host round trips (oracle/problems.hpp BallBounce; reference examples/bouncing_ball.py:14-36
// restarts solve_ivp after every terminal event, src/solout.rs:18-29 describes the in-solver alternative used here).
struct PBallBounce : ProblemDefaults<2, 3, 0> {
  IVPB_DEV void ode(double, const double* s, const double* p, double* d) {
    const double vy = s[1];
    d[0] = vy;
    d[1] = -p[0] - p[1] * vy * fabs(vy);
  }
  static constexpr bool HAS_SOLOUT = true;
  template <class Interp, class Emit>
  IVPB_DEV int solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit) {
    if (!dense.valid()) { emit(x, y); return 0; }
    if (!(y[0] < 0.0)) return 0;
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
    if (fabs(y[1]) < 0.1) return 1;
    return 2;
  }
};

struct PCr3bp : ProblemDefaults<6, 1, 0> {      // reference examples/cr3bp.rs:24-35
  IVPB_DEV void ode(double, const double* s, const double* p, double* d) {
    const double mu = p[0];
    const double x = s[0], y = s[1], z = s[2];
    d[0] = s[3]; d[1] = s[4]; d[2] = s[5];
#ifdef IVPB_STRICT
    // the reference's 2 sqrt + 6 divisions, correctly rounded (ivpb_exact.cuh): the six quotients have two distinct
    // denominators, so two refined reciprocals serve all of them.  All eight run branch-free with ONE test of their
    // guards at the end; the guarded forms repeat the group if an operand left the fast-path range (never on an orbit).
    bool ok = true;
    {
      const double r1 = ex::sqrt_fast((x + mu) * (x + mu) + y * y + z * z, ok);
      const double r2 = ex::sqrt_fast((x - 1.0 + mu) * (x - 1.0 + mu) + y * y + z * z, ok);
      const ex::Recip r13 = ex::recip((r1 * r1) * r1), r23 = ex::recip((r2 * r2) * r2);   // powi(3)
      d[3] = x + 2.0 * s[4] - ex::div_fast((1.0 - mu) * (x + mu), r13, ok) - ex::div_fast(mu * (x - 1.0 + mu), r23, ok);
      d[4] = y - 2.0 * s[3] - ex::div_fast((1.0 - mu) * y, r13, ok) - ex::div_fast(mu * y, r23, ok);
      d[5] = ex::div_fast(-(1.0 - mu) * z, r13, ok) - ex::div_fast(mu * z, r23, ok);
    }
    if (!ok) {
      const double r1 = ex::sqrt((x + mu) * (x + mu) + y * y + z * z);
      const double r2 = ex::sqrt((x - 1.0 + mu) * (x - 1.0 + mu) + y * y + z * z);
      const ex::Recip r13 = ex::recip((r1 * r1) * r1), r23 = ex::recip((r2 * r2) * r2);
      d[3] = x + 2.0 * s[4] - ex::div((1.0 - mu) * (x + mu), r13) - ex::div(mu * (x - 1.0 + mu), r23);
      d[4] = y - 2.0 * s[3] - ex::div((1.0 - mu) * y, r13) - ex::div(mu * y, r23);
      d[5] = ex::div(-(1.0 - mu) * z, r13) - ex::div(mu * z, r23);
    }
#else
    // fast build: 1/r^3 = rsqrt(r^2)^3 -- two slow-path-free reciprocal square roots (ivpb_fastmath.cuh)
    // replace the reference's 2 sqrt + 6 divisions (a few ulp apart, like every other fast-mode operation)
    const double xm = x + mu, xn = x - 1.0 + mu;
    const double yz = fma(y, y, z * z);
    const double i1 = fm::rsqrt(fma(xm, xm, yz)), i2 = fm::rsqrt(fma(xn, xn, yz));
    const double g1 = (1.0 - mu) * ((i1 * i1) * i1), g2 = mu * ((i2 * i2) * i2);
    d[3] = fma(-g2, xn, fma(-g1, xm, fma(2.0, s[4], x)));
    d[4] = fma(-(g1 + g2), y, fma(-2.0, s[3], y));
    d[5] = -(g1 + g2) * z;
#endif
  }
};

struct PBall : ProblemDefaults<2, 2, 1> {       // reference examples/bouncing_ball.rs:11-31
  IVPB_DEV void ode(double, const double* s, const double* p, double* d) {
    const double vy = s[1];
    d[0] = vy;
    d[1] = -p[0] - p[1] * vy * fabs(vy);
  }
  IVPB_DEV void events(double, const double* s, const double*, double* g) { g[0] = s[0]; }
  IVPB_HD int default_dir(int) { return -1; }       // config.negative()
  IVPB_HD i64 default_term(int) { return 1; }       // config.terminal()
};

struct PRobertson : ProblemDefaults<3, 3, 0> {  // reference tests/test_stiff.py:104-110
  IVPB_DEV void ode(double, const double* s, const double* p, double* d) {
    const double x = s[0], y = s[1], z = s[2];
    d[0] = -p[0] * x + p[1] * y * z;
    d[1] = p[0] * x - p[1] * y * z - p[2] * y * y;
    d[2] = p[2] * y * y;
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* s, const double* p, double* J) {
    const double y = s[1], z = s[2];
    J[0] = -p[0]; J[1] = p[1] * z;                     J[2] = p[1] * y;
    J[3] = p[0];  J[4] = -p[1] * z - 2.0 * p[2] * y;   J[5] = -p[1] * y;
    J[6] = 0.0;   J[7] = 2.0 * p[2] * y;               J[8] = 0.0;
  }
};

struct PSho : ProblemDefaults<2, 0, 1> {        // reference tests/common.rs:3-9, tests/ivp.rs:151-220
  IVPB_DEV void ode(double, const double* y, const double*, double* d) { d[0] = y[1]; d[1] = -y[0]; }
  IVPB_DEV void events(double, const double* y, const double*, double* g) { g[0] = y[0]; }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double*, const double*, double* J) { J[0] = 0.0; J[1] = 1.0; J[2] = -1.0; J[3] = 0.0; }
};

struct PZero3 : ProblemDefaults<3, 0, 0> {      // reference tests/ivp.rs:11-18
  IVPB_DEV void ode(double, const double*, const double*, double* d) { d[0] = 0.0; d[1] = 0.0; d[2] = 0.0; }
};

struct PExp2 : ProblemDefaults<2, 0, 0> {       // reference tests/ivp.rs:291-297
  IVPB_DEV void ode(double, const double* y, const double*, double* d) { d[0] = y[0]; d[1] = y[1]; }
};

struct PRational : ProblemDefaults<2, 0, 0> {   // reference tests/test_helpers.py:23-25, 34-40
  IVPB_DEV void ode(double t, const double* y, const double*, double* d) {
    d[0] = y[1] / t;
    d[1] = y[1] * (y[0] + 2.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double t, const double* y, const double*, double* J) {
    J[0] = 0.0; J[1] = 1.0 / t;
    J[2] = -2.0 * y[1] * y[1] / (t * (y[0] - 1.0) * (y[0] - 1.0));
    J[3] = (y[0] + 4.0 * y[1] - 1.0) / (t * (y[0] - 1.0));
  }
};

struct PCannon : ProblemDefaults<2, 0, 1> {     // reference tests/test_ivp.py:153-160
  IVPB_DEV void ode(double, const double* y, const double*, double* d) { d[0] = y[1]; d[1] = -9.80665; }
  IVPB_DEV void events(double, const double* y, const double*, double* g) { g[0] = y[0]; }
  IVPB_HD int default_dir(int) { return -1; }
  IVPB_HD i64 default_term(int) { return 1; }
};

// ---- large systems: one trajectory per warp (WarpLayout), the RHS is given per component ----
struct PLinear100 : ProblemDefaults<100, 0, 0> {  // reference benches/benchmark.py:39-41,137-146 (dy/dt = -y, N = 100)
  static constexpr bool HAS_ODE_I = true;
  IVPB_DEV double ode_i(double, const double* y, const double*, int i) { return -y[i]; }
};

// MEDAKZO (reference tests/test_ivp.py:77-101 `fun_medazko`) on 32 grid points: n = 64, stiff reaction-diffusion.
struct PMedakzo64 : ProblemDefaults<64, 0, 0> {
  static constexpr bool HAS_ODE_I = true;
  IVPB_DEV double z(double t, const double* y, int m) {      // y padded as hstack((phi, 0, y, y[-2]))
    if (m == 0) return t <= 5.0 ? 2.0 : 0.0;
    if (m == 1) return 0.0;
    if (m == 2 * 32 + 2) return y[2 * 32 - 2];
    return y[m - 2];
  }
  IVPB_DEV double ode_i(double t, const double* y, const double*, int i) {
    constexpr int NG = 32;
    const double k = 100.0, c = 4.0, d = 1.0 / (double)NG;
    const int j = i / 2 + 1;
    const double u = z(t, y, 2 * j), v = z(t, y, 2 * j + 1);
    if (i & 1) return -k * v * u;
    const double w = (double)j * d - 1.0;
    const double alpha = 2.0 * ((w * w) * w) / (c * c), beta = ((w * w) * (w * w)) / (c * c);
    const double zp = z(t, y, 2 * j + 2), zm = z(t, y, 2 * j - 2);
    return alpha * (zp - zm) / (2.0 * d) + beta * (zm - 2.0 * u + zp) / (d * d) - k * u * v;
  }
  // analytic Jacobian for jac_mode = 1 (row-major 64 x 64); oracle/problems.hpp Medakzo64::jac holds the same expressions
  static constexpr bool HAS_JAC = true;
  IVPB_DEV void jac(double, const double* y, const double*, double* J) {
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

}  // namespace ivpb
