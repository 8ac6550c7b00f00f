// ivp_oracle.hpp -- CPU restatement of Ryan-D-Gast/ivp (crate `ivp` v0.5.1) for the
// batched-solve hot path.  TEST INFRASTRUCTURE ONLY: nothing under ivp_b200/ may include,
// link or call this file.  It exists so that tests/, __graft_entry__.smoke() and the
// `cpu_baseline` / `--impl reference` legs of bench.py have something to compare and time
// against (the Rust crate itself cannot be built: no cargo/rustc in this image).
//
// PARITY PINNING: the reference cannot be executed here, and its own tests are all
// tolerance based (no golden step counts / bit patterns).  The oracle is therefore pinned
// against every known-answer expectation those tests hold for this path (see
// tests/test_oracle_pins.py, each citing tests/*.rs / tests/test_ivp.py line ranges, and
// tests/test_reference_pytests.py, the reference's Python test files restated one by one), but
// step-count parity is defined oracle-vs-GPU only: "parity unpinned at the bit level".
//
// Every function cites the reference file:line it follows.  Semantics mirrored from Rust:
// no FMA contraction (compile with -ffp-contract=off), powi(2) == x*x, f64::signum(+0)=+1,
// f64::max/min ignore NaN (== fmax/fmin), usize::MAX default step cap.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace oracle {

// src/status.rs:4-19 (declaration order == integer code used on the device)
enum class Status : int {
  Success = 0, UserInterrupt = 1, NeedLargerNMax = 2, StepSizeTooSmall = 3,
  ProbablyStiff = 4, SingularMatrix = 5, PoorConvergence = 6
};
// src/solve/options.rs:14-27
enum class Method : int { RK23 = 0, DOPRI5 = 1, DOP853 = 2, RK4 = 3, RADAU = 4, BDF = 5 };
// src/solve/event.rs:62-77
enum class Direction : int { All = 0, Positive = 1, Negative = -1 };
struct EventConfig {            // src/solve/event.rs:5-27
  Direction direction = Direction::All;
  long terminal_count = -1;     // -1 == None
};
// src/solout.rs:73-78 (DefaultSolOut only ever returns Continue / Interrupt)
enum class Flag { Continue, Interrupt, ModifiedSolution };   // src/solout.rs:73-78 (XOut only gates the interpolant: == Continue here)

struct ConfigError : std::runtime_error { using std::runtime_error::runtime_error; };

// src/python/sparsity.rs:13-154 -- SparsityStructure: per-column row lists and the greedy first-fit column groups
// (two columns share a group iff they have no structural non-zero row in common).
struct Sparsity {
  size_t n = 0, n_groups = 0;
  std::vector<std::vector<size_t>> col_to_rows;
  std::vector<size_t> groups;
  // from compressed columns (include/ivpb.h: jac_sparsity_colptr / jac_sparsity_rows)
  Sparsity(size_t n_, const int* colptr, const int* rows) : n(n_), col_to_rows(n_), groups(n_, (size_t)-1) {
    for (size_t c = 0; c < n; ++c)
      for (int k = colptr[c]; k < colptr[c + 1]; ++k) col_to_rows[c].push_back((size_t)rows[k]);
    std::vector<std::vector<bool>> group_rows;           // group_columns, sparsity.rs:109-154
    for (size_t col = 0; col < n; ++col) {
      const std::vector<size_t>& rs = col_to_rows[col];
      bool assigned = false;
      for (size_t g = 0; g < group_rows.size() && !assigned; ++g) {
        bool can_use = true;
        for (size_t r : rs) if (group_rows[g][r]) { can_use = false; break; }
        if (can_use) {
          groups[col] = g;
          for (size_t r : rs) group_rows[g][r] = true;
          assigned = true;
        }
      }
      if (!assigned) {
        groups[col] = n_groups++;
        std::vector<bool> used(n, false);
        for (size_t r : rs) used[r] = true;
        group_rows.push_back(used);
      }
    }
  }
};

// src/methods/mod.rs:104-214 -- scalar or per-component tolerance
struct Tol {
  std::vector<double> v;
  bool scalar = true;
  Tol() : v{0.0} {}
  Tol(double s) : v{s} {}
  static Tol vec(std::vector<double> w) { Tol t; t.v = std::move(w); t.scalar = false; return t; }
  // Tolerance::Vector indexes out of bounds => panic in Rust (methods/mod.rs:201); .at() throws here.
  double operator[](size_t i) const { return scalar ? v[0] : v.at(i); }
};

inline double signum(double x) {              // f64::signum
  if (std::isnan(x)) return x;
  return std::signbit(x) ? -1.0 : 1.0;
}
inline double sq(double x) { return x * x; }  // powi(2)

// src/methods/mod.rs:29-97
struct IntegrationResult {
  double h = 0.0;
  Status status = Status::Success;
  size_t nfev = 0, njev = 0, nlu = 0;
  size_t nstep = 0, naccpt = 0, nrejct = 0;
};

// src/dense.rs:17,32-97 -- borrowed per-step interpolant
using InterpFn = void (*)(double xi, double* yi, size_t n, const double* cont, double xold, double h);
struct StepInterp {
  const double* cont; size_t cont_len; double xold, h; InterpFn fn;
  void interpolate(double xi, double* yi, size_t n) const { fn(xi, yi, n, cont, xold, h); }
};

// ---------------------------------------------------------------------------------------
// hinit -- src/methods/mod.rs:217-281
template <class F>
double hinit(const F& f, double x, const std::vector<double>& y, double posneg,
             const std::vector<double>& f0, std::vector<double>& f1, std::vector<double>& y1,
             int iord, double hmax, const Tol& atol, const Tol& rtol) {
  const size_t n = y.size();
  double dnf = 0.0, dny = 0.0;
  for (size_t i = 0; i < n; ++i) {
    double sk = atol[i] + rtol[i] * std::fabs(y[i]);
    dnf += (f0[i] / sk) * (f0[i] / sk);
    dny += (y[i] / sk) * (y[i] / sk);
  }
  double h;
  if (dnf <= 1e-10 || dny <= 1e-10) h = 1.0e-6;
  else h = std::sqrt(dny / dnf) * 0.01;
  if (h > std::fabs(hmax)) h = std::fabs(hmax);
  h = std::fabs(h) * signum(posneg);
  for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * f0[i];
  f.ode(x + h, y1.data(), f1.data());
  double der2 = 0.0;
  for (size_t i = 0; i < n; ++i) {
    double sk = atol[i] + rtol[i] * std::fabs(y[i]);
    double df = (f1[i] - f0[i]) / sk;
    der2 += df * df;
  }
  der2 = std::sqrt(der2) / std::fabs(h);
  double der12 = std::fmax(std::fabs(der2), std::sqrt(dnf));
  double h1;
  if (der12 <= 1.0e-15) h1 = std::fmax(1.0e-6, std::fabs(h) * 1.0e-3);
  else h1 = std::pow(0.01 / der12, 1.0 / (double)iord);
  // mod.rs:279 -- includes the extra |h| term (reference quirk)
  double hf = std::fmin(std::fmin(std::fmin(std::fabs(h), 100.0 * std::fabs(h)), h1), std::fabs(hmax));
  return std::fabs(hf) * signum(posneg);
}

// ---------------------------------------------------------------------------------------
// Shared settings handed from solve_ivp to each method (the struct fields solve_ivp sets:
// src/solve/solve_ivp.rs:184-286; everything else keeps the struct defaults).
struct StepCfg {
  bool has_max_step = false;  double max_step = 0.0;
  bool has_first_step = false; double first_step = 0.0;
  bool has_min_step = false;  double min_step = 0.0;
  size_t max_steps = std::numeric_limits<size_t>::max();
  // RADAU only (solve_ivp.rs:246-258): Options.mass_storage == Full, Options.nind1..3 (< 0 => None)
  bool mass_full = false;
  long nind1 = -1, nind2 = -1, nind3 = -1;
};

// ======================================================================================
// DOP853 -- src/methods/dop853.rs:114-670
namespace dop853 {
#include "dop853_coeffs.inc"

inline void interpolate(double xi, double* yi, size_t n, const double* c, double xold, double h) {
  // dop853.rs:659-670
  double s = (xi - xold) / h;
  double s1 = 1.0 - s;
  for (size_t i = 0; i < n; ++i) {
    double conpar = c[4 * n + i] + s * (c[5 * n + i] + s1 * (c[6 * n + i] + s * c[7 * n + i]));
    yi[i] = c[i] + s * (c[n + i] + s1 * (c[2 * n + i] + s * (c[3 * n + i] + s1 * conpar)));
  }
}

template <class F, class S>
IntegrationResult solve(const F& f, double x0, const std::vector<double>& y0, double xend,
                        const Tol& rtol, const Tol& atol, const StepCfg& cfg, S* so) {
  // struct defaults dop853.rs:33-63
  const double uround = 2.3e-16, safe = 0.9, beta = 0.0;
  const double facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0;
  const size_t nstiff = 1000;
  const bool dense = true;  // never overridden by solve_ivp (solve_ivp.rs:231-235)
  double x = x0;
  std::vector<double> y = y0;
  const double h_max = cfg.has_max_step ? std::fabs(cfg.max_step) : std::fabs(xend - x);
  const size_t nmax = cfg.max_steps;
  if (nmax == 0) throw ConfigError("max_steps must be positive");

  const size_t n = y.size();
  std::vector<double> y1(n), k1(n), k2(n), k3(n), k4(n), k5(n), k6(n), k7(n), k8(n), k9(n), k10(n);
  std::vector<double> cont(8 * n);
  int nonstiff = 0, iasti = 0;
  double facold = 1e-4, hlamb = 0.0;
  bool last = false, reject = false;
  IntegrationResult R;
  double xold = x;
  const double expo1 = 1.0 / 8.0 - beta * 0.2;
  const double posneg = signum(xend - x);

  f.ode(x, y.data(), k1.data());
  R.nfev += 1;
  double h;
  if (cfg.has_first_step) h = std::fabs(cfg.first_step) * posneg;
  else { R.nfev += 1; h = hinit(f, x, y, posneg, k1, k2, y1, 8, h_max, atol, rtol); }

  if (so) {
    const Flag fl0 = so->solout(xold, x, y, nullptr);
    if (fl0 == Flag::Interrupt) {
      R.h = h; R.status = Status::UserInterrupt; return R;
    }
    if (fl0 == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // e.g. dop853.rs:258-262
  }

  for (;;) {
    if (R.nstep > nmax) { R.status = Status::NeedLargerNMax; break; }
    if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { R.status = Status::StepSizeTooSmall; break; }
    if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
    R.nstep += 1;

    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * a21 * k1[i];
    f.ode(x + c2 * h, y1.data(), k2.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
    f.ode(x + c3 * h, y1.data(), k3.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a41 * k1[i] + a43 * k3[i]);
    f.ode(x + c4 * h, y1.data(), k4.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a51 * k1[i] + a53 * k3[i] + a54 * k4[i]);
    f.ode(x + c5 * h, y1.data(), k5.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a61 * k1[i] + a64 * k4[i] + a65 * k5[i]);
    f.ode(x + c6 * h, y1.data(), k6.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a71 * k1[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
    f.ode(x + c7 * h, y1.data(), k7.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a81 * k1[i] + a84 * k4[i] + a85 * k5[i] + a86 * k6[i] + a87 * k7[i]);
    f.ode(x + c8 * h, y1.data(), k8.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a91 * k1[i] + a94 * k4[i] + a95 * k5[i] + a96 * k6[i] + a97 * k7[i] + a98 * k8[i]);
    f.ode(x + c9 * h, y1.data(), k9.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a101 * k1[i] + a104 * k4[i] + a105 * k5[i] + a106 * k6[i] + a107 * k7[i] +
                          a108 * k8[i] + a109 * k9[i]);
    f.ode(x + c10 * h, y1.data(), k10.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a111 * k1[i] + a114 * k4[i] + a115 * k5[i] + a116 * k6[i] + a117 * k7[i] +
                          a118 * k8[i] + a119 * k9[i] + a1110 * k10[i]);
    f.ode(x + c11 * h, y1.data(), k2.data());
    const double xph = x + h;
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a121 * k1[i] + a124 * k4[i] + a125 * k5[i] + a126 * k6[i] + a127 * k7[i] +
                          a128 * k8[i] + a129 * k9[i] + a1210 * k10[i] + a1211 * k2[i]);
    f.ode(xph, y1.data(), k3.data());
    R.nfev += 11;

    for (size_t i = 0; i < n; ++i) {
      k4[i] = b1 * k1[i] + b6 * k6[i] + b7 * k7[i] + b8 * k8[i] + b9 * k9[i] + b10 * k10[i] +
              b11 * k2[i] + b12 * k3[i];
      k5[i] = y[i] + h * k4[i];
    }

    double err = 0.0, err2 = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double sk = atol[i] + rtol[i] * std::fmax(std::fabs(y[i]), std::fabs(k5[i]));
      double erri = k4[i] - bh1 * k1[i] - bh2 * k9[i] - bh3 * k3[i];
      err2 += sq(erri / sk);
      erri = er1 * k1[i] + er6 * k6[i] + er7 * k7[i] + er8 * k8[i] + er9 * k9[i] + er10 * k10[i] +
             er11 * k2[i] + er12 * k3[i];
      err += sq(erri / sk);
    }
    double deno = err + 0.01 * err2;
    if (deno <= 0.0) deno = 1.0;
    err = std::fabs(h) * err * std::sqrt(1.0 / ((double)n * deno));

    const double fac11 = std::pow(err, expo1);
    double fac = fac11 / std::pow(facold, beta);
    fac = std::fmax(facc2, std::fmin(facc1, fac / safe));
    double hnew = h / fac;

    if (err <= 1.0) {
      facold = std::fmax(err, 1.0e-4);
      R.naccpt += 1;
      f.ode(xph, k5.data(), k4.data());
      R.nfev += 1;

      if ((R.naccpt % nstiff == 0) || (iasti > 0)) {
        double stnum = 0.0, stden = 0.0;
        for (size_t i = 0; i < n; ++i) {
          double d1 = k4[i] - k3[i], d2 = k5[i] - y1[i];
          stnum += d1 * d1; stden += d2 * d2;
        }
        if (stden > 0.0) hlamb = std::fabs(h) * std::sqrt(stnum / stden);
        if (hlamb > 6.1) {
          nonstiff = 0; iasti += 1;
          if (iasti == 15) { R.status = Status::ProbablyStiff; break; }
        } else {
          nonstiff += 1;
          if (nonstiff == 6) iasti = 0;
        }
      }

      if (dense) {
        for (size_t i = 0; i < n; ++i) {
          cont[i] = y[i];
          double ydiff = k5[i] - y[i];
          cont[n + i] = ydiff;
          double bspl = h * k1[i] - ydiff;
          cont[2 * n + i] = bspl;
          cont[3 * n + i] = ydiff - h * k4[i] - bspl;
          cont[4 * n + i] = d41 * k1[i] + d46 * k6[i] + d47 * k7[i] + d48 * k8[i] + d49 * k9[i] +
                            d410 * k10[i] + d411 * k2[i] + d412 * k3[i];
          cont[5 * n + i] = d51 * k1[i] + d56 * k6[i] + d57 * k7[i] + d58 * k8[i] + d59 * k9[i] +
                            d510 * k10[i] + d511 * k2[i] + d512 * k3[i];
          cont[6 * n + i] = d61 * k1[i] + d66 * k6[i] + d67 * k7[i] + d68 * k8[i] + d69 * k9[i] +
                            d610 * k10[i] + d611 * k2[i] + d612 * k3[i];
          cont[7 * n + i] = d71 * k1[i] + d76 * k6[i] + d77 * k7[i] + d78 * k8[i] + d79 * k9[i] +
                            d710 * k10[i] + d711 * k2[i] + d712 * k3[i];
        }
        for (size_t i = 0; i < n; ++i)
          y1[i] = y[i] + h * (a141 * k1[i] + a147 * k7[i] + a148 * k8[i] + a149 * k9[i] + a1410 * k10[i] +
                              a1411 * k2[i] + a1412 * k3[i] + a1413 * k4[i]);
        f.ode(x + c14 * h, y1.data(), k10.data());
        for (size_t i = 0; i < n; ++i)
          y1[i] = y[i] + h * (a151 * k1[i] + a156 * k6[i] + a157 * k7[i] + a158 * k8[i] + a1511 * k2[i] +
                              a1512 * k3[i] + a1513 * k4[i] + a1514 * k10[i]);
        f.ode(x + c15 * h, y1.data(), k2.data());
        for (size_t i = 0; i < n; ++i)
          y1[i] = y[i] + h * (a161 * k1[i] + a166 * k6[i] + a167 * k7[i] + a168 * k8[i] + a169 * k9[i] +
                              a1613 * k4[i] + a1614 * k10[i] + a1615 * k2[i]);
        f.ode(x + c16 * h, y1.data(), k3.data());
        R.nfev += 3;
        for (size_t i = 0; i < n; ++i) {
          cont[4 * n + i] = h * (cont[4 * n + i] + d413 * k4[i] + d414 * k10[i] + d415 * k2[i] + d416 * k3[i]);
          cont[5 * n + i] = h * (cont[5 * n + i] + d513 * k4[i] + d514 * k10[i] + d515 * k2[i] + d516 * k3[i]);
          cont[6 * n + i] = h * (cont[6 * n + i] + d613 * k4[i] + d614 * k10[i] + d615 * k2[i] + d616 * k3[i]);
          cont[7 * n + i] = h * (cont[7 * n + i] + d713 * k4[i] + d714 * k10[i] + d715 * k2[i] + d716 * k3[i]);
        }
      }

      k1 = k4; y = k5; xold = x; x = xph;

      if (so) {
        StepInterp ip{cont.data(), cont.size(), xold, h, &interpolate};
        const Flag fl = so->solout(xold, x, y, &ip);
        if (fl == Flag::Interrupt) { R.status = Status::UserInterrupt; break; }
        if (fl == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // dop853.rs:613-617, dopri5.rs:422-426
      }
      if (last) { h = hnew; R.status = Status::Success; break; }
      if (std::fabs(hnew) > std::fabs(h_max)) hnew = posneg * std::fabs(h_max);
      if (reject) { hnew = posneg * std::fmin(std::fabs(hnew), std::fabs(h)); reject = false; }
    } else {
      hnew = h / std::fmin(facc1, fac11 / safe);
      reject = true;
      if (R.naccpt > 1) R.nrejct += 1;   // dop853.rs:647 (Hairer: >= 1)
      last = false;
    }
    h = hnew;
  }
  R.h = h;
  return R;
}
}  // namespace dop853

// ======================================================================================
// DOPRI5 -- src/methods/dopri5.rs:122-520
namespace dopri5 {
static constexpr double c2 = 0.2, c3 = 0.3, c4 = 0.8, c5 = 8.0 / 9.0;
static constexpr double a21 = 0.2, a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
static constexpr double a41 = 44.0 / 45.0, a42 = -56.0 / 15.0, a43 = 32.0 / 9.0;
static constexpr double a51 = 19372.0 / 6561.0, a52 = -25360.0 / 2187.0, a53 = 64448.0 / 6561.0, a54 = -212.0 / 729.0;
static constexpr double a61 = 9017.0 / 3168.0, a62 = -355.0 / 33.0, a63 = 46732.0 / 5247.0, a64 = 49.0 / 176.0,
                        a65 = -5103.0 / 18656.0;
static constexpr double a71 = 35.0 / 384.0, a73 = 500.0 / 1113.0, a74 = 125.0 / 192.0, a75 = -2187.0 / 6784.0,
                        a76 = 11.0 / 84.0;
static constexpr double e1 = 71.0 / 57600.0, e3 = -71.0 / 16695.0, e4 = 71.0 / 1920.0, e5 = -17253.0 / 339200.0,
                        e6 = 22.0 / 525.0, e7 = -1.0 / 40.0;
static constexpr double d1 = -12715105075.0 / 11282082432.0, d3 = 87487479700.0 / 32700410799.0,
                        d4 = -10690763975.0 / 1880347072.0, d5 = 701980252875.0 / 199316789632.0,
                        d6 = -1453857185.0 / 822651844.0, d7 = 69997945.0 / 29380423.0;

inline void interpolate(double xi, double* yi, size_t n, const double* c, double xold, double h) {
  // dopri5.rs:467-478
  double th = (xi - xold) / h, th1 = 1.0 - th;
  for (size_t i = 0; i < n; ++i)
    yi[i] = c[i] + th * (c[n + i] + th1 * (c[2 * n + i] + th * (c[3 * n + i] + th1 * c[4 * n + i])));
}

template <class F, class S>
IntegrationResult solve(const F& f, double x0, const std::vector<double>& y0, double xend,
                        const Tol& rtol, const Tol& atol, const StepCfg& cfg, S* so) {
  // struct defaults dopri5.rs:33-72
  const double uround = 2.3e-16, safe = 0.9, beta = 0.04;
  const double facc1 = 1.0 / 0.2, facc2 = 1.0 / 10.0;
  const size_t nstiff = 1000;
  const bool dense = true;
  double x = x0;
  std::vector<double> y = y0;
  // dopri5.rs:180 -- NO abs on a user max_step
  const double h_max = cfg.has_max_step ? cfg.max_step : std::fabs(xend - x);
  const size_t nmax = cfg.max_steps;
  if (nmax == 0) throw ConfigError("max_steps must be positive");

  const size_t n = y.size();
  std::vector<double> k1(n), k2(n), k3(n), k4(n), k5(n), k6(n), y1(n), cont(5 * n);
  double facold = 1e-4, hlamb = 0.0;
  bool last = false, reject = false;
  int nonstiff = 0, iasti = 0;
  IntegrationResult R;
  double xold = x;
  const double expo1 = 0.2 - beta * 0.75;
  const double posneg = signum(xend - x);

  f.ode(x, y.data(), k1.data());
  R.nfev += 1;
  double h;
  if (cfg.has_first_step) h = std::fabs(cfg.first_step) * posneg;
  else { R.nfev += 1; h = hinit(f, x, y, posneg, k1, k2, k3, 5, h_max, atol, rtol); }

  if (so) {
    const Flag fl0 = so->solout(xold, x, y, nullptr);
    if (fl0 == Flag::Interrupt) {
      R.h = h; R.status = Status::UserInterrupt; return R;
    }
    if (fl0 == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // e.g. dop853.rs:258-262
  }

  for (;;) {
    if (R.nstep > nmax) { R.status = Status::NeedLargerNMax; break; }
    if (0.1 * std::fabs(h) <= std::fabs(x) * uround) { R.status = Status::StepSizeTooSmall; break; }
    if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = true; }
    R.nstep += 1;

    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * a21 * k1[i];
    f.ode(x + c2 * h, y1.data(), k2.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
    f.ode(x + c3 * h, y1.data(), k3.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
    f.ode(x + c4 * h, y1.data(), k4.data());
    for (size_t i = 0; i < n; ++i) y1[i] = y[i] + h * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
    f.ode(x + c5 * h, y1.data(), k5.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
    const double xph = x + h;
    f.ode(xph, y1.data(), k6.data());
    for (size_t i = 0; i < n; ++i)
      y1[i] = y[i] + h * (a71 * k1[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
    f.ode(xph, y1.data(), k2.data());
    R.nfev += 6;

    if (dense)
      for (size_t i = 0; i < n; ++i)
        cont[4 * n + i] = h * (d1 * k1[i] + d3 * k3[i] + d4 * k4[i] + d5 * k5[i] + d6 * k6[i] + d7 * k2[i]);

    for (size_t i = 0; i < n; ++i)
      k4[i] = (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k2[i]) * h;

    double err = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double sk = atol[i] + rtol[i] * std::fmax(std::fabs(y[i]), std::fabs(y1[i]));
      err += (k4[i] / sk) * (k4[i] / sk);
    }
    err = std::sqrt(err / (double)n);

    const double fac11 = std::pow(err, expo1);
    double fac = fac11 / std::pow(facold, beta);
    fac = std::fmax(facc2, std::fmin(facc1, fac / safe));
    double hnew = h / fac;

    if (err <= 1.0) {
      facold = std::fmax(err, 1.0e-4);
      R.naccpt += 1;
      if ((R.naccpt % nstiff == 0) || (iasti > 0)) {
        double stnum = 0.0, stden = 0.0;
        for (size_t i = 0; i < n; ++i) {
          double dd1 = k2[i] - k6[i];
          // dopri5.rs:369-370 -- uses the overwritten k4 (error vector): reference quirk
          double ysti = y[i] + h * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
          double dd2 = y1[i] - ysti;
          stnum += dd1 * dd1; stden += dd2 * dd2;
        }
        if (stden > 0.0) hlamb = std::fabs(h) * std::sqrt(stnum / stden);
        if (hlamb > 3.25) {
          nonstiff = 0; iasti += 1;
          if (iasti == 15) { R.status = Status::ProbablyStiff; break; }
        } else {
          nonstiff += 1;
          if (nonstiff == 6) iasti = 0;
        }
      }
      if (dense) {
        for (size_t i = 0; i < n; ++i) {
          double ydiff = y1[i] - y[i];
          double bspl = h * k1[i] - ydiff;
          cont[i] = y[i];
          cont[n + i] = ydiff;
          cont[2 * n + i] = bspl;
          cont[3 * n + i] = -h * k2[i] + ydiff - bspl;
        }
      }
      k1 = k2; y = y1; xold = x; x = xph;
      if (so) {
        StepInterp ip{cont.data(), cont.size(), xold, h, &interpolate};
        const Flag fl = so->solout(xold, x, y, &ip);
        if (fl == Flag::Interrupt) { R.status = Status::UserInterrupt; break; }
        if (fl == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // dop853.rs:613-617, dopri5.rs:422-426
      }
      if (last) { h = hnew; R.status = Status::Success; break; }
      if (std::fabs(hnew) > std::fabs(h_max)) hnew = posneg * std::fabs(h_max);
      if (reject) { hnew = posneg * std::fmin(std::fabs(hnew), std::fabs(h)); reject = false; }
    } else {
      hnew = h / std::fmin(facc1, fac11 / safe);
      reject = true;
      if (R.naccpt > 1) R.nrejct += 1;
      last = false;
    }
    h = hnew;
  }
  R.h = h;
  return R;
}
}  // namespace dopri5

// ======================================================================================
// RK23 -- src/methods/rk23.rs:81-347
namespace rk23 {
static constexpr double c2 = 0.5, c3 = 0.75, a21 = 0.5, a32 = 0.75;
static constexpr double b1 = 2.0 / 9.0, b2 = 1.0 / 3.0, b3 = 4.0 / 9.0;
static constexpr double e1 = 5.0 / 72.0, e2 = -1.0 / 12.0, e3 = -1.0 / 9.0, e4 = 1.0 / 8.0;
static constexpr double d21 = -4.0 / 3.0, d22 = 1.0, d23 = 4.0 / 3.0, d24 = -1.0;
static constexpr double d31 = 5.0 / 9.0, d32 = -2.0 / 3.0, d33 = -8.0 / 9.0, d34 = 1.0;

inline void interpolate(double xi, double* yi, size_t n, const double* c, double xold, double h) {
  // rk23.rs:313-321
  double xc = (xi - xold) / h, x2 = xc * xc, x3 = x2 * xc;
  for (size_t i = 0; i < n; ++i) yi[i] = c[i] + h * (c[n + i] * xc + c[2 * n + i] * x2 + c[3 * n + i] * x3);
}

template <class F, class S>
IntegrationResult solve(const F& f, double x0, const std::vector<double>& y0, double xend,
                        const Tol& rtol, const Tol& atol, const StepCfg& cfg, S* so) {
  const double safe = 0.9, scale_min = 0.2, scale_max = 10.0;   // rk23.rs:15-36
  const double error_exponent = -1.0 / 3.0;
  double x = x0;
  std::vector<double> y = y0;
  const size_t nmax = cfg.max_steps;
  if (nmax == 0) throw ConfigError("max_steps must be positive");
  const double hmax = cfg.has_max_step ? std::fabs(cfg.max_step) : std::fabs(xend - x);
  const size_t n = y.size();
  std::vector<double> k1(n), k2(n), k3(n), k4(n), yt(n), ye(n), cont(4 * n);
  IntegrationResult R;
  R.status = Status::Success;
  double xold = x;
  const double posneg = signum(xend - x);

  f.ode(x, y.data(), k1.data());
  R.nfev += 1;
  double h;
  if (cfg.has_first_step) h = std::fabs(cfg.first_step) * posneg;
  else { R.nfev += 1; h = hinit(f, x, y, posneg, k1, k2, k3, 3, hmax, atol, rtol); }
  if (so) {
    const Flag fl0 = so->solout(xold, x, y, nullptr);
    if (fl0 == Flag::Interrupt) {
      R.h = h; R.status = Status::UserInterrupt; return R;
    }
    if (fl0 == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // e.g. dop853.rs:258-262
  }
  for (;;) {
    if (R.nstep >= nmax) { R.status = Status::NeedLargerNMax; break; }
    if ((x + h - xend) * posneg > 0.0) h = xend - x;
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * a21 * k1[i];
    f.ode(x + c2 * h, yt.data(), k2.data());
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * a32 * k2[i];
    f.ode(x + c3 * h, yt.data(), k3.data());
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * (b1 * k1[i] + b2 * k2[i] + b3 * k3[i]);
    f.ode(x + h, yt.data(), k4.data());
    R.nfev += 3;
    for (size_t i = 0; i < n; ++i) ye[i] = h * (e1 * k1[i] + e2 * k2[i] + e3 * k3[i] + e4 * k4[i]);
    double err = 0.0;
    for (size_t i = 0; i < n; ++i) {
      double tol = atol[i] + rtol[i] * std::fmax(std::fabs(yt[i]), std::fabs(y[i]));
      err += sq(ye[i] / tol);
    }
    err = std::sqrt(err / (double)n);
    if (err <= 1.0) {
      R.nstep += 1; R.naccpt += 1;
      ye = y; y = yt; xold = x; x += h;
      if (so) {   // dense_output (default true) && solout.is_some()
        for (size_t i = 0; i < n; ++i) {
          cont[i] = ye[i];
          cont[n + i] = k1[i];
          cont[2 * n + i] = d21 * k1[i] + d22 * k2[i] + d23 * k3[i] + d24 * k4[i];
          cont[3 * n + i] = d31 * k1[i] + d32 * k2[i] + d33 * k3[i] + d34 * k4[i];
        }
        StepInterp ip{cont.data(), cont.size(), xold, h, &interpolate};
        const Flag fl = so->solout(xold, x, y, &ip);
        if (fl == Flag::Interrupt) { R.status = Status::UserInterrupt; break; }
        if (fl == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // rk23.rs:270-275
        else k1 = k4;  // rk23.rs:276-284 (only inside the solout branch)
      }
      h *= std::fmax(std::fmin(safe * std::pow(err, error_exponent), scale_max), scale_min);
      if (std::fabs(h) > hmax) h = hmax * posneg;
      if (x == xend) break;
    } else {
      R.nrejct += 1;
      h *= std::fmax(std::fmin(safe * std::pow(err, error_exponent), 1.0), scale_min);
    }
  }
  R.h = h;
  return R;
}
}  // namespace rk23

// ======================================================================================
// RK4 -- src/methods/rk4.rs:64-257
namespace rk4 {
static constexpr double c2 = 0.5, c3 = 0.5, c4 = 1.0, a21 = 0.5, a32 = 0.5, a43 = 1.0;
static constexpr double b1 = 1.0 / 6.0, b2 = 1.0 / 3.0, b3 = 1.0 / 3.0, b4 = 1.0 / 6.0;

inline void interpolate(double xi, double* yi, size_t n, const double* c, double xold, double h) {
  // rk4.rs:229-244 -- cubic Hermite; left slope is the k4 *stage* (reference quirk)
  double t = (xi - xold) / h, t2 = t * t, t3 = t2 * t;
  double h00 = 2.0 * t3 - 3.0 * t2 + 1.0;
  double h10 = t3 - 2.0 * t2 + t;
  double h01 = -2.0 * t3 + 3.0 * t2;
  double h11 = t3 - t2;
  for (size_t i = 0; i < n; ++i)
    yi[i] = h00 * c[i] + h10 * h * c[n + i] + h01 * c[3 * n + i] + h11 * h * c[2 * n + i];
}

template <class F, class S>
IntegrationResult solve(const F& f, double x0, const std::vector<double>& y0, double xend, double h,
                        const StepCfg& cfg, S* so) {
  double x = x0;
  std::vector<double> y = y0;
  const double posneg = signum(xend - x);
  if (h == 0.0 || signum(h) != posneg) throw ConfigError("RK4: invalid step size sign");   // rk4.rs:85
  const size_t nmax = cfg.max_steps;
  if (nmax == 0) throw ConfigError("max_steps must be positive");
  const size_t n = y.size();
  std::vector<double> k1(n), k2(n), k3(n), k4(n), yt(n), cont(4 * n);
  IntegrationResult R;
  R.status = Status::Success;
  double xold = x;
  f.ode(x, y.data(), k1.data());   // not counted (rk4.rs:116)
  if (so) {
    const Flag fl0 = so->solout(xold, x, y, nullptr);
    if (fl0 == Flag::Interrupt) {
      R.h = h; R.status = Status::UserInterrupt; return R;
    }
    if (fl0 == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // e.g. dop853.rs:258-262
  }
  for (;;) {
    if (R.nstep >= nmax) { R.status = Status::NeedLargerNMax; break; }
    bool last = false;
    if ((x + 1.01 * h - xend) * signum(h) > 0.0) last = true;
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * a21 * k1[i];
    f.ode(x + c2 * h, yt.data(), k2.data());
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * a32 * k2[i];
    f.ode(x + c3 * h, yt.data(), k3.data());
    for (size_t i = 0; i < n; ++i) yt[i] = y[i] + h * a43 * k3[i];
    f.ode(x + c4 * h, yt.data(), k4.data());
    xold = x;
    yt = y;
    x += h;
    for (size_t i = 0; i < n; ++i) y[i] += h * (b1 * k1[i] + b2 * k2[i] + b3 * k3[i] + b4 * k4[i]);
    f.ode(x, y.data(), k1.data());
    R.nfev += 4;
    R.nstep += 1;
    if (so) {
      for (size_t i = 0; i < n; ++i) {
        cont[i] = yt[i]; cont[n + i] = k4[i]; cont[2 * n + i] = k1[i]; cont[3 * n + i] = y[i];
      }
      StepInterp ip{cont.data(), cont.size(), xold, h, &interpolate};
      const Flag fl = so->solout(xold, x, y, &ip);
      if (fl == Flag::Interrupt) { R.status = Status::UserInterrupt; break; }
      if (fl == Flag::ModifiedSolution) { f.ode(x, y.data(), k1.data()); R.nfev += 1; }   // rk4.rs:206-210
    }
    if (last) break;
  }
  R.h = h;
  return R;
}
}  // namespace rk4

// ======================================================================================
// Solution -- src/solve/solution.rs:7-20 (y flattened row-major [len(t)][n])
struct DenseSeg { std::vector<double> cont; double xold, h; };
struct Solution {
  size_t n = 0;
  std::vector<double> t;
  std::vector<double> y;
  std::vector<std::vector<double>> t_events;
  std::vector<std::vector<double>> y_events;   // per event: flattened [k][n]
  size_t nfev = 0, njev = 0, nlu = 0, nstep = 0, naccpt = 0, nrejct = 0;
  Status status = Status::Success;
  double h_next = 0.0;                  // IntegrationResult.h (dropped by the reference's solve_ivp)
  double x_last = 0.0; std::vector<double> y_last;   // harness addition: integrator's last accepted (x, y)
  bool has_dense = false;
  Method method = Method::DOPRI5;
  std::vector<DenseSeg> segs;           // ContinuousOutput (src/solve/cont.rs:9-30)
};

inline InterpFn interp_fn(Method m);

// ======================================================================================
// A problem's own SolOut (src/solout.rs:55-63) behind the solvers' callback slot.  The problem supplies
//   int solout(xold, x&, y*, state*, interp, emit)  ->  0 Continue | 1 Interrupt | 2 ModifiedSolution
// `state` is the SolOut struct's own fields (NSTATE doubles, zero at the start), `interp.valid()` is false at the initial
// call (interpolant == None), `emit(t, y)` appends a sample to Solution.t / .y.
template <class F>
struct UserSolOut {
  const F& f; size_t n;
  std::vector<double> t, y, state;
  double x_last = 0.0; std::vector<double> y_last;
  UserSolOut(const F& f_, size_t n_) : f(f_), n(n_), state(F::NSTATE, 0.0) {}
  struct Interp {
    const StepInterp* ip; size_t n;
    mutable std::vector<double> buf;
    double* buffer() const { buf.resize(n); return buf.data(); }      // n doubles for eval() (device: WarpHook, ivpb_erk.cuh)
    bool valid() const { return ip != nullptr; }
    void eval(double tt, double* yi) const { ip->interpolate(tt, yi, n); }
  };
  struct Emit {
    UserSolOut* self;
    void operator()(double tt, const double* yy) { self->t.push_back(tt); self->y.insert(self->y.end(), yy, yy + self->n); }
  };
  Flag solout(double xold, double& x, std::vector<double>& yv, const StepInterp* ip) {
    Interp in{ip, n, {}};
    Emit em{this};
    const int fl = f.solout(xold, x, yv.data(), state.data(), in, em);
    x_last = x; y_last = yv;
    return fl == 1 ? Flag::Interrupt : (fl == 2 ? Flag::ModifiedSolution : Flag::Continue);
  }
};

// ======================================================================================
// DefaultSolOut -- src/solve/solout.rs:15-432
template <class F>
struct DefaultSolOut {
  const F& ode;
  bool has_t_eval; std::vector<double> t_eval; size_t next_idx = 0;
  double tol = 1e-12;
  size_t n;
  std::vector<double> t, y;
  std::vector<std::vector<double>> t_events, y_events;
  bool collect_dense; std::vector<DenseSeg> dense_segs;
  std::vector<double> yold; bool yold_set = false;
  std::vector<EventConfig> event_config;
  std::vector<double> prev_event; std::vector<size_t> event_hits;
  bool has_first_step; double first_step; double x0; bool first_output_done = false;
  std::vector<double> g_curr, y_mid, g_mid;
  // Harness addition (not in the reference): the integrator's own last accepted (x, y), so the batch
  // ABI can report t_final / y_final in t_eval mode too.
  double x_last = 0.0; std::vector<double> y_last;

  DefaultSolOut(const F& f, bool has_te, const std::vector<double>& te, bool dense, bool has_fs, double fs,
                double x0_, size_t n_states)
      : ode(f), has_t_eval(has_te), t_eval(te), n(n_states), collect_dense(dense), has_first_step(has_fs),
        first_step(fs), x0(x0_) {
    size_t ne = (size_t)f.n_events();
    for (size_t i = 0; i < ne; ++i) event_config.push_back(f.event_config((int)i));
    t_events.resize(ne); y_events.resize(ne);
    prev_event.assign(ne, 0.0); event_hits.assign(ne, 0);
    g_curr.assign(ne, 0.0); y_mid.assign(n_states, 0.0); g_mid.assign(ne, 0.0);
  }

  static bool crossed(double l, double r, Direction d) {   // solout.rs:168-176
    switch (d) {
      case Direction::All: return (l <= 0.0 && r >= 0.0) || (l >= 0.0 && r <= 0.0);
      case Direction::Positive: return l < 0.0 && r >= 0.0;
      default: return l > 0.0 && r <= 0.0;
    }
  }
  void push(double tt, const double* yy) { t.push_back(tt); y.insert(y.end(), yy, yy + n); }

  Flag solout(double xold, double& x, std::vector<double>& yv, const StepInterp* ip) {
    x_last = x; y_last = yv;
    // dense capture: solout.rs:141-146
    if (collect_dense && x != xold && ip) {
      if (ip->h != 0.0) dense_segs.push_back({std::vector<double>(ip->cont, ip->cont + ip->cont_len), ip->xold, ip->h});
    }
    const size_t ne = event_config.size();
    if (ne > 0) {
      ode.events(x, yv.data(), g_curr.data());
      if (!yold_set) {
        prev_event = g_curr;
      } else {
        struct Det { double t; size_t i; std::vector<double> y; };
        std::vector<Det> det;
        for (size_t i = 0; i < ne; ++i) {
          double g_prev = prev_event[i], g_c = g_curr[i];
          if (!crossed(g_prev, g_c, event_config[i].direction)) continue;
          const double XTOL = 2e-12, RTOL = std::numeric_limits<double>::epsilon();
          const int MAXITER = 100;
          double a = xold, b = x, fa = g_prev, fb = g_c;
          double et; std::vector<double> ey;
          if (std::fabs(fa) <= XTOL) { et = a; ey = yold; }
          else if (std::fabs(fb) <= XTOL) { et = b; ey = yv; }
          else {
            // Brent (scipy brentq port) solout.rs:204-291
            double c = a, fc = fa, d = b - a, e = d;
            for (int it = 0; it < MAXITER; ++it) {
              if (fb * fc > 0.0) { c = a; fc = fa; d = b - a; e = d; }
              if (std::fabs(fc) < std::fabs(fb)) { a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
              double tol1 = 2.0 * RTOL * std::fabs(b) + 0.5 * XTOL;
              double xm = 0.5 * (c - b);
              if (std::fabs(xm) <= tol1 || fb == 0.0) break;
              if (std::fabs(e) >= tol1 && std::fabs(fa) > std::fabs(fb)) {
                double s, p, q;
                if (a == c) {
                  s = fb / fa; p = 2.0 * xm * s; q = 1.0 - s;
                } else {
                  double qv = fa / fc, r = fb / fc;
                  s = fb / fa;
                  p = s * (2.0 * xm * qv * (qv - r) - (b - a) * (r - 1.0));
                  q = (qv - 1.0) * (r - 1.0) * (s - 1.0);
                }
                if (q > 0.0) p = -p; else q = -q;
                if (2.0 * p < std::fmin(3.0 * xm * q - std::fabs(tol1 * q), std::fabs(e * q))) { e = d; d = p / q; }
                else { d = xm; e = d; }
              } else { d = xm; e = d; }
              a = b; fa = fb;
              if (std::fabs(d) > tol1) b += d;
              else b += (xm > 0.0 ? tol1 : -tol1);
              ip->interpolate(b, y_mid.data(), n);
              ode.events(b, y_mid.data(), g_mid.data());
              fb = g_mid[i];
            }
            ip->interpolate(b, y_mid.data(), n);
            et = b; ey = y_mid;
          }
          det.push_back({et, i, ey});
        }
        bool forward = x > xold;
        // Rust sort_by is stable
        if (forward) std::stable_sort(det.begin(), det.end(), [](const Det& p, const Det& q) { return p.t < q.t; });
        else std::stable_sort(det.begin(), det.end(), [](const Det& p, const Det& q) { return p.t > q.t; });
        for (auto& dd : det) {
          t_events[dd.i].push_back(dd.t);
          y_events[dd.i].insert(y_events[dd.i].end(), dd.y.begin(), dd.y.end());
          event_hits[dd.i] += 1;
          long lim = event_config[dd.i].terminal_count;
          if (lim >= 0 && event_hits[dd.i] >= (size_t)lim) {
            push(dd.t, dd.y.data());
            prev_event = g_curr;
            return Flag::Interrupt;
          }
        }
        prev_event = g_curr;
      }
    }
    yold = yv; yold_set = true;

    if (has_t_eval) {
      size_t i = next_idx;
      if (std::fabs(xold - x) <= tol) {
        while (i < t_eval.size() && std::fabs(t_eval[i] - x) <= tol) { push(t_eval[i], yv.data()); ++i; }
      } else {
        bool forward = x > xold;
        std::vector<double> yi(n);
        if (forward) {
          while (i < t_eval.size() && t_eval[i] <= x + tol) {
            if (t_eval[i] >= xold - tol) { ip->interpolate(t_eval[i], yi.data(), n); push(t_eval[i], yi.data()); }
            ++i;
          }
        } else {
          while (i < t_eval.size() && t_eval[i] >= x - tol) {
            if (t_eval[i] <= xold + tol) { ip->interpolate(t_eval[i], yi.data(), n); push(t_eval[i], yi.data()); }
            ++i;
          }
        }
      }
      next_idx = i;
    } else {
      if (has_first_step) {
        if (!first_output_done && std::fabs(xold - x) > tol) {
          double direction = signum(x - xold);
          double target = x0 + direction * first_step;
          if (direction * (x - target) >= -tol) {
            if (ip) {
              std::vector<double> yi(n);
              ip->interpolate(target, yi.data(), n);
              push(target, yi.data());
              first_output_done = true;
            }
            if (std::fabs(x - target) > tol) push(x, yv.data());
            return Flag::Continue;
          } else {
            return Flag::Continue;
          }
        }
      }
      if (t.empty() || std::fabs(t.back() - x) > tol) push(x, yv.data());
    }
    return Flag::Continue;
  }
};

// ======================================================================================
// Options -- src/solve/options.rs:75-123 (defaults mirrored)
struct Options {
  Method method = Method::DOPRI5;
  Tol rtol = Tol(1e-3), atol = Tol(1e-6);
  bool has_max_steps = false; size_t max_steps = 0;
  bool has_t_eval = false; std::vector<double> t_eval;
  bool has_first_step = false; double first_step = 0.0;
  bool has_max_step = false; double max_step = 0.0;
  bool has_min_step = false; double min_step = 0.0;
  bool dense_output = false;
  bool user_solout = false;                // batch ABI: the problem's own SolOut replaces DefaultSolOut (Method::solve, e.g. dop853.rs:114-127)
  bool mass_full = false;                  // options.rs:109-112 (Identity | Full)
  long nind1 = -1, nind2 = -1, nind3 = -1; // options.rs:114-122 (None => -1)
};

inline size_t coeffs_per_state(Method m) {   // options.rs:34-43
  switch (m) { case Method::DOPRI5: return 5; case Method::DOP853: return 8; case Method::BDF: return 7; default: return 4; }
}

}  // namespace oracle

#include "ivp_oracle_implicit.hpp"

namespace oracle {

inline InterpFn interp_fn(Method m) {         // options.rs:49-58
  switch (m) {
    case Method::RK4: return &rk4::interpolate;
    case Method::RK23: return &rk23::interpolate;
    case Method::DOPRI5: return &dopri5::interpolate;
    case Method::DOP853: return &dop853::interpolate;
    case Method::RADAU: return &radau::interpolate;
    default: return &bdf::interpolate;
  }
}

// solve_ivp -- src/solve/solve_ivp.rs:99-313
template <class F>
Solution solve_ivp(const F& f, double x0, double xend, const std::vector<double>& y0, const Options& o) {
  Solution S;
  S.n = y0.size();
  S.method = o.method;
  const size_t ne = (size_t)f.n_events();
  S.t_events.resize(ne); S.y_events.resize(ne);
  if (std::fabs(xend - x0) < 1e-15) {          // :110-145
    if (o.has_t_eval) {
      for (double te : o.t_eval) if (std::fabs(te - x0) < 1e-12) { S.t.push_back(te); S.y.insert(S.y.end(), y0.begin(), y0.end()); }
    } else { S.t.push_back(x0); S.y = y0; }
    if (o.dense_output) {
      S.has_dense = true;
      DenseSeg seg; seg.cont.assign(S.n * coeffs_per_state(o.method), 0.0); seg.xold = x0; seg.h = 1e-15;
      if (o.method == Method::BDF) for (size_t i = 0; i < S.n; ++i) { seg.cont[i * 7] = y0[i]; seg.cont[i * 7 + 6] = 1.0; }
      else std::copy(y0.begin(), y0.end(), seg.cont.begin());
      S.segs.push_back(seg);
    }
    return S;
  }
  if (y0.empty()) {                            // :148-176
    if (o.has_t_eval) S.t = o.t_eval; else S.t = {x0, xend};
    S.has_dense = o.dense_output;
    return S;
  }
  StepCfg cfg;
  cfg.has_max_step = o.has_max_step; cfg.max_step = o.max_step;
  cfg.has_first_step = o.has_first_step; cfg.first_step = o.first_step;
  cfg.has_min_step = o.has_min_step; cfg.min_step = o.min_step;
  cfg.max_steps = o.has_max_steps ? o.max_steps : std::numeric_limits<size_t>::max();
  cfg.mass_full = o.mass_full; cfg.nind1 = o.nind1; cfg.nind2 = o.nind2; cfg.nind3 = o.nind3;
  IntegrationResult R;
  if constexpr (F::HAS_SOLOUT) {
    if (o.user_solout) {
      // The low-level entry `Method::solve(f, x0, y0, xend, rtol, atol, Some(&mut user_solout))` (dop853.rs:114-127,
      // dopri5.rs:122-135, rk23.rs:81-94, rk4.rs:64-77) with the problem's own SolOut (src/solout.rs:55-63)
      UserSolOut<F> us(f, y0.size());
      switch (o.method) {
        case Method::RK4: R = rk4::solve(f, x0, y0, xend, o.has_first_step ? o.first_step : (xend - x0) / 100.0, cfg, &us); break;
        case Method::RK23: R = rk23::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &us); break;
        case Method::DOPRI5: R = dopri5::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &us); break;
        case Method::DOP853: R = dop853::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &us); break;
        case Method::RADAU: R = radau::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &us); break;
        case Method::BDF: R = bdf::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &us); break;
      }
      S.t = std::move(us.t); S.y = std::move(us.y);
      S.nfev = R.nfev; S.njev = R.njev; S.nlu = R.nlu; S.nstep = R.nstep; S.naccpt = R.naccpt; S.nrejct = R.nrejct;
      S.status = R.status; S.h_next = R.h;
      S.x_last = us.x_last; S.y_last = std::move(us.y_last);
      return S;
    }
  }
  if (o.user_solout) throw ConfigError("user_solout = 1, but the problem defines no SolOut hook");
  DefaultSolOut<F> so(f, o.has_t_eval, o.t_eval, o.dense_output, o.has_first_step, o.first_step, x0, y0.size());
  switch (o.method) {
    case Method::RK4: {
      double h = o.has_first_step ? o.first_step : (xend - x0) / 100.0;   // :185
      R = rk4::solve(f, x0, y0, xend, h, cfg, &so);
      break;
    }
    case Method::RK23: R = rk23::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &so); break;
    case Method::DOPRI5: R = dopri5::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &so); break;
    case Method::DOP853: R = dop853::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &so); break;
    case Method::RADAU: R = radau::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &so); break;
    case Method::BDF: R = bdf::solve(f, x0, y0, xend, o.rtol, o.atol, cfg, &so); break;
  }
  S.t = std::move(so.t); S.y = std::move(so.y);
  S.t_events = std::move(so.t_events); S.y_events = std::move(so.y_events);
  S.nfev = R.nfev; S.njev = R.njev; S.nlu = R.nlu;
  S.nstep = R.nstep; S.naccpt = R.naccpt; S.nrejct = R.nrejct;
  S.status = R.status; S.h_next = R.h;
  S.x_last = so.x_last; S.y_last = std::move(so.y_last);
  if (o.dense_output) {
    S.has_dense = true;
    for (auto& sg : so.dense_segs) if (sg.h != 0.0) S.segs.push_back(std::move(sg));   // cont.rs:15-28
  }
  return S;
}

// ContinuousOutput::evaluate / t_span -- src/solve/cont.rs:66-117; Solution::sol -- solution.rs:25-44
inline bool sol_span(const Solution& S, double& a, double& b) {
  if (!S.has_dense || S.segs.empty()) return false;
  a = S.segs.front().xold; b = S.segs.back().xold + S.segs.back().h; return true;
}
inline bool sol_eval(const Solution& S, double t, double* yi) {
  double a, b;
  if (!sol_span(S, a, b)) return false;
  double lo = std::fmin(a, b), hi = std::fmax(a, b);
  if (t < lo || t > hi) return false;
  const double tol = 1e-12;
  for (const auto& sg : S.segs) {
    double left = std::fmin(sg.xold, sg.xold + sg.h), right = std::fmax(sg.xold, sg.xold + sg.h);
    if (t >= left - tol && t <= right + tol) { interp_fn(S.method)(t, yi, S.n, sg.cont.data(), sg.xold, sg.h); return true; }
  }
  return false;
}

// ContinuousOutput::evaluate_extrapolate / find_segment_extrapolate -- src/solve/cont.rs:91-150 (the rule behind the
// reference's Python OdeSolution.__call__, src/python/solution.rs:41,116)
inline bool sol_eval_extrapolate(const Solution& S, double t, double* yi) {
  if (!S.has_dense || S.segs.empty()) return false;
  const double tol = 1e-12;
  const DenseSeg* use = nullptr;
  for (const auto& sg : S.segs) {
    double left = std::fmin(sg.xold, sg.xold + sg.h), right = std::fmax(sg.xold, sg.xold + sg.h);
    if (t >= left - tol && t <= right + tol) { use = &sg; break; }
  }
  if (!use) {
    const DenseSeg& first = S.segs.front();
    const DenseSeg& last = S.segs.back();
    const double first_left = std::fmin(first.xold, first.xold + first.h);
    const double last_right = std::fmax(last.xold, last.xold + last.h);
    if (t < first_left) use = &first;
    else if (t > last_right) use = &last;
    else return false;
  }
  interp_fn(S.method)(t, yi, S.n, use->cont.data(), use->xold, use->h);
  return true;
}

}  // namespace oracle
