// C++ host-API tests: the reference's own integration tests (tests/ivp.rs, tests/accuracy.rs,
// tests/backward_and_bounds.rs, examples/*.rs) restated against ivp::solve_ivp / ivp::solve_ivp_batch.
// Built and run by tests/test_cpp_host_api.py.   usage: test_ivp_batch [--api-only]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ivp_batch.hpp"

using namespace ivp;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...)                                                                  \
  do {                                                                                    \
    ++g_checks;                                                                           \
    if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d: %s -- ", __FILE__, __LINE__, #cond); std::printf(__VA_ARGS__); std::printf("\n"); } \
  } while (0)

static std::vector<Method> all_methods() { return {Method::RK23, Method::DOPRI5, Method::DOP853, Method::RADAU, Method::BDF}; }
static const char* name(Method m) { static const char* n[] = {"RK23", "DOPRI5", "DOP853", "RK4", "RADAU", "BDF"}; return n[(int)m]; }
static Options default_opts(Method m) { return Options::builder().method(m).rtol(1e-9).atol(1e-9).build(); }   // tests/common.rs:21-27

// ---- API surface only: runs without a GPU ----------------------------------------------------------------
static void api_surface() {
  Options d = Options::builder().build();                       // options.rs:77-123 defaults
  CHECK(d.method == Method::DOPRI5 && d.rtol[0] == 1e-3 && d.atol[0] == 1e-6, "defaults");
  CHECK(!d.max_steps && !d.t_eval && !d.first_step && !d.max_step && !d.min_step && !d.dense_output, "Option fields default to None");
  CHECK(method_from_str("rk45") == Method::DOPRI5 && method_from_str("Radau5") == Method::RADAU &&
        method_from_str("bdf15") == Method::BDF && method_from_str("nonsense") == Method::DOPRI5, "From<&str> for Method");
  CHECK(coeffs_per_state(Method::DOP853) == 8 && coeffs_per_state(Method::BDF) == 7 && coeffs_per_state(Method::RK4) == 4, "coeffs_per_state");
  CHECK((int)Status::Success == 0 && (int)Status::UserInterrupt == 1 && (int)Status::PoorConvergence == 6, "Status declaration order");
  CHECK(direction_from(3) == Direction::Positive && direction_from(-1) == Direction::Negative && direction_from(0) == Direction::All, "From<i32> for Direction");
  EventConfig c; c.negative().terminal();
  CHECK(c.direction == Direction::Negative && c.terminal_count_ && *c.terminal_count_ == 1, "EventConfig setters");
  Tolerance tv{1e-2, 1e-10};
  CHECK(tv.is_vector() && tv[1] == 1e-10 && Tolerance(1e-3)[5] == 1e-3, "Tolerance scalar broadcast / vector index");
  Problem p = Problem::builtin("cr3bp");
  CHECK(p.n() == 6 && p.n_params() == 1 && p.n_events() == 0, "built-in problem sizes");
  bool threw = false;
  try { Problem::builtin("no_such_problem"); } catch (const ConfigError&) { threw = true; }
  CHECK(threw, "unknown built-in problem is a ConfigError");
  // options.rs:105-122 (mass_storage, nind1..3) and the SolOut slot of Method::solve
  Options m = Options::builder().method(Method::RADAU).mass_storage(Options::MatrixStorage::Full).nind2(1).user_solout(true).build();
  CHECK(d.mass_storage == Options::MatrixStorage::Identity && !d.nind1 && !d.nind2 && !d.nind3 && !d.user_solout, "mass / DAE / hook defaults");
  CHECK(m.mass_storage == Options::MatrixStorage::Full && m.nind2 && *m.nind2 == 1 && !m.nind1 && m.user_solout, "mass / DAE / hook setters");
  Problem dae = Problem::builtin("robertson_dae"), bounce = Problem::builtin("ball_bounce");
  CHECK(dae.n() == 3 && dae.n_params() == 3 && bounce.n() == 2 && bounce.n_params() == 3, "built-ins of the widened rows");
}

static void no_device_fails_loudly() {
  bool threw = false;
  try { Context c; } catch (const DeviceError& e) { threw = std::strstr(e.what(), "no CPU fallback") != nullptr; }
  CHECK(threw, "without a CUDA device the context must fail loudly");
}

// ---- tests/ivp.rs:20-46 ----------------------------------------------------------------------------------
static void integration_zero_rhs_all_methods() {
  Problem f = Problem::builtin("zero3");
  std::vector<double> t_eval;
  for (int i = 0; i <= 20; ++i) t_eval.push_back(0.0 + 10.0 * i / 20.0);
  for (Method m : {Method::RK23, Method::DOPRI5, Method::DOP853, Method::RADAU}) {      // BDF excluded, as in the reference
    Options o = Options::builder().method(m).rtol(1e-9).atol(1e-12).t_eval(t_eval).build();
    Solution sol = solve_ivp(f, 0.0, 10.0, {1.0, 1.0, 1.0}, o);
    CHECK(sol.t == t_eval, "%s: sol.t == t_eval", name(m));
    for (auto& yi : sol.y) for (double v : yi) CHECK(std::fabs(v - 1.0) <= 1e-12, "%s", name(m));
  }
}

// ---- tests/ivp.rs:48-104 ---------------------------------------------------------------------------------
static void max_step_and_first_step_controls() {
  Problem sho = Problem::builtin("sho");
  const double max_step = 0.05;
  for (Method m : all_methods()) {
    Options o = Options::builder().method(m).rtol(1e-6).atol(1e-9).max_step(max_step).build();
    Solution sol = solve_ivp(sho, 0.0, 3.0, {1.0, 0.0}, o);
    CHECK(sol.status == Status::Success && sol.t.size() > 10, "%s", name(m));
    for (size_t i = 1; i < sol.t.size(); ++i)
      CHECK(std::fabs(sol.t[i] - sol.t[i - 1]) <= max_step * 1.01 + 1e-12, "dt exceeds max_step for %s", name(m));
  }
  for (Method m : {Method::RK23, Method::DOPRI5, Method::DOP853, Method::RADAU}) {
    Options o = Options::builder().method(m).rtol(1e-3).atol(1e-6).first_step(0.1).build();
    Solution sol = solve_ivp(sho, 0.0, 3.0, {1.0, 0.0}, o);
    CHECK(sol.t.size() >= 2, "%s", name(m));
    if (sol.t.size() >= 2) CHECK(std::fabs(std::fabs(sol.t[1] - sol.t[0]) - 0.1) <= 1e-6, "first step mismatch for %s: %g", name(m), sol.t[1] - sol.t[0]);
  }
}

// ---- tests/ivp.rs:151-275 --------------------------------------------------------------------------------
static void event_detection_all_and_directional() {
  Problem sho = Problem::builtin("sho");     // events: g = y[0]
  const double pi_2 = std::acos(-1.0) / 2;
  auto run = [&](EventConfig c) {
    Options o = default_opts(Method::DOPRI5);
    o.event_config = std::vector<EventConfig>{c};
    return solve_ivp(sho, 0.0, 6.0, {1.0, 0.0}, o);
  };
  Solution all = run(EventConfig().all().terminal_count(2));
  CHECK(all.t_events[0].size() == 2, "two zero crossings recorded");
  if (all.t_events[0].size() == 2) {
    CHECK(std::fabs(all.t_events[0][0] - pi_2) < 5e-3 && std::fabs(all.t_events[0][1] - 3 * pi_2) < 5e-3, "zeros at pi/2, 3pi/2");
    CHECK(std::fabs(all.y_events[0][0][0]) <= 1e-8 && std::fabs(all.y_events[0][1][0]) <= 1e-8, "event states on the zero");
    CHECK(all.status == Status::UserInterrupt && all.t.back() == all.t_events[0][1], "terminal event point appended to t");
  }
  Solution pos = run(EventConfig().positive().terminal());
  CHECK(!pos.t_events[0].empty() && std::fabs(pos.t_events[0][0] - 3 * pi_2) < 5e-3, "positive crossing near 3pi/2");
  Solution neg = run(EventConfig().negative().terminal());
  CHECK(!neg.t_events[0].empty() && std::fabs(neg.t_events[0][0] - pi_2) < 5e-3, "negative crossing near pi/2");
}

// ---- tests/ivp.rs:277-289 --------------------------------------------------------------------------------
static void zero_interval_returns_initial_state() {
  Problem sho = Problem::builtin("sho");
  for (Method m : all_methods()) {
    Solution sol = solve_ivp(sho, 1.23, 1.23, {2.0, 3.0}, default_opts(m));
    CHECK(!sol.t.empty(), "%s", name(m));
    CHECK(std::fabs(sol.y.back()[0] - 2.0) <= 1e-12 && std::fabs(sol.y.back()[1] - 3.0) <= 1e-12, "%s", name(m));
  }
}

// ---- tests/ivp.rs:299-334 --------------------------------------------------------------------------------
static void vector_rtol_componentwise_control() {
  Problem f = Problem::builtin("exp2");
  Options loose = Options::builder().method(Method::DOPRI5).rtol({1e-2, 1e-2}).atol(1e-10).build();
  Options tight = Options::builder().method(Method::DOPRI5).rtol({1e-2, 1e-10}).atol(1e-10).build();
  Solution a = solve_ivp(f, 0.0, 1.0, {1.0, 1.0}, loose), b = solve_ivp(f, 0.0, 1.0, {1.0, 1.0}, tight);
  const double e = std::exp(1.0);
  CHECK(std::fabs(b.y.back()[1] - e) < 0.5 * std::fabs(a.y.back()[1] - e), "tightening the second component reduces its error");
  CHECK(std::fabs(b.y.back()[0] - e) <= 10.0 * std::fabs(a.y.back()[0] - e), "first component not dramatically worse");
}

// ---- tests/accuracy.rs:17-76 -----------------------------------------------------------------------------
static void harmonic_accuracy_and_t_eval() {
  Problem sho = Problem::builtin("sho");
  const double xend = 2.0 * std::acos(-1.0);
  for (Method m : {Method::RK4, Method::RK23, Method::DOPRI5, Method::DOP853, Method::RADAU, Method::BDF}) {
    Options o = m == Method::RK4 ? Options::builder().method(m).first_step(xend / 2000.0).build() : default_opts(m);
    Solution sol = solve_ivp(sho, 0.0, xend, {1.0, 0.0}, o);
    CHECK(std::fabs(sol.y.back()[0] - 1.0) < 1e-5 && std::fabs(sol.y.back()[1]) < 1e-5, "end state after one period, %s", name(m));
    std::vector<double> te;
    for (int i = 0; i <= 10; ++i) te.push_back(i / 10.0);
    Options ot = m == Method::RK4 ? Options::builder().method(m).first_step(0.01).t_eval(te).build()
                                  : Options::builder().method(m).rtol(1e-9).atol(1e-9).t_eval(te).build();
    Solution st = solve_ivp(sho, 0.0, 1.0, {1.0, 0.0}, ot);
    CHECK(st.t == te && st.y.size() == st.t.size(), "t_eval respected for %s", name(m));
    size_t k = 0;
    for (auto [t, y] : st) { CHECK(std::fabs(y[0] - std::cos(t)) < 1e-4, "%s sample %zu", name(m), k); ++k; }   // Solution::iter
  }
}

// ---- tests/backward_and_bounds.rs:6-31 (end state instead of the dense span) --------------------------------
static void backward_integration_works() {
  Problem sho = Problem::builtin("sho");
  const double x0 = 2.0 * std::acos(-1.0);
  for (Method m : all_methods()) {
    Options o = Options::builder().method(m).rtol(1e-9).atol(1e-9).t_eval({x0, 0.5 * x0, 0.0}).build();
    Solution sol = solve_ivp(sho, x0, 0.0, {1.0, 0.0}, o);
    CHECK(sol.status == Status::Success && sol.t.size() == 3, "%s", name(m));
    if (sol.t.size() == 3) CHECK(std::fabs(sol.y[1][0] - std::cos(0.5 * x0)) < 1e-6 && std::fabs(sol.y[1][1] + std::sin(0.5 * x0)) < 1e-6, "%s mid point", name(m));
  }
}

// ---- tests/ivp.rs:106-149, tests/backward_and_bounds.rs:6-31 (dense output) -------------------------------
static void dense_output_tests() {
  Problem sho = Problem::builtin("sho");
  for (Method m : {Method::RK23, Method::DOPRI5, Method::DOP853, Method::RADAU}) {
    Options o = Options::builder().method(m).rtol(1e-8).atol(1e-10).dense_output(true).build();
    Solution sol = solve_ivp(sho, 0.0, 2.0, {1.0, 0.0}, o);
    CHECK(sol.sol_span().has_value(), "%s: span exists", name(m));
    auto ys = sol.sol_many(sol.t);
    CHECK(ys.size() == sol.y.size(), "%s", name(m));
    for (size_t k = 0; k < ys.size(); ++k)
      for (size_t c = 0; c < ys[k].size(); ++c) CHECK(std::fabs(ys[k][c] - sol.y[k][c]) <= 1e-8, "dense vs stored mismatch (%s)", name(m));
  }
  {
    Solution sol = solve_ivp(sho, 0.0, 1.0, {1.0, 0.0}, Options::builder().method(Method::DOPRI5).rtol(1e-9).atol(1e-9).dense_output(true).build());
    auto span = *sol.sol_span();
    int errs = 0;
    try { sol.sol(span.first - 0.1); } catch (const InterpolationError&) { ++errs; }
    try { sol.sol(span.second + 0.1); } catch (const InterpolationError&) { ++errs; }
    CHECK(errs == 2, "out-of-range dense evaluation is an error");
    Solution plain = solve_ivp(sho, 0.0, 1.0, {1.0, 0.0}, default_opts(Method::DOPRI5));
    bool threw = false;
    try { plain.sol(0.5); } catch (const InterpolationError&) { threw = true; }
    CHECK(threw && !plain.sol_span(), "dense output disabled => NotEnabled");
  }
  const double x0 = 2.0 * std::acos(-1.0);
  for (Method m : all_methods()) {
    Solution sol = solve_ivp(sho, x0, 0.0, {1.0, 0.0}, Options::builder().method(m).rtol(1e-9).atol(1e-9).dense_output(true).build());
    auto span = sol.sol_span();
    CHECK(span && span->first > span->second, "%s: backward span", name(m));
    if (span) {
      const double mid = 0.5 * (span->first + span->second);
      auto y = sol.sol(mid);
      CHECK(std::fabs(y[0] - std::cos(mid)) < 1e-6 && std::fabs(y[1] + std::sin(mid)) < 1e-6, "%s: dense mid point", name(m));
    }
  }
}

// ---- examples/bouncing_ball.rs, examples/van_der_pol.rs, examples/cr3bp.rs as batches ----------------------
static void example_programs_as_batches() {
  const size_t N = 1000;
  {  // bouncing ball: terminal event y[0] = 0, negative direction (bouncing_ball.rs:11-31)
    Problem ball = Problem::builtin("bouncing_ball");
    std::vector<double> y0, par;
    for (size_t i = 0; i < N; ++i) { y0.push_back(1.0 + 0.01 * i); y0.push_back(0.0); par.push_back(9.81); par.push_back(0.0); }
    auto sols = solve_ivp_batch(ball, 0.0, 10.0, y0, par, Options::builder().method(Method::DOPRI5).rtol(1e-8).atol(1e-10).build());
    for (size_t i = 0; i < N; i += 111) {
      const double t_hit = std::sqrt(2.0 * (1.0 + 0.01 * i) / 9.81);
      CHECK(sols[i].status == Status::UserInterrupt && sols[i].t_events[0].size() == 1, "ball %zu", i);
      CHECK(std::fabs(sols[i].t_events[0][0] - t_hit) < 1e-7, "impact time %g vs %g", sols[i].t_events[0][0], t_hit);
    }
  }
  {  // stiff Van der Pol, eps = 1e-3, BDF, t_eval 0..2 (van_der_pol.rs:17-29), whole batch identical rows
    Problem vdp = Problem::builtin("vdp_eps");
    std::vector<double> y0, par, te;
    for (size_t i = 0; i < 64; ++i) { y0.push_back(2.0); y0.push_back(0.0); par.push_back(1e-3); }
    for (int i = 0; i <= 20; ++i) te.push_back(0.1 * i);
    for (Method m : {Method::BDF, Method::RADAU}) {
      auto sols = solve_ivp_batch(vdp, 0.0, 2.0, y0, par, Options::builder().method(m).rtol(1e-6).atol(1e-8).t_eval(te).build());
      CHECK(sols[0].status == Status::Success && sols[0].t == te && sols[0].njev > 0 && sols[0].nlu > 0, "%s", name(m));
      for (auto& s : sols) CHECK(s.y == sols[0].y && s.naccpt == sols[0].naccpt, "identical rows give identical solutions (%s)", name(m));
      CHECK(std::fabs(sols[0].y.back()[0] - 1.7632) < 1e-3, "y(2) on the slow branch (%s): %g", name(m), sols[0].y.back()[0]);
    }
  }
  {  // user problem as CUDA C through NVRTC == `impl IVP for Decay` (exponential_decay.rs:10-12)
    Problem decay = Problem::from_cuda_source(
        "__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = -p[0] * y[0]; }", 1, 1);
    std::vector<double> y0(N, 1.0), k(N);
    for (size_t i = 0; i < N; ++i) k[i] = 0.1 + 0.001 * i;
    auto sols = solve_ivp_batch(decay, 0.0, 5.0, y0, k, Options::builder().method(Method::DOP853).rtol(1e-10).atol(1e-12).build());
    for (size_t i = 0; i < N; i += 97) CHECK(std::fabs(sols[i].y.back()[0] - std::exp(-5.0 * k[i])) < 1e-9, "decay %zu", i);
    bool threw = false;
    try { solve_ivp(decay, 0.0, 1.0, {1.0}, Options::builder().method(Method::RK4).first_step(-0.1).build(), {0.5}); }
    catch (const ConfigError&) { threw = true; }                  // rk4.rs:84-90
    CHECK(threw, "RK4 with a step of the wrong sign is Error::Config");
    threw = false;
    try { solve_ivp(decay, 0.0, 1.0, {1.0}, Options::builder().max_steps(0).build(), {0.5}); } catch (const ConfigError&) { threw = true; }
    CHECK(threw, "max_steps == 0 is Error::Config");
    Solution s = solve_ivp(decay, 0.0, 1.0, {1.0}, Options::builder().rtol(1e-10).atol(1e-12).max_steps(3).build(), {0.5});
    CHECK(s.status == Status::NeedLargerNMax, "numerical failure is a status, not an error: %s", to_string(s.status));
  }
}

// ---- RADAU with a mass matrix, SolOut hooks, extrapolating dense output ---------------------
static void mass_matrix_hooks_and_extrapolation() {
  {  // Robertson as an index-1 DAE (M = diag(1, 1, 0)) against its ODE form, radau.rs:375-386,525-539,626-634
    Problem dae = Problem::builtin("robertson_dae"), ode = Problem::builtin("robertson");
    const std::vector<double> y0 = {1.0, 0.0, 0.0}, par = {0.04, 1e4, 3e7};
    Options od = Options::builder().method(Method::RADAU).rtol(1e-8).atol(1e-12).mass_storage(Options::MatrixStorage::Full).build();
    Options oo = Options::builder().method(Method::RADAU).rtol(1e-8).atol(1e-12).build();
    Solution a = solve_ivp(dae, 0.0, 1e5, y0, od, par), b = solve_ivp(ode, 0.0, 1e5, y0, oo, par);
    CHECK(a.status == Status::Success && b.status == Status::Success, "DAE / ODE status");
    for (int i = 0; i < 3; ++i)
      CHECK(std::fabs(a.y.back()[i] - b.y.back()[i]) <= 1e-6 * std::fabs(b.y.back()[i]) + 1e-11, "DAE == ODE form, component %d", i);
    CHECK(std::fabs(a.y.back()[0] + a.y.back()[1] + a.y.back()[2] - 1.0) < 1e-10, "the algebraic row holds");
    bool threw = false;
    try { solve_ivp(dae, 0.0, 1.0, y0, oo, par); } catch (const ConfigError&) { threw = true; }
    CHECK(threw, "a mass-matrix problem without mass_storage = Full is Error::Config");
    threw = false;
    try { solve_ivp(dae, 0.0, 1.0, y0, Options::builder().method(Method::RADAU).mass_storage(Options::MatrixStorage::Full).nind1(1).nind2(1).build(), par); }
    catch (const ConfigError&) { threw = true; }                  // radau.rs:236-245 InvalidDAEPartition
    CHECK(threw, "nind1 + nind2 + nind3 != n is Error::Config");
  }
  {  // the ball that bounces inside one solve: SolOut + ControlFlag::ModifiedSolution (src/solout.rs:18-29)
    Problem bounce = Problem::builtin("ball_bounce");
    for (Method m : {Method::DOPRI5, Method::DOP853, Method::RADAU, Method::BDF}) {
      Solution s = solve_ivp(bounce, 0.0, 15.0, {10.0, 5.0},
                             Options::builder().method(m).rtol(1e-8).atol(1e-10).user_solout(true).build(), {9.81, 0.02, 0.75});
      CHECK(s.status == Status::UserInterrupt && s.t.size() == 18, "%s: 17 impacts + the start point, then |v| < 0.1 stops it (%zu)", name(m), s.t.size());
      CHECK(std::fabs(s.t[1] - 2.0726) < 1e-3 && std::fabs(s.t.back() - 9.1664) < 1e-3, "%s: first / last impact %g %g", name(m), s.t[1], s.t.back());
    }
  }
  {  // ContinuousOutput::evaluate_extrapolate (cont.rs:91-150)
    Problem sho = Problem::builtin("sho");
    Solution s = solve_ivp(sho, 0.0, 1.0, {1.0, 0.0}, Options::builder().method(Method::DOP853).rtol(1e-10).atol(1e-12).dense_output(true).build());
    auto in = s.sol_extrapolate(0.5), out = s.sol_extrapolate(1.05);
    CHECK(in && std::fabs((*in)[0] - std::cos(0.5)) < 1e-9, "inside the span: interpolation");
    CHECK(out && std::fabs((*out)[0] - std::cos(1.05)) < 1e-6, "beyond the span: the last segment extrapolates");
    bool threw = false;
    try { s.sol(1.05); } catch (const InterpolationError&) { threw = true; }
    CHECK(threw, "Solution::sol outside the span stays an error (solution.rs:25-44)");
  }
}

int main(int argc, char** argv) {
  const bool api_only = argc > 1 && std::strcmp(argv[1], "--api-only") == 0;
  api_surface();
  if (api_only) {
    if (argc > 2 && std::strcmp(argv[2], "--expect-no-device") == 0) no_device_fails_loudly();
  } else {
    integration_zero_rhs_all_methods();
    max_step_and_first_step_controls();
    event_detection_all_and_directional();
    zero_interval_returns_initial_state();
    vector_rtol_componentwise_control();
    harmonic_accuracy_and_t_eval();
    backward_integration_works();
    dense_output_tests();
    example_programs_as_batches();
    mass_matrix_hooks_and_extrapolation();
  }
  std::printf("%d checks, %d failed\n", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
