// ivp_batch.hpp -- C++17 host API above the C ABI (include/ivpb.h): the reference crate's API surface for
// batched solves.  The reference is Rust (crate `ivp` v0.5.1); neither rustc nor cargo exists in this image,
// so the host side the north star asks for is written in C++ with the reference's names, argument meaning,
// defaults and error behaviour, and the equivalent Rust binding is kept as source under rust/ (INTEGRATION.md).
//
//   reference item                                   here
//   Method              src/solve/options.rs:14-73   ivp::Method, ivp::method_from_str, coeffs_per_state
//   Options::builder()  src/solve/options.rs:75-123  ivp::Options::builder() ... .build()
//   Tolerance           src/methods/mod.rs:104-214   ivp::Tolerance (scalar or per-component vector)
//   EventConfig         src/solve/event.rs:5-77      ivp::EventConfig, ivp::Direction
//   Status              src/status.rs:4-26           ivp::Status (same declaration order == device codes)
//   Solution            src/solve/solution.rs:7-97   ivp::Solution (t, y, t_events, y_events, counters, status)
//   Error::Config       src/error.rs:7-80            ivp::ConfigError (thrown before stepping)
//   trait IVP           src/ivp.rs:27-121            ivp::Problem: a built-in device problem, or CUDA C source
//                                                    defining ivp_ode / ivp_events / ivp_jac (NVRTC)
//   solve_ivp           src/solve/solve_ivp.rs:99    ivp::solve_ivp (one trajectory) and the new
//                                                    ivp::solve_ivp_batch(problem, t0, tf, Y0[N x n], params[N x p], options)
//
// Header-only; link with -livpb.  Every solve runs on the GPU through ivpb_solve_batch -- there is no CPU path.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "ivpb.h"

namespace ivp {

using Float = double;   // reference src/lib.rs:84-85 (f64 default)

// ---- errors (reference src/error.rs) ----------------------------------------------------------
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
struct ConfigError : Error { using Error::Error; };          // Error::Config(..)
struct DeviceError : Error { using Error::Error; };          // CUDA / NVRTC failures (no reference equivalent)
struct InterpolationError : Error { using Error::Error; };   // Error::Interpolation(..)

// ---- Method (reference src/solve/options.rs:14-73) ---------------------------------------------
enum class Method : int32_t { RK23 = 0, DOPRI5 = 1, DOP853 = 2, RK4 = 3, RADAU = 4, BDF = 5 };

inline Method method_from_str(std::string s) {   // impl From<&str> for Method; unknown => DOPRI5
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
  if (s == "RK23") return Method::RK23;
  if (s == "DOPRI5" || s == "RK45") return Method::DOPRI5;
  if (s == "DOP853") return Method::DOP853;
  if (s == "RK4") return Method::RK4;
  if (s == "RADAU" || s == "RADAU5") return Method::RADAU;
  if (s == "BDF" || s == "BDF15") return Method::BDF;
  return Method::DOPRI5;
}
constexpr size_t coeffs_per_state(Method m) {    // Method::coeffs_per_state
  return m == Method::DOPRI5 ? 5 : m == Method::DOP853 ? 8 : m == Method::BDF ? 7 : 4;
}

// ---- Status (reference src/status.rs:4-19) -----------------------------------------------------
enum class Status : int32_t {
  Success = 0, UserInterrupt = 1, NeedLargerNMax = 2, StepSizeTooSmall = 3, ProbablyStiff = 4,
  SingularMatrix = 5, PoorConvergence = 6
};
inline const char* to_string(Status s) {
  static const char* names[] = {"Success", "UserInterrupt", "NeedLargerNMax", "StepSizeTooSmall",
                                "ProbablyStiff", "SingularMatrix", "PoorConvergence"};
  return names[(int)s];
}

// ---- Tolerance (reference src/methods/mod.rs:104-214) ------------------------------------------
class Tolerance {
 public:
  Tolerance(Float v) : v_{v} {}                                  // From<Float>
  Tolerance(std::vector<Float> v) : v_(std::move(v)), vec_(true) {}   // From<Vec<Float>>
  Tolerance(std::initializer_list<Float> v) : v_(v), vec_(true) {}    // From<[Float; N]>
  bool is_vector() const { return vec_; }
  size_t len() const { return v_.size(); }
  Float operator[](size_t i) const { return vec_ ? v_.at(i) : v_[0]; }   // impl Index: scalar broadcasts
  const Float* data() const { return v_.data(); }
 private:
  std::vector<Float> v_;
  bool vec_ = false;
};

// ---- events (reference src/solve/event.rs) -----------------------------------------------------
enum class Direction : int32_t { All = 0, Positive = 1, Negative = -1 };
inline Direction direction_from(int v) { return v > 0 ? Direction::Positive : v < 0 ? Direction::Negative : Direction::All; }
struct EventConfig {
  Direction direction = Direction::All;
  std::optional<size_t> terminal_count_;       // None => never terminate
  EventConfig& terminal_count(size_t n) { terminal_count_ = n; return *this; }
  EventConfig& terminal() { terminal_count_ = 1; return *this; }
  EventConfig& all() { direction = Direction::All; return *this; }
  EventConfig& positive() { direction = Direction::Positive; return *this; }
  EventConfig& negative() { direction = Direction::Negative; return *this; }
};

// ---- Options (reference src/solve/options.rs:75-123) -------------------------------------------
class OptionsBuilder;
struct Options {
  Method method = Method::DOPRI5;
  Tolerance rtol = 1e-3;
  Tolerance atol = 1e-6;
  std::optional<size_t> max_steps;
  std::optional<std::vector<Float>> t_eval;
  std::optional<Float> first_step, max_step, min_step;
  bool dense_output = false;
  int max_segments = 4096;   // dense_output: interpolant segments kept per trajectory (one per accepted step)
  // mass_storage / nind1-3 (options.rs:105-122; RADAU only, solve_ivp.rs:246-258): Full => M y' = f with the problem's
  // constant mass matrix (IVP::mass), nind2 / nind3 => index-2 / index-3 variable counts of a DAE.  jac_storage of the
  // reference selects a Banded Jacobian layout; device Jacobians are always Full.
  enum class MatrixStorage { Identity, Full };
  MatrixStorage mass_storage = MatrixStorage::Identity;
  std::optional<size_t> nind1, nind2, nind3;
  // the problem's own SolOut hook (src/solout.rs:55-63) instead of DefaultSolOut: the reference's low-level
  // Method::solve(.., Some(&mut solout)) call.  Samples the hook emits come back in Solution::t / y.
  bool user_solout = false;
  // ---- batched-solve additions (no reference equivalent) ----
  std::optional<std::vector<EventConfig>> event_config;   // overrides the problem's IVP::event_config
  int max_events = 8;        // event hits stored per event function and trajectory
  int max_out = 4096;        // step-mode samples stored per trajectory when t_eval is None (0: endpoints only)
  bool analytic_jac = false; // use the problem's ivp_jac instead of the default finite differences
  // `jac_sparsity` of the reference's Python front end (src/python/sparsity.rs): structural non-zeros as (row, col) pairs;
  // RADAU / BDF then take one RHS evaluation per group of structurally orthogonal columns instead of one per column.
  std::optional<std::vector<std::pair<int, int>>> jac_sparsity;
  bool strict_fp = false;    // IVPB_FLAG_STRICT_FP: the reference's rounding, operation for operation (explicit methods;
                             // RADAU / BDF use it by default)
  bool fast_fp = false;      // IVPB_FLAG_FAST_FP: FMA-contracted RADAU / BDF kernels
  static OptionsBuilder builder();
};
class OptionsBuilder {
 public:
  OptionsBuilder& method(Method m) { o_.method = m; return *this; }
  OptionsBuilder& method(const char* s) { o_.method = method_from_str(s); return *this; }
  OptionsBuilder& rtol(Tolerance t) { o_.rtol = std::move(t); return *this; }
  OptionsBuilder& atol(Tolerance t) { o_.atol = std::move(t); return *this; }
  OptionsBuilder& max_steps(size_t n) { o_.max_steps = n; return *this; }
  OptionsBuilder& t_eval(std::vector<Float> t) { o_.t_eval = std::move(t); return *this; }
  OptionsBuilder& first_step(Float h) { o_.first_step = h; return *this; }
  OptionsBuilder& max_step(Float h) { o_.max_step = h; return *this; }
  OptionsBuilder& min_step(Float h) { o_.min_step = h; return *this; }
  OptionsBuilder& dense_output(bool b) { o_.dense_output = b; return *this; }
  OptionsBuilder& max_segments(int n) { o_.max_segments = n; return *this; }
  OptionsBuilder& mass_storage(Options::MatrixStorage s) { o_.mass_storage = s; return *this; }
  OptionsBuilder& user_solout(bool b) { o_.user_solout = b; return *this; }
  OptionsBuilder& nind1(size_t k) { o_.nind1 = k; return *this; }
  OptionsBuilder& nind2(size_t k) { o_.nind2 = k; return *this; }
  OptionsBuilder& nind3(size_t k) { o_.nind3 = k; return *this; }
  OptionsBuilder& event_config(std::vector<EventConfig> c) { o_.event_config = std::move(c); return *this; }
  OptionsBuilder& max_events(int n) { o_.max_events = n; return *this; }
  OptionsBuilder& max_out(int n) { o_.max_out = n; return *this; }
  OptionsBuilder& analytic_jac(bool b) { o_.analytic_jac = b; return *this; }
  OptionsBuilder& jac_sparsity(std::vector<std::pair<int, int>> nz) { o_.jac_sparsity = std::move(nz); return *this; }
  OptionsBuilder& strict_fp(bool b) { o_.strict_fp = b; return *this; }
  OptionsBuilder& fast_fp(bool b) { o_.fast_fp = b; return *this; }
  Options build() { return std::move(o_); }
 private:
  Options o_;
};
inline OptionsBuilder Options::builder() { return OptionsBuilder(); }

// ---- ContinuousOutput handle (reference src/solve/cont.rs:9-154) ---------------------------------
// The segments stay on the device; evaluation is ivpb_dense_eval.  Valid until the next dense_output solve on
// the same context (the context retains one log).
struct ContinuousOutput {
  std::shared_ptr<ivpb_ctx> ctx;
  int64_t index = 0;
  int n = 0;
  uint64_t generation = 0;     // ivpb_dense_generation at solve time: a later dense solve invalidates this handle (IVPB_ERR_CONFIG)
};

// ---- Solution (reference src/solve/solution.rs:7-97) -------------------------------------------
struct Solution {
  std::vector<Float> t;
  std::vector<std::vector<Float>> y;
  std::vector<std::vector<Float>> t_events;
  std::vector<std::vector<std::vector<Float>>> y_events;
  size_t nfev = 0, njev = 0, nlu = 0, nstep = 0, naccpt = 0, nrejct = 0;
  Status status = Status::Success;
  std::optional<ContinuousOutput> continuous_sol;
  // Solution::sol_span / sol / sol_many (solution.rs:25-72)
  std::optional<std::pair<Float, Float>> sol_span() const {
    if (!continuous_sol) return std::nullopt;
    double a = 0, b = 0; int32_t m = 0;
    if (ivpb_dense_span(continuous_sol->ctx.get(), continuous_sol->generation, continuous_sol->index, 1, &a, &b, &m) != IVPB_OK || m <= 0) return std::nullopt;
    return std::make_pair(a, b);
  }
  std::vector<std::vector<Float>> sol_many(const std::vector<Float>& ts) const {
    if (!continuous_sol) throw InterpolationError("dense output not enabled");
    const auto span = sol_span();
    if (!span) throw InterpolationError("dense output not enabled");
    const Float lo = std::min(span->first, span->second), hi = std::max(span->first, span->second);
    for (Float t : ts) if (t < lo || t > hi) throw InterpolationError("t outside the dense output span");
    const int n = continuous_sol->n;
    std::vector<int64_t> tr(ts.size(), continuous_sol->index);
    std::vector<Float> y(ts.size() * (size_t)n);
    std::vector<int32_t> ok(ts.size());
    if (ivpb_dense_eval(continuous_sol->ctx.get(), continuous_sol->generation, n, (int64_t)ts.size(), tr.data(), ts.data(), y.data(), ok.data()) != IVPB_OK)
      throw InterpolationError(ivpb_last_error(continuous_sol->ctx.get()));
    std::vector<std::vector<Float>> out(ts.size());
    for (size_t k = 0; k < ts.size(); ++k) {
      if (!ok[k]) throw InterpolationError("t outside the dense output span");
      out[k].assign(&y[k * n], &y[k * n] + n);
    }
    return out;
  }
  std::vector<Float> sol(Float t) const { return sol_many({t}).at(0); }
  // ContinuousOutput::evaluate_extrapolate (cont.rs:91-150): outside the stored steps the first / last one answers
  std::optional<std::vector<Float>> sol_extrapolate(Float t) const {
    if (!continuous_sol) return std::nullopt;
    const int n = continuous_sol->n;
    std::vector<Float> y((size_t)n);
    int32_t ok = 0;
    const int64_t tr = continuous_sol->index;
    if (ivpb_dense_eval_extrapolate(continuous_sol->ctx.get(), continuous_sol->generation, n, 1, &tr, &t, y.data(), &ok) != IVPB_OK || !ok) return std::nullopt;
    return y;
  }
  // batched-solve additions
  Float h_next = 0.0;          // IntegrationResult.h (src/methods/mod.rs:31-32)
  bool truncated = false;      // more samples / event hits than max_out / max_events allowed
  // Iterate over stored sample pairs (t_i, y_i): Solution::iter
  struct Iter {
    const Solution* s; size_t i;
    bool operator!=(const Iter& o) const { return i != o.i; }
    void operator++() { ++i; }
    std::pair<Float, const std::vector<Float>&> operator*() const { return {s->t[i], s->y[i]}; }
  };
  Iter begin() const { return {this, 0}; }
  Iter end() const { return {this, t.size()}; }
};

// ---- the IVP trait, device form (reference src/ivp.rs:27-121) -----------------------------------
class Context {
 public:
  explicit Context(const std::vector<int>& devices = {}) {
    ivpb_ctx* c = nullptr;
    const int rc = ivpb_create(&c, devices.empty() ? nullptr : devices.data(), (int)devices.size());
    if (rc != IVPB_OK) throw DeviceError(std::string("ivpb_create: ") + ivpb_last_error(nullptr));
    ctx_.reset(c, [](ivpb_ctx* p) { ivpb_destroy(p); });
  }
  ivpb_ctx* get() const { return ctx_.get(); }
  const std::shared_ptr<ivpb_ctx>& shared() const { return ctx_; }
  int device_count() const { return ivpb_device_count(ctx_.get()); }
 private:
  std::shared_ptr<ivpb_ctx> ctx_;
};

class Problem {
 public:
  // Built-in problems (ivpb_builtin in ivpb.h): "decay", "vdp_eps", "vdp_mu", "lorenz", "cr3bp",
  // "bouncing_ball", "robertson", "sho", "zero3", "exp2", "rational", "cannon", "linear100", "medakzo64",
  // "robertson_dae", "mass_linear3" (the last two carry a mass matrix: RADAU with mass_storage = Full).
  static Problem builtin(const std::string& name) {
    static const char* names[] = {"decay", "vdp_eps", "vdp_mu", "lorenz", "cr3bp", "bouncing_ball", "robertson",
                                  "sho", "zero3", "exp2", "rational", "cannon", "linear100", "medakzo64",
                                  "robertson_dae", "mass_linear3", "ball_bounce"};
    static_assert(sizeof(names) / sizeof(names[0]) == IVPB_P_BUILTIN_COUNT, "one name per ivpb_builtin id");
    for (int i = 0; i < IVPB_P_BUILTIN_COUNT; ++i)
      if (name == names[i]) {
        Problem p;
        p.builtin_ = i;
        if (ivpb_builtin_problem(nullptr, i, &p.n_, &p.p_, &p.n_events_) != IVPB_OK) throw ConfigError("unknown built-in problem");
        return p;
      }
    throw ConfigError("unknown built-in problem: " + name);
  }
  // `impl IVP for MyProblem` as CUDA C: the source must define
  //   __device__ void ivp_ode(double t, const double* y, const double* p, double* dydt);
  // and, if n_events > 0 / has_jac, ivp_events(t, y, p, g) / ivp_jac(t, y, p, J row-major).
  // `has_mass`: the source also defines `ivp_mass(const double* p, double* M)` (IVP::mass, row-major n x n).
  static Problem from_cuda_source(std::string src, int n, int p = 0, int n_events = 0, bool has_jac = false,
                                  bool has_mass = false, bool has_solout = false) {
    Problem q;
    q.src_ = std::move(src); q.n_ = n; q.p_ = p; q.n_events_ = n_events;
    q.has_jac_ = (has_jac ? 1 : 0) | (has_mass ? 2 : 0) | (has_solout ? 4 : 0);     // `ivp_solout`: see include/ivpb.h
    return q;
  }
  int n() const { return n_; }
  int n_params() const { return p_; }
  int n_events() const { return n_events_; }
  int handle(const Context& ctx) const {      // user problems are registered per context, lazily
    if (builtin_ >= 0) return builtin_;
    for (auto& h : handles_) if (h.first == ctx.get()) return h.second;
    int h = -1;
    if (ivpb_nvrtc_problem(ctx.get(), src_.c_str(), n_, p_, n_events_, has_jac_, &h) != IVPB_OK)
      throw ConfigError(ivpb_last_error(ctx.get()));
    handles_.push_back({ctx.get(), h});
    return h;
  }
 private:
  int builtin_ = -1, n_ = 0, p_ = 0, n_events_ = 0;
  int has_jac_ = 0;            // bit 0: ivp_jac, bit 1: ivp_mass
  std::string src_;
  mutable std::vector<std::pair<ivpb_ctx*, int>> handles_;
};

inline Context& default_context() {
  static Context ctx;
  return ctx;
}

// ---- solve_ivp_batch -----------------------------------------------------------------------------
// N trajectories of one problem: Y0 is [N x n] row-major, params [N x p] row-major (nullptr if p == 0).
// Returns one Solution per trajectory with the reference's semantics (t / y hold the t_eval samples, or the
// accepted step end points when t_eval is None; a terminal event appends its point; numerical failures are
// reported in Solution::status, configuration errors throw ConfigError before any stepping).
inline std::vector<Solution> solve_ivp_batch(const Problem& f, Float t0, Float tf, const Float* Y0, const Float* params,
                                            size_t N, const Options& options, Context* ctx_in = nullptr) {
  Context& ctx = ctx_in ? *ctx_in : default_context();
  const int n = f.n(), ne = f.n_events();
  if (options.rtol.is_vector() && (int)options.rtol.len() != n)      // Tolerance::Vector length mismatch panics
    throw ConfigError("rtol vector length does not match the state size");
  if (options.atol.is_vector() && (int)options.atol.len() != n)
    throw ConfigError("atol vector length does not match the state size");
  ivpb_options o{};
  o.method = (int32_t)options.method;
  o.n_rtol = options.rtol.is_vector() ? n : 1; o.rtol = options.rtol.data();
  o.n_atol = options.atol.is_vector() ? n : 1; o.atol = options.atol.data();
  o.has_first_step = options.first_step.has_value(); o.first_step = options.first_step.value_or(0.0);
  o.has_max_step = options.max_step.has_value(); o.max_step = options.max_step.value_or(0.0);
  o.has_min_step = options.min_step.has_value(); o.min_step = options.min_step.value_or(0.0);
  o.has_max_steps = options.max_steps.has_value(); o.max_steps = options.max_steps.value_or(0);
  o.has_t_eval = options.t_eval.has_value();
  o.n_t_eval = o.has_t_eval ? (int32_t)options.t_eval->size() : 0;
  o.t_eval = o.has_t_eval ? options.t_eval->data() : nullptr;
  o.dense_output = options.dense_output;
  o.max_segments = options.dense_output ? options.max_segments : 0;
  o.mass_storage = options.mass_storage == Options::MatrixStorage::Full ? 1 : 0;
  o.user_solout = options.user_solout ? 1 : 0;
  o.nind1 = options.nind1 ? (int32_t)*options.nind1 : -1;
  o.nind2 = options.nind2 ? (int32_t)*options.nind2 : -1;
  o.nind3 = options.nind3 ? (int32_t)*options.nind3 : -1;
  std::vector<int32_t> dirs;
  std::vector<int64_t> terms;
  if (options.event_config) {
    if ((int)options.event_config->size() != ne) throw ConfigError("event_config length must equal the problem's n_events");
    for (auto& c : *options.event_config) {
      dirs.push_back((int32_t)c.direction);
      terms.push_back(c.terminal_count_ ? (int64_t)*c.terminal_count_ : -1);
    }
    o.n_event_cfg = ne; o.ev_direction = dirs.data(); o.ev_terminal_count = terms.data();
  }
  o.max_events = ne > 0 ? options.max_events : 0;
  o.max_out = options.max_out;
  o.jac_mode = options.analytic_jac ? 1 : 0;
  std::vector<int32_t> sp_colptr, sp_rows;
  if (options.jac_sparsity) {        // (row, col) pairs -> compressed columns, rows ascending, duplicates dropped
    std::vector<std::vector<int32_t>> cols((size_t)n);
    for (auto& rc : *options.jac_sparsity) {
      if (rc.first < 0 || rc.first >= n || rc.second < 0 || rc.second >= n) throw ConfigError("jac_sparsity entry out of range");
      cols[(size_t)rc.second].push_back(rc.first);
    }
    sp_colptr.push_back(0);
    for (auto& c : cols) {
      std::sort(c.begin(), c.end());
      c.erase(std::unique(c.begin(), c.end()), c.end());
      sp_rows.insert(sp_rows.end(), c.begin(), c.end());
      sp_colptr.push_back((int32_t)sp_rows.size());
    }
    if (sp_rows.empty()) sp_rows.push_back(0);
    o.has_jac_sparsity = 1; o.jac_sparsity_colptr = sp_colptr.data(); o.jac_sparsity_rows = sp_rows.data();
  }
  o.flags = (options.strict_fp ? IVPB_FLAG_STRICT_FP : 0u) | (options.fast_fp ? IVPB_FLAG_FAST_FP : 0u);

  const size_t cap = o.has_t_eval ? (size_t)o.n_t_eval + 1 : (size_t)o.max_out;
  std::vector<int32_t> status(N), n_out(N), ev_count(N * (size_t)ne);
  std::vector<uint32_t> counters(N * 6);
  std::vector<Float> t_final(N), y_final(N * (size_t)n), h_next(N), t_out(N * cap), y_out(N * cap * n),
      ev_t(N * (size_t)ne * o.max_events), ev_y(N * (size_t)ne * o.max_events * n);
  ivpb_outputs out{};
  out.status = status.data(); out.counters = counters.data(); out.t_final = t_final.data();
  out.y_final = y_final.data(); out.h_next = h_next.data();
  out.n_out = n_out.data(); out.t_out = cap ? t_out.data() : nullptr; out.y_out = cap ? y_out.data() : nullptr;
  if (ne > 0) { out.ev_count = ev_count.data(); out.ev_t = ev_t.data(); out.ev_y = ev_y.data(); }
  const int rc = ivpb_solve_batch(ctx.get(), f.handle(ctx), &o, (int64_t)N, t0, tf, Y0, params, &out);
  if (rc == IVPB_ERR_CONFIG) throw ConfigError(ivpb_last_error(ctx.get()));
  if (rc != IVPB_OK) throw DeviceError(ivpb_last_error(ctx.get()));

  std::vector<Solution> sols(N);
  for (size_t i = 0; i < N; ++i) {
    Solution& s = sols[i];
    s.status = (Status)status[i];
    const uint32_t* c = &counters[6 * i];
    s.nfev = c[0]; s.njev = c[1]; s.nlu = c[2]; s.nstep = c[3]; s.naccpt = c[4]; s.nrejct = c[5];
    s.h_next = h_next[i];
    if (options.dense_output) s.continuous_sol = ContinuousOutput{ctx.shared(), (int64_t)i, n, ivpb_dense_generation(ctx.get())};
    const size_t m = std::min<size_t>((size_t)std::max(n_out[i], 0), cap);
    s.truncated = (size_t)std::max(n_out[i], 0) > cap;
    s.t.assign(&t_out[i * cap], &t_out[i * cap] + m);
    s.y.resize(m);
    for (size_t k = 0; k < m; ++k) s.y[k].assign(&y_out[(i * cap + k) * n], &y_out[(i * cap + k) * n] + n);
    if (cap == 0) {     // samples not stored: keep the reference's "last point" reachable
      s.t = {t_final[i]};
      s.y = {std::vector<Float>(&y_final[i * n], &y_final[i * n] + n)};
    }
    s.t_events.resize(ne); s.y_events.resize(ne);
    for (int e = 0; e < ne; ++e) {
      const size_t hits = (size_t)ev_count[i * ne + e], keep = std::min<size_t>(hits, (size_t)o.max_events);
      s.truncated = s.truncated || hits > keep;
      const size_t base = (i * ne + e) * (size_t)o.max_events;
      s.t_events[e].assign(&ev_t[base], &ev_t[base] + keep);
      s.y_events[e].resize(keep);
      for (size_t k = 0; k < keep; ++k) s.y_events[e][k].assign(&ev_y[(base + k) * n], &ev_y[(base + k) * n] + n);
    }
  }
  return sols;
}

inline std::vector<Solution> solve_ivp_batch(const Problem& f, Float t0, Float tf, const std::vector<Float>& Y0,
                                            const std::vector<Float>& params, const Options& options,
                                            Context* ctx = nullptr) {
  if (f.n() <= 0 || Y0.size() % (size_t)f.n() != 0) throw ConfigError("Y0 must hold N x n values");
  const size_t N = Y0.size() / (size_t)f.n();
  if (params.size() != N * (size_t)f.n_params()) throw ConfigError("params must hold N x p values");
  return solve_ivp_batch(f, t0, tf, Y0.data(), f.n_params() ? params.data() : nullptr, N, options, ctx);
}

// solve_ivp(f, x0, xend, y0, options) -- the reference's signature (src/solve/solve_ivp.rs:99-108), one
// trajectory, parameters (the fields of the reference's problem struct) passed explicitly.
inline Solution solve_ivp(const Problem& f, Float x0, Float xend, const std::vector<Float>& y0, const Options& options,
                          const std::vector<Float>& params = {}, Context* ctx = nullptr) {
  if ((int)y0.size() != f.n()) throw ConfigError("y0 length does not match the problem's state size");
  return solve_ivp_batch(f, x0, xend, y0, params, options, ctx).at(0);
}

}  // namespace ivp
