// capi.cpp -- C entry points of the CPU oracle (liboracle.so), driven from Python via ctypes.
// TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Output layout is the one of include/ivpb.h so arrays compare directly.
//
// The threaded batch driver is the stand-in for "the reference CPU path run across all host cores":
// the reference has no batch API and no rayon dependency, so the outer loop over
// trajectories is a std::thread pool pulling chunks from an atomic counter; each trajectory runs
// oracle::solve_ivp, the restatement of reference src/solve/solve_ivp.rs:99-313.
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>

#include "../include/ivpb.h"
#include "test_problems.hpp"

using namespace oracle;

namespace {

thread_local std::string g_err;

Options make_options(const ivpb_options* o, int n) {
  Options O;
  O.method = (Method)o->method;
  if (o->rtol) O.rtol = (o->n_rtol == 1) ? Tol(o->rtol[0]) : Tol::vec(std::vector<double>(o->rtol, o->rtol + o->n_rtol));
  if (o->atol) O.atol = (o->n_atol == 1) ? Tol(o->atol[0]) : Tol::vec(std::vector<double>(o->atol, o->atol + o->n_atol));
  (void)n;
  O.has_max_steps = o->has_max_steps; O.max_steps = (size_t)o->max_steps;
  O.has_t_eval = o->has_t_eval;
  if (o->has_t_eval && o->n_t_eval > 0) O.t_eval.assign(o->t_eval, o->t_eval + o->n_t_eval);
  O.has_first_step = o->has_first_step; O.first_step = o->first_step;
  O.has_max_step = o->has_max_step; O.max_step = o->max_step;
  O.has_min_step = o->has_min_step; O.min_step = o->min_step;
  O.dense_output = o->dense_output != 0;
  O.user_solout = o->user_solout != 0;
  O.mass_full = o->mass_storage == 1;
  O.nind1 = o->nind1; O.nind2 = o->nind2; O.nind3 = o->nind3;
  return O;
}

template <class P>
void scatter(const Solution& S, const P&, const ivpb_options* o, const ivpb_outputs* out, int64_t i,
             double x0, const double* y0) {
  constexpr int n = P::N, ne = P::NEV;
  if (out->status) out->status[i] = (int32_t)S.status;
  if (out->counters) {
    uint32_t* c = out->counters + 6 * i;
    c[0] = (uint32_t)S.nfev; c[1] = (uint32_t)S.njev; c[2] = (uint32_t)S.nlu;
    c[3] = (uint32_t)S.nstep; c[4] = (uint32_t)S.naccpt; c[5] = (uint32_t)S.nrejct;
  }
  const bool term = S.status == Status::UserInterrupt && !S.t.empty();
  if (out->t_final) out->t_final[i] = term ? S.t.back() : (S.y_last.empty() ? x0 : S.x_last);
  if (out->y_final) {
    const double* src = term ? &S.y[S.y.size() - n] : (S.y_last.empty() ? y0 : S.y_last.data());
    std::memcpy(out->y_final + (size_t)n * i, src, sizeof(double) * n);
  }
  if (out->h_next) out->h_next[i] = S.h_next;
  const int64_t cap = o->has_t_eval ? (int64_t)o->n_t_eval + 1 : (int64_t)o->max_out;
  if (out->n_out) out->n_out[i] = cap > 0 ? (int32_t)S.t.size() : 0;   // only meaningful when samples are stored
  const int64_t m = std::min<int64_t>((int64_t)S.t.size(), cap);
  if (out->t_out && m > 0) std::memcpy(out->t_out + cap * i, S.t.data(), sizeof(double) * m);
  if (out->y_out && m > 0) std::memcpy(out->y_out + cap * n * i, S.y.data(), sizeof(double) * m * n);
  if (o->dense_output && o->max_segments > 0) {   // ContinuousOutput segments in the layout of include/ivpb.h
    const int64_t sc = o->max_segments, nc = (int64_t)coeffs_per_state((Method)o->method) * n;
    if (out->n_seg) out->n_seg[i] = (int32_t)S.segs.size();
    const int64_t ms = std::min<int64_t>((int64_t)S.segs.size(), sc);
    for (int64_t k = 0; k < ms; ++k) {
      if (out->seg_x) { out->seg_x[(sc * i + k) * 2] = S.segs[k].xold; out->seg_x[(sc * i + k) * 2 + 1] = S.segs[k].h; }
      if (out->seg_cont) std::memcpy(out->seg_cont + (sc * i + k) * nc, S.segs[k].cont.data(), sizeof(double) * nc);
    }
  }
  for (int e = 0; e < ne; ++e) {
    const int64_t k = (int64_t)S.t_events[e].size();
    if (out->ev_count) out->ev_count[(int64_t)ne * i + e] = (int32_t)k;
    const int64_t mk = std::min<int64_t>(k, o->max_events);
    if (out->ev_t && mk > 0)
      std::memcpy(out->ev_t + ((int64_t)ne * i + e) * o->max_events, S.t_events[e].data(), sizeof(double) * mk);
    if (out->ev_y && mk > 0)
      std::memcpy(out->ev_y + ((int64_t)ne * i + e) * o->max_events * n, S.y_events[e].data(), sizeof(double) * mk * n);
  }
}

template <class P>
int run_batch(const ivpb_options* o, int64_t N, double t0, double tf, const double* y0, const double* params,
              const ivpb_outputs* out, int nthreads) {
  constexpr int n = P::N, p = P::P, ne = P::NEV;
  const Options O = make_options(o, n);
  std::vector<EventConfig> evc;
  if (o->n_event_cfg > 0) {
    if (o->n_event_cfg != ne) { g_err = "n_event_cfg must equal the problem's n_events"; return 1; }
    for (int e = 0; e < ne; ++e) {
      EventConfig c;
      c.direction = o->ev_direction[e] > 0 ? Direction::Positive : (o->ev_direction[e] < 0 ? Direction::Negative : Direction::All);
      c.terminal_count = (long)o->ev_terminal_count[e];
      evc.push_back(c);
    }
  }
  std::unique_ptr<Sparsity> sparsity;
  if (o->has_jac_sparsity && o->jac_sparsity_colptr) sparsity.reset(new Sparsity((size_t)n, o->jac_sparsity_colptr, o->jac_sparsity_rows));
  if (nthreads < 1) nthreads = 1;
  std::atomic<int64_t> next{0};
  std::atomic<int> failed{0};
  std::string err;
  const int64_t chunk = 64;
  auto worker = [&]() {
    for (;;) {
      int64_t b = next.fetch_add(chunk);
      if (b >= N) break;
      int64_t e = std::min(N, b + chunk);
      for (int64_t i = b; i < e; ++i) {
        P prob;
        prob.p = (p > 0 && params) ? params + (int64_t)p * i : nullptr;
        prob.ev_cfg = evc.empty() ? nullptr : evc.data();
        prob.jac_mode = o->jac_mode;
        prob.sparsity = sparsity.get();
        std::vector<double> yy(y0 + (int64_t)n * i, y0 + (int64_t)n * (i + 1));
        try {
          Solution S = solve_ivp(prob, t0, tf, yy, O);
          scatter<P>(S, prob, o, out, i, t0, yy.data());
        } catch (const std::exception& ex) {
          if (!failed.exchange(1)) err = ex.what();
          return;
        }
      }
    }
  };
  if (nthreads == 1) worker();
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  if (failed.load()) { g_err = err; return 2; }
  return 0;
}

template <class P>
int dense_eval(const ivpb_options* o, double t0, double tf, const double* y0, const double* params,
               const double* ts, int nts, double* ys, int32_t* ok, double* span, int extrapolate = 0) {
  constexpr int n = P::N;
  Options O = make_options(o, n);
  O.dense_output = true;
  P prob; prob.p = params; prob.jac_mode = o->jac_mode;
  std::vector<double> yy(y0, y0 + n);
  try {
    Solution S = solve_ivp(prob, t0, tf, yy, O);
    double a = 0, b = 0;
    bool has = sol_span(S, a, b);
    if (span) { span[0] = a; span[1] = b; span[2] = has ? 1.0 : 0.0; }
    for (int k = 0; k < nts; ++k) ok[k] = (extrapolate ? sol_eval_extrapolate(S, ts[k], ys + (size_t)n * k) : sol_eval(S, ts[k], ys + (size_t)n * k)) ? 1 : 0;
  } catch (const std::exception& ex) { g_err = ex.what(); return 2; }
  return 0;
}

}  // namespace

extern "C" {

// Column groups of a sparsity structure (src/python/sparsity.rs:109-154), for tests of the runtime's own grouping.
int oracle_group_columns(int n, const int32_t* colptr, const int32_t* rows, int32_t* groups) {
  Sparsity sp((size_t)n, colptr, rows);
  for (int c = 0; c < n; ++c) groups[c] = (int32_t)sp.groups[c];
  return (int)sp.n_groups;
}

int oracle_problem_dims(int problem, int* n, int* p, int* ne) {
#define DIMS(T) { *n = T::N; *p = T::P; *ne = T::NEV; return 0; }
  switch (problem) {
    case IVPB_P_DECAY: DIMS(Decay) case IVPB_P_VDP_EPS: DIMS(VdpEps) case IVPB_P_VDP_MU: DIMS(VdpMu)
    case IVPB_P_LORENZ: DIMS(Lorenz) case IVPB_P_CR3BP: DIMS(Cr3bp) case IVPB_P_BALL: DIMS(Ball)
    case IVPB_P_ROBERTSON: DIMS(Robertson) case IVPB_P_SHO: DIMS(Sho) case IVPB_P_ZERO3: DIMS(Zero3)
    case IVPB_P_EXP2: DIMS(Exp2) case IVPB_P_RATIONAL: DIMS(Rational) case IVPB_P_CANNON: DIMS(Cannon)
    case IVPB_P_LINEAR100: DIMS(Linear100) case IVPB_P_MEDAKZO64: DIMS(Medakzo64)
    case IVPB_P_ROBERTSON_DAE: DIMS(RobertsonDae) case IVPB_P_MASS_LINEAR3: DIMS(MassLinear3) case IVPB_P_BALL_BOUNCE: DIMS(BallBounce)
    case 100: DIMS(RationalEv) case 101: DIMS(Sys3) case 102: DIMS(Scale1) case 103: DIMS(Scale2)
    case 104: DIMS(Radial) case 105: DIMS(ConstRates) case 106: DIMS(Linear2) case 107: DIMS(MassLinear4) case 108: DIMS(Medakzo400) case 109: DIMS(MassLinear12) case 110: DIMS(DecayKick40)
    default: return 1;
  }
}

int oracle_solve_batch(int problem, const ivpb_options* o, int64_t N, double t0, double tf, const double* y0,
                       const double* params, const ivpb_outputs* out, int nthreads) {
#define RB(T) run_batch<T>(o, N, t0, tf, y0, params, out, nthreads)
  switch (problem) {
    case IVPB_P_DECAY: return RB(Decay); case IVPB_P_VDP_EPS: return RB(VdpEps); case IVPB_P_VDP_MU: return RB(VdpMu);
    case IVPB_P_LORENZ: return RB(Lorenz); case IVPB_P_CR3BP: return RB(Cr3bp); case IVPB_P_BALL: return RB(Ball);
    case IVPB_P_ROBERTSON: return RB(Robertson); case IVPB_P_SHO: return RB(Sho); case IVPB_P_ZERO3: return RB(Zero3);
    case IVPB_P_EXP2: return RB(Exp2); case IVPB_P_RATIONAL: return RB(Rational); case IVPB_P_CANNON: return RB(Cannon);
    case IVPB_P_LINEAR100: return RB(Linear100); case IVPB_P_MEDAKZO64: return RB(Medakzo64);
    case IVPB_P_ROBERTSON_DAE: return RB(RobertsonDae); case IVPB_P_MASS_LINEAR3: return RB(MassLinear3);
    case IVPB_P_BALL_BOUNCE: return RB(BallBounce);
    case 100: return RB(RationalEv); case 101: return RB(Sys3); case 102: return RB(Scale1); case 103: return RB(Scale2);
    case 104: return RB(Radial); case 105: return RB(ConstRates); case 106: return RB(Linear2); case 107: return RB(MassLinear4); case 108: return RB(Medakzo400); case 109: return RB(MassLinear12); case 110: return RB(DecayKick40);
    default: g_err = "unknown problem id"; return 1;
  }
}

// Solve ONE trajectory with dense_output=true and evaluate Solution::sol at ts (reference
// src/solve/solution.rs:25-44).  span = {t_start, t_end, has_span}.
int oracle_dense_eval(int problem, const ivpb_options* o, double t0, double tf, const double* y0,
                      const double* params, const double* ts, int nts, double* ys, int32_t* ok, double* span) {
#define DE(T) dense_eval<T>(o, t0, tf, y0, params, ts, nts, ys, ok, span)
  switch (problem) {
    case IVPB_P_DECAY: return DE(Decay); case IVPB_P_VDP_EPS: return DE(VdpEps); case IVPB_P_VDP_MU: return DE(VdpMu);
    case IVPB_P_LORENZ: return DE(Lorenz); case IVPB_P_CR3BP: return DE(Cr3bp); case IVPB_P_BALL: return DE(Ball);
    case IVPB_P_ROBERTSON: return DE(Robertson); case IVPB_P_SHO: return DE(Sho); case IVPB_P_ZERO3: return DE(Zero3);
    case IVPB_P_EXP2: return DE(Exp2); case IVPB_P_RATIONAL: return DE(Rational); case IVPB_P_CANNON: return DE(Cannon);
    case IVPB_P_LINEAR100: return DE(Linear100); case IVPB_P_MEDAKZO64: return DE(Medakzo64);
    case IVPB_P_ROBERTSON_DAE: return DE(RobertsonDae); case IVPB_P_MASS_LINEAR3: return DE(MassLinear3);
    case IVPB_P_BALL_BOUNCE: return DE(BallBounce);
    case 100: return DE(RationalEv); case 101: return DE(Sys3); case 102: return DE(Scale1); case 103: return DE(Scale2);
    case 104: return DE(Radial); case 105: return DE(ConstRates); case 106: return DE(Linear2); case 107: return DE(MassLinear4); case 108: return DE(Medakzo400); case 109: return DE(MassLinear12); case 110: return DE(DecayKick40);
    default: g_err = "unknown problem id"; return 1;
  }
}

// The same with ContinuousOutput::evaluate_extrapolate (src/solve/cont.rs:91-150).
int oracle_dense_eval_extrapolate(int problem, const ivpb_options* o, double t0, double tf, const double* y0,
                      const double* params, const double* ts, int nts, double* ys, int32_t* ok, double* span) {
#define DX(T) dense_eval<T>(o, t0, tf, y0, params, ts, nts, ys, ok, span, 1)
  switch (problem) {
    case IVPB_P_DECAY: return DX(Decay); case IVPB_P_VDP_EPS: return DX(VdpEps); case IVPB_P_VDP_MU: return DX(VdpMu);
    case IVPB_P_LORENZ: return DX(Lorenz); case IVPB_P_CR3BP: return DX(Cr3bp); case IVPB_P_BALL: return DX(Ball);
    case IVPB_P_ROBERTSON: return DX(Robertson); case IVPB_P_SHO: return DX(Sho); case IVPB_P_ZERO3: return DX(Zero3);
    case IVPB_P_EXP2: return DX(Exp2); case IVPB_P_RATIONAL: return DX(Rational); case IVPB_P_CANNON: return DX(Cannon);
    case IVPB_P_LINEAR100: return DX(Linear100); case IVPB_P_MEDAKZO64: return DX(Medakzo64);
    case IVPB_P_ROBERTSON_DAE: return DX(RobertsonDae); case IVPB_P_MASS_LINEAR3: return DX(MassLinear3);
    case IVPB_P_BALL_BOUNCE: return DX(BallBounce);
    case 100: return DX(RationalEv); case 101: return DX(Sys3); case 102: return DX(Scale1); case 103: return DX(Scale2);
    case 104: return DX(Radial); case 105: return DX(ConstRates); case 106: return DX(Linear2); case 107: return DX(MassLinear4); case 108: return DX(Medakzo400); case 109: return DX(MassLinear12); case 110: return DX(DecayKick40);
    default: g_err = "unknown problem id"; return 1;
  }
}

const char* oracle_last_error(void) { return g_err.c_str(); }
int oracle_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
