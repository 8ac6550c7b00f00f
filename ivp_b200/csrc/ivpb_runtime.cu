// ivpb_runtime.cu -- host runtime behind the C ABI of include/ivpb.h.
//
// Owns the device set (one stream + work-queue counter + grow-only staging buffers per device),
// validates options the way the reference's solvers do before stepping (Error::Config), picks the
// kernel specialisation, and launches the persistent kernels.  No numerics live here, and there is no
// CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ivpb.h"
#include "ivpb_common.cuh"
#include "ivpb_runtime.h"

using ivpb::KArgs;
using ivpb::u64;

// ------------------------------------------------------------------------------------------------
// Built-in kernel tables (one lookup function per problem and floating-point mode; ivpb_inst.cu)
typedef const void* (*lookup_fn)(int method, int feat, ivpb_pinfo* info);
typedef const void* (*lookup_impl_fn)(int method, int feat, int* block, int* smem, int* units);
#define DECL(tag)                                                                  \
  extern "C" const void* ivpb_lookup_##tag(int, int, ivpb_pinfo*);                 \
  extern "C" const void* ivpb_lookup_strict_##tag(int, int, ivpb_pinfo*);          \
  extern "C" const void* ivpb_lookup_strictd_##tag(int, int, ivpb_pinfo*);         \
  extern "C" const void* ivpb_lookup_impl_strictd_##tag(int, int, int*, int*, int*); \
  extern "C" const void* ivpb_lookup_impl_##tag(int, int, int*, int*, int*);       \
  extern "C" const void* ivpb_lookup_impl_strict_##tag(int, int, int*, int*, int*);
DECL(decay) DECL(vdp_eps) DECL(vdp_mu) DECL(lorenz) DECL(cr3bp) DECL(ball) DECL(robertson) DECL(sho)
DECL(zero3) DECL(exp2) DECL(rational) DECL(cannon) DECL(linear100) DECL(medakzo64) DECL(robertson_dae) DECL(mass_linear3) DECL(ball_bounce)
#undef DECL
// [0] FMA build, [1] strict build, [2] strict build with deferred guards (first pass of a strict solve, ivpb_exact.cuh)
#define ROW(tag) {ivpb_lookup_##tag, ivpb_lookup_strict_##tag, ivpb_lookup_strictd_##tag}
static const lookup_fn BUILTIN[IVPB_P_BUILTIN_COUNT][3] = {
    ROW(decay), ROW(vdp_eps), ROW(vdp_mu), ROW(lorenz), ROW(cr3bp), ROW(ball),
    ROW(robertson), ROW(sho), ROW(zero3), ROW(exp2), ROW(rational), ROW(cannon), ROW(linear100), ROW(medakzo64),
    ROW(robertson_dae), ROW(mass_linear3), ROW(ball_bounce)};
#undef ROW
// dense-output evaluation kernels (ivpb_dense.cu)
extern "C" cudaError_t ivpb_launch_dense_eval(int method, int n, int n_cont, int cap, const int* seg_n, const double* seg_x,
                                              const double* seg_cont, long long M, const long long* traj, long long lo,
                                              long long Ng, const double* ts, double* y, int* ok, int extrapolate,
                                              cudaStream_t stream);
extern "C" cudaError_t ivpb_launch_dense_span(int cap, const int* seg_n, const double* seg_x, long long first, long long count,
                                              double* t_start, double* t_end, int* n_out, cudaStream_t stream);
// RADAU / BDF kernels (ivpb_inst_implicit.cu)
#define ROW(tag) {ivpb_lookup_impl_##tag, ivpb_lookup_impl_strict_##tag, ivpb_lookup_impl_strictd_##tag}
static const lookup_impl_fn BUILTIN_IMPL[IVPB_P_BUILTIN_COUNT][3] = {
    ROW(decay), ROW(vdp_eps), ROW(vdp_mu), ROW(lorenz), ROW(cr3bp), ROW(ball),
    ROW(robertson), ROW(sho), ROW(zero3), ROW(exp2), ROW(rational), ROW(cannon), ROW(linear100), ROW(medakzo64),
    ROW(robertson_dae), ROW(mass_linear3), ROW(ball_bounce)};
#undef ROW

namespace {

std::string g_create_err;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int MAX_CHUNKS = 64;
enum { OUT_STATUS, OUT_COUNTERS, OUT_TFINAL, OUT_YFINAL, OUT_HNEXT, OUT_NOUT, OUT_TOUT, OUT_YOUT, OUT_EVCOUNT,
       OUT_EVT, OUT_EVY, OUT_NSEG, OUT_SEGX, OUT_SEGC, OUT_FIELDS };

struct Device {
  int id = 0;
  int sms = 0;
  cudaStream_t stream = nullptr;
  int sp_ngroups = 0;               // column groups of the staged jac_sparsity (stage_shared)
  cudaStream_t stream2 = nullptr;   // second lane of the chunked host-buffer pipeline (ivpb_solve_batch)
  cudaEvent_t ev_shared = nullptr;  // "t_eval / tolerance staging uploaded" (recorded on stream, awaited by stream2)
  u64* queue = nullptr;             // work-queue head(s) of the launches in flight on this device
  unsigned* chunk_count = nullptr;  // completion counters of the host-buffer pipeline (device, MAX_CHUNKS)
  int* chunk_flag = nullptr;        // completion flags (page-locked host memory, mapped; MAX_CHUNKS)
  int* chunk_flag_dev = nullptr;    // the same flags as the device sees them
  int* in_flag = nullptr;           // arrival flags of the input chunks (device, MAX_CHUNKS; set by H2D copies of `ones`)
  int* ones = nullptr;              // page-locked host array of MAX_CHUNKS ones, the source of those copies
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;   // cross-device ordering for the peer-copy path
  cudaEvent_t ev_last = nullptr;    // end of the most recent solve enqueued on this device: the per-device queue counter,
                                    // t_eval / tolerance staging and sort buffers are shared, so solves on one context are
                                    // serialised on the device even when the caller hands in different streams
  Buf y0, params, t_eval, tol_ext, sparsity, scratch, rerun, out[OUT_FIELDS];
  Buf sort_keys, sort_vals, sort_tmp, sort_minmax;   // locality order of the shard (locality_order)
  Buf q_traj, q_ts, q_y, q_ok;      // ivpb_dense_eval query staging (grow-only)
  Buf dense_nseg, dense_segx, dense_segc;   // the retained dense log of this device's shard (never shared with dev.out[])
};

// bytes per trajectory of every output field (include/ivpb.h `ivpb_outputs`)
int coeffs_per_state(int method) {   // Method::coeffs_per_state, src/solve/options.rs:34-43
  return method == IVPB_DOPRI5 ? 5 : method == IVPB_DOP853 ? 8 : method == IVPB_BDF ? 7 : 4;
}
void field_bytes(int n, int nev, int64_t cap, int max_events, int64_t seg_cap, int n_cont, size_t per[OUT_FIELDS]) {
  const size_t v[OUT_FIELDS] = {4, 24, 8, 8u * n, 8, 4, 8u * (size_t)cap, 8u * (size_t)cap * n,
                                4u * nev, 8u * (size_t)nev * max_events, 8u * (size_t)nev * max_events * n,
                                4, 16u * (size_t)seg_cap, 8u * (size_t)seg_cap * n_cont};
  for (int f = 0; f < OUT_FIELDS; ++f) per[f] = v[f];
}
void out_to_array(const ivpb_outputs* o, void* a[OUT_FIELDS]) {
  void* v[OUT_FIELDS] = {o->status, o->counters, o->t_final, o->y_final, o->h_next, o->n_out,
                         o->t_out, o->y_out, o->ev_count, o->ev_t, o->ev_y, o->n_seg, o->seg_x, o->seg_cont};
  for (int f = 0; f < OUT_FIELDS; ++f) a[f] = v[f];
}
void array_to_out(void* const a[OUT_FIELDS], ivpb_outputs* d) {
  d->status = (int32_t*)a[OUT_STATUS]; d->counters = (uint32_t*)a[OUT_COUNTERS];
  d->t_final = (double*)a[OUT_TFINAL]; d->y_final = (double*)a[OUT_YFINAL]; d->h_next = (double*)a[OUT_HNEXT];
  d->n_out = (int32_t*)a[OUT_NOUT]; d->t_out = (double*)a[OUT_TOUT]; d->y_out = (double*)a[OUT_YOUT];
  d->ev_count = (int32_t*)a[OUT_EVCOUNT]; d->ev_t = (double*)a[OUT_EVT]; d->ev_y = (double*)a[OUT_EVY];
  d->n_seg = (int32_t*)a[OUT_NSEG]; d->seg_x = (double*)a[OUT_SEGX]; d->seg_cont = (double*)a[OUT_SEGC];
}

}  // namespace

// The dense output retained by the last host-buffer solve with dense_output = 1 (ivpb_dense_eval).
struct DenseLog {
  bool valid = false;
  uint64_t generation = 0;            // identity of the retained log (ivpb_dense_generation)
  int method = 0, n = 0, n_cont = 0, cap = 0;
  int64_t N = 0;
  std::vector<int64_t> lo, count;     // shard of every device
};

struct ivpb_ctx {
  std::vector<Device> devs;
  std::string err;
  uint64_t launches = 0;
  DenseLog dense;
  std::vector<ivpb_user_problem> user;   // NVRTC problems (ivpb_nvrtc.cpp)
  // floating-point mode of the explicit methods (resolve_fp): decisions of the parity pilot, keyed by the solve's
  // configuration, and what the most recent solve ran (ivpb_last_fp_mode)
  struct FpChoice { uint64_t key; int strict; int32_t info[5]; };
  std::vector<FpChoice> fp_cache;
  int fp_override = -1;                  // set around launch_shard by the entry points: 0 FMA build, 1 strict build
  int32_t last_fp[6] = {0, 0, 0, 0, 0, 0};   // {strict, source, sample, status mismatches, step-count mismatches, out of tolerance}
  Buf pilot_in, pilot_out[2], pilot_res;
};

namespace {

}  // namespace
void ivpb_set_error(ivpb_ctx* ctx, const std::string& msg) { if (ctx) ctx->err = msg; }
namespace {
int fail(ivpb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return fail(ctx, IVPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
  } while (0)

struct ProblemInfo {
  int n = 0, p = 0, nev = 0, has_jac = 0;
  bool user = false; int uidx = -1;
  int ev_dir[8] = {0}; long long ev_term[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
};

int problem_info(ivpb_ctx* ctx, int problem, ProblemInfo* pi) {
  if (problem >= 0 && problem < IVPB_P_BUILTIN_COUNT) {
    ivpb_pinfo info;
    BUILTIN[problem][0](-1, 0, &info);
    pi->n = info.n; pi->p = info.p; pi->nev = info.nev; pi->has_jac = info.has_jac;
    for (int e = 0; e < 8; ++e) { pi->ev_dir[e] = info.ev_dir[e]; pi->ev_term[e] = info.ev_term[e]; }
    return 0;
  }
  int u = problem - IVPB_USER_HANDLE_BASE;
  if (u >= 0 && u < (int)ctx->user.size()) {
    const ivpb_user_problem& up = ctx->user[u];
    pi->n = up.n; pi->p = up.p; pi->nev = up.n_events; pi->has_jac = up.has_jac; pi->user = true; pi->uidx = u;
    return 0;
  }
  return fail(ctx, IVPB_ERR_CONFIG, "unknown problem handle");
}

// Error::Config checks the reference performs before stepping, plus ABI sanity.
int validate(ivpb_ctx* ctx, const ProblemInfo& pi, const ivpb_options* o, int64_t N, double t0, double tf) {
  if (!o) return fail(ctx, IVPB_ERR_CONFIG, "options is null");
  if (N < 0) return fail(ctx, IVPB_ERR_CONFIG, "N must be >= 0");
  if (o->method < IVPB_RK23 || o->method > IVPB_BDF) return fail(ctx, IVPB_ERR_CONFIG, "unknown method");
  if (!o->rtol || !o->atol) return fail(ctx, IVPB_ERR_CONFIG, "rtol/atol must be given");
  // Tolerance::Vector length mismatch panics in the reference (src/methods/mod.rs:156-161)
  if (!(o->n_rtol == 1 || o->n_rtol == pi.n)) return fail(ctx, IVPB_ERR_CONFIG, "rtol length must be 1 or n");
  if (!(o->n_atol == 1 || o->n_atol == pi.n)) return fail(ctx, IVPB_ERR_CONFIG, "atol length must be 1 or n");
  // ConfigError::MustBePositive { max_steps } (e.g. src/methods/dop853.rs:178-184)
  if (o->has_max_steps && o->max_steps == 0) return fail(ctx, IVPB_ERR_CONFIG, "max_steps must be positive");
  if (o->has_t_eval && o->n_t_eval > 0 && !o->t_eval) return fail(ctx, IVPB_ERR_CONFIG, "t_eval is null");
  if (o->has_t_eval && o->n_t_eval < 0) return fail(ctx, IVPB_ERR_CONFIG, "n_t_eval < 0");
  if (!o->has_t_eval && o->max_out < 0) return fail(ctx, IVPB_ERR_CONFIG, "max_out < 0");
  if (pi.nev > ivpb::MAX_EVENTS_FN) return fail(ctx, IVPB_ERR_CONFIG, "too many event functions");
  if (pi.nev > 0 && o->max_events < 1) return fail(ctx, IVPB_ERR_CONFIG, "max_events must be >= 1 for a problem with events");
  if (o->n_event_cfg != 0 && o->n_event_cfg != pi.nev)
    return fail(ctx, IVPB_ERR_CONFIG, "n_event_cfg must be 0 or the problem's n_events");
  if (o->n_event_cfg != 0 && (!o->ev_direction || !o->ev_terminal_count))
    return fail(ctx, IVPB_ERR_CONFIG, "event config arrays are null");
  if (o->dense_output && o->max_segments < 1)
    return fail(ctx, IVPB_ERR_CONFIG, "dense_output needs max_segments >= 1 (interpolant segments stored per trajectory)");
  if (o->method == IVPB_RK4 && std::fabs(tf - t0) >= 1e-15) {
    // ConfigError::InvalidStepSize (src/methods/rk4.rs:84-90); h as chosen by solve_ivp.rs:185
    const double h = o->has_first_step ? o->first_step : (tf - t0) / 100.0;
    const double posneg = std::signbit(tf - t0) ? -1.0 : 1.0;
    const double sg = std::signbit(h) ? -1.0 : 1.0;
    if (h == 0.0 || sg != posneg || std::isnan(h))
      return fail(ctx, IVPB_ERR_CONFIG, "RK4: step size is zero or its sign does not match tf - t0");
  }
  if (o->method == IVPB_RADAU && std::fabs(tf - t0) >= 1e-15) {
    // radau.rs:250-262: the initial step is clamp(h, -hmax, hmax) -- f64::clamp panics for min > max -- and a
    // zero step is ConfigError::InvalidStepSize; the accepted-step clamp(hmin, hmax) needs hmin <= hmax (:752)
    const double hmax = o->has_max_step ? o->max_step : std::fabs(tf - t0);
    const double hmin = o->has_min_step ? o->min_step : 0.0;
    if (!(hmax >= 0.0)) return fail(ctx, IVPB_ERR_CONFIG, "RADAU: max_step must be non-negative");
    if (hmin > hmax) return fail(ctx, IVPB_ERR_CONFIG, "RADAU: min_step exceeds max_step");
    if (o->has_first_step && o->first_step == 0.0) return fail(ctx, IVPB_ERR_CONFIG, "RADAU: first_step is zero");
    if (hmax == 0.0) return fail(ctx, IVPB_ERR_CONFIG, "RADAU: max_step is zero");
    for (int i = 0; i < (o->n_rtol == 1 ? 1 : pi.n); ++i)
      if (!(o->rtol[i] > 0.0)) return fail(ctx, IVPB_ERR_CONFIG, "RADAU: rtol must be positive");
  }
  if (o->method == IVPB_BDF && std::fabs(tf - t0) >= 1e-15) {
    // bdf.rs:113-136 (negative tolerances), :192-197 (zero first step)
    for (int i = 0; i < (o->n_rtol == 1 ? 1 : pi.n); ++i)
      if (o->rtol[i] < 0.0) return fail(ctx, IVPB_ERR_CONFIG, "BDF: negative relative tolerance");
    for (int i = 0; i < (o->n_atol == 1 ? 1 : pi.n); ++i)
      if (o->atol[i] < 0.0) return fail(ctx, IVPB_ERR_CONFIG, "BDF: negative absolute tolerance");
    if (o->has_first_step && o->first_step == 0.0) return fail(ctx, IVPB_ERR_CONFIG, "BDF: first_step is zero");
  }
  if (o->has_jac_sparsity) {       // SparsityStructure (src/python/sparsity.rs:13-84): compressed columns, rows in range
    if (!o->jac_sparsity_colptr || !o->jac_sparsity_rows) return fail(ctx, IVPB_ERR_CONFIG, "has_jac_sparsity = 1 but the colptr / rows arrays are null");
    if (o->jac_sparsity_colptr[0] != 0) return fail(ctx, IVPB_ERR_CONFIG, "jac_sparsity_colptr[0] must be 0");
    for (int c = 0; c < pi.n; ++c) {
      if (o->jac_sparsity_colptr[c + 1] < o->jac_sparsity_colptr[c]) return fail(ctx, IVPB_ERR_CONFIG, "jac_sparsity_colptr must be non-decreasing");
      for (int k = o->jac_sparsity_colptr[c]; k < o->jac_sparsity_colptr[c + 1]; ++k)
        if (o->jac_sparsity_rows[k] < 0 || o->jac_sparsity_rows[k] >= pi.n) return fail(ctx, IVPB_ERR_CONFIG, "jac_sparsity row index out of range");
    }
  }
  if (o->jac_mode == 1 && !(pi.has_jac & 1) && (o->method == IVPB_RADAU || o->method == IVPB_BDF))
    return fail(ctx, IVPB_ERR_CONFIG, "jac_mode=1 but the problem has no analytic Jacobian");
  if (o->user_solout) {
    // Method::solve(.., Some(&mut user_solout)) (e.g. dop853.rs:114-127): DefaultSolOut's services do not exist on this path
    if (!(pi.has_jac & 4)) return fail(ctx, IVPB_ERR_CONFIG, "user_solout = 1, but the problem defines no SolOut hook (solout / ivp_solout)");
    if (o->has_t_eval || o->dense_output) return fail(ctx, IVPB_ERR_CONFIG, "user_solout = 1: t_eval / dense_output belong to DefaultSolOut; emit samples from the hook instead");
  }
  if (o->method == IVPB_RADAU) {
    // Options.mass_storage / nind1..3 reach RADAU only (solve_ivp.rs:246-258); DAE partition rules of radau.rs:210-245
    const bool has_mass = (pi.has_jac & 2) != 0;
    if (o->mass_storage != 0 && o->mass_storage != 1) return fail(ctx, IVPB_ERR_CONFIG, "mass_storage must be 0 (Identity) or 1 (Full)");
    if (o->mass_storage == 1 && !has_mass)
      return fail(ctx, IVPB_ERR_CONFIG, "mass_storage = Full, but the problem defines no mass matrix (IVP::mass / ivp_mass)");
    if (o->mass_storage == 0 && has_mass)
      return fail(ctx, IVPB_ERR_CONFIG, "the problem defines a mass matrix: RADAU needs mass_storage = Full (Identity storage cannot hold it)");
    const int64_t k1 = o->nind1, k2 = o->nind2 < 0 ? 0 : o->nind2, k3 = o->nind3 < 0 ? 0 : o->nind3;
    if (o->nind1 >= 0 || o->nind2 >= 0 || o->nind3 >= 0) {
      if (o->nind1 < 0 ? (k2 + k3 > pi.n) : (k1 + k2 + k3 != pi.n))
        return fail(ctx, IVPB_ERR_CONFIG, "RADAU: invalid DAE partition (ConfigError::InvalidDAEPartition: nind1 + nind2 + nind3 must equal n)");
      if ((k2 > 0 || k3 > 0) && !has_mass)
        return fail(ctx, IVPB_ERR_CONFIG, "index-2 / index-3 variables (nind2, nind3) need a problem with a mass matrix");
    }
  }
  return 0;
}

__global__ void zero_interval_kernel(KArgs a, int n, int nev, int method) {
  // reference src/solve/solve_ivp.rs:110-145: |xend - x0| < 1e-15 => Success, nothing evaluated.
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  if (a.status) a.status[i] = ivpb::ST_SUCCESS;
  if (a.counters) for (int c = 0; c < 6; ++c) a.counters[i * 6 + c] = 0u;
  if (a.t_final) a.t_final[i] = a.t0;
  if (a.y_final) for (int c = 0; c < n; ++c) a.y_final[i * n + c] = a.y0[i * n + c];
  if (a.h_next) a.h_next[i] = 0.0;
  int m = 0;
  if (a.n_t_eval >= 0) {
    for (int j = 0; j < a.n_t_eval; ++j) {
      if (fabs(a.t_eval[j] - a.t0) < 1e-12) {
        if (m < a.out_cap) {
          if (a.t_out) a.t_out[i * a.out_cap + m] = a.t_eval[j];
          if (a.y_out) for (int c = 0; c < n; ++c) a.y_out[(i * a.out_cap + m) * n + c] = a.y0[i * n + c];
        }
        ++m;
      }
    }
  } else if (a.out_cap > 0) {
    if (a.t_out) a.t_out[i * a.out_cap] = a.t0;
    if (a.y_out) for (int c = 0; c < n; ++c) a.y_out[i * a.out_cap * n + c] = a.y0[i * n + c];
    m = 1;
  }
  if (a.n_out) a.n_out[i] = m;
  if (a.ev_count) for (int e = 0; e < nev; ++e) a.ev_count[i * nev + e] = 0;
  if (a.seg_cap > 0) {   // ContinuousOutput::constant (src/solve/cont.rs:32-64): one segment at x0 with h = 1e-15
    a.seg_n[i] = 1;
    a.seg_x[2 * i * a.seg_cap] = a.t0;
    a.seg_x[2 * i * a.seg_cap + 1] = 1e-15;
    double* c = a.seg_cont + i * (long long)a.seg_cap * a.n_cont;
    for (int k = 0; k < a.n_cont; ++k) c[k] = 0.0;
    if (method == ivpb::M_BDF) for (int k = 0; k < n; ++k) { c[k * 7] = a.y0[i * n + k]; c[k * 7 + 6] = 1.0; }
    else for (int k = 0; k < n; ++k) c[k] = a.y0[i * n + k];
  }
}

// reference src/solve/solve_ivp.rs:148-176: y0.is_empty() => nothing to integrate; t = t_eval (all of it) or [x0, xend],
// every y an empty vector, Success, all counters zero.
__global__ void empty_state_kernel(KArgs a, int nev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  if (a.status) a.status[i] = ivpb::ST_SUCCESS;
  if (a.counters) for (int c = 0; c < 6; ++c) a.counters[i * 6 + c] = 0u;
  if (a.t_final) a.t_final[i] = a.tf;
  if (a.h_next) a.h_next[i] = 0.0;
  int m = 0;
  if (a.n_t_eval >= 0) {
    for (int j = 0; j < a.n_t_eval; ++j, ++m)
      if (m < a.out_cap && a.t_out) a.t_out[i * a.out_cap + m] = a.t_eval[j];
  } else {
    if (a.out_cap > 0 && a.t_out) a.t_out[i * a.out_cap] = a.t0;
    if (a.out_cap > 1 && a.t_out) a.t_out[i * a.out_cap + 1] = a.tf;
    m = 2;
  }
  if (a.n_out) a.n_out[i] = m;
  if (a.ev_count) for (int e = 0; e < nev; ++e) a.ev_count[i * nev + e] = 0;
  if (a.seg_cap > 0) {   // ContinuousOutput::constant with an empty state: one segment at x0
    a.seg_n[i] = 1;
    a.seg_x[2 * i * a.seg_cap] = a.t0;
    a.seg_x[2 * i * a.seg_cap + 1] = 1e-15;
  }
}

// ---- locality order ---------------------------------------------------------------------------------------------
// Trajectories are independent, so the order in which the work queue hands them out is free.  Neighbouring initial
// conditions (and parameters) mostly take the same accept / reject / Newton decisions, so a warp whose 32 lanes hold
// neighbours diverges much less than one holding 32 random members of the ensemble.  The shard is therefore ordered along
// a Morton curve through its first (up to three) coordinates -- y0[0], y0[1], ... then params -- before the solver kernel
// starts: a min/max pass, a key pass and one cub radix sort, all on the solve's stream (~0.1 ms per 2^20 trajectories).
// Outputs are still written at the trajectory's own index, so nothing is permuted back, and every trajectory's arithmetic
// is untouched (results are bit-identical with and without the order).
__device__ __forceinline__ unsigned long long ordered_bits(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
  return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}
__device__ __forceinline__ double sort_feature(const double* y0, const double* params, int n, int p, long long i, int f) {
  return f < n ? y0[i * n + f] : params[i * p + (f - n)];
}
__global__ void sort_minmax_kernel(const double* y0, const double* params, int n, int p, int nf, long long N,
                                   unsigned long long* mm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int f = 0; f < nf; ++f) {
    unsigned long long lo = ~0ull, hi = 0ull;
    if (i < N) {
      const double v = sort_feature(y0, params, n, p, i, f);
      if (v == v) lo = hi = ordered_bits(v);
    }
    for (int s = 16; s > 0; s >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xffffffffu, lo, s), oh = __shfl_xor_sync(0xffffffffu, hi, s);
      lo = ol < lo ? ol : lo; hi = oh > hi ? oh : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(mm + 2 * f, lo); atomicMax(mm + 2 * f + 1, hi); }
  }
}
__device__ __forceinline__ unsigned spread_bits(unsigned v, int nf) {      // v's bits at stride nf (Morton interleave)
  unsigned r = 0;
  for (int b = 0; b < 16; ++b) if ((unsigned)(b * nf) < 32u) r |= ((v >> b) & 1u) << (b * nf);
  return r;
}
// Key = Morton code of the first (up to three) coordinates that actually vary over the shard (constant ones carry no order).
__global__ void sort_keys_kernel(const double* y0, const double* params, int n, int p, int nscan, long long N,
                                 const unsigned long long* mm, unsigned* keys, unsigned* vals) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int sel[3], nsel = 0;
  for (int f = 0; f < nscan && nsel < 3; ++f)
    if (mm[2 * f + 1] > mm[2 * f]) sel[nsel++] = f;
  const int bits = nsel <= 1 ? 16 : (nsel == 2 ? 15 : 10);
  unsigned key = 0;
  for (int k = 0; k < nsel; ++k) {
    const int f = sel[k];
    const double lo = from_ordered_bits(mm[2 * f]), hi = from_ordered_bits(mm[2 * f + 1]);
    const double v = sort_feature(y0, params, n, p, i, f);
    double q = (v == v) ? (v - lo) / (hi - lo) : 0.0;
    q = fmin(fmax(q, 0.0), 1.0);
    const unsigned cell = (unsigned)(q * (double)((1u << bits) - 1u));
    key |= spread_bits(cell, nsel) << k;
  }
  keys[i] = key;
  vals[i] = (unsigned)i;
}

__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
  double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
  for (int i = 0; i < iters; ++i) {
    c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
    c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
}

// Fill KArgs from options (host side of the by-value kernel parameter).
void fill_args(KArgs& a, const ProblemInfo& pi, const ivpb_options* o, int64_t N, double t0, double tf) {
  std::memset(&a, 0, sizeof(a));
  a.N = N; a.t0 = t0; a.tf = tf;
  for (int i = 0; i < ivpb::MAX_N; ++i) {
    a.rtol[i] = o->rtol[o->n_rtol == 1 ? 0 : (i < pi.n ? i : 0)];
    a.atol[i] = o->atol[o->n_atol == 1 ? 0 : (i < pi.n ? i : 0)];
  }
  a.first_step = o->first_step; a.max_step = o->max_step; a.min_step = o->min_step;
  a.has_first_step = o->has_first_step; a.has_max_step = o->has_max_step; a.has_min_step = o->has_min_step;
  a.static_sched = (o->flags & IVPB_FLAG_NO_REFILL) ? 1 : 0;
  a.max_steps = o->has_max_steps ? o->max_steps : std::numeric_limits<u64>::max();   // usize::MAX
  a.n_t_eval = o->has_t_eval ? o->n_t_eval : -1;
  a.out_cap = o->has_t_eval ? o->n_t_eval + 1 : o->max_out;
  a.max_events = o->max_events;
  a.jac_mode = o->jac_mode;
  {   // resolved DAE partition (radau.rs:210-245; validated above)
    const int k2 = o->nind2 < 0 ? 0 : o->nind2, k3 = o->nind3 < 0 ? 0 : o->nind3;
    a.nind2 = k2; a.nind3 = k3;
    a.nind1 = (o->nind1 < 0 && o->nind2 < 0 && o->nind3 < 0) ? pi.n : (o->nind1 < 0 ? pi.n - k2 - k3 : o->nind1);
  }
  a.seg_cap = o->dense_output ? o->max_segments : 0;
  a.n_cont = coeffs_per_state(o->method) * pi.n;
  if (o->method == IVPB_RADAU) {
    // Tolerance transform and Newton tolerance of radau.rs:188-205, evaluated here with the host libm (the
    // same pow the reference calls) so every trajectory sees bit-identical scaled tolerances.
    const double uround = 2.3e-16;
    for (int i = 0; i < ivpb::MAX_N; ++i) {
      const double rt = a.rtol[i], at = a.atol[i];
      const double quot = at / rt;
      a.rtol[i] = 0.1 * std::pow(rt, 2.0 / 3.0);
      a.atol[i] = a.rtol[i] * quot;
    }
    const double tolst = a.rtol[0];
    a.newton_tol = std::fmax(10.0 * uround / tolst, std::fmin(0.03, std::sqrt(tolst)));
  } else if (o->method == IVPB_BDF) {
    // bdf.rs:174-184
    const double eps = std::numeric_limits<double>::epsilon();
    double rtol_min = std::numeric_limits<double>::infinity();      // over ALL n components (Tolerance::iter(n)), also n > MAX_N
    for (int i = 0; i < pi.n; ++i) rtol_min = std::fmin(rtol_min, o->rtol[o->n_rtol == 1 ? 0 : i]);
    rtol_min = std::fmax(rtol_min, eps);
    a.newton_tol = std::fmax(10.0 * eps / rtol_min, std::fmin(std::sqrt(rtol_min), 0.03));
    if (a.newton_tol <= 0.0) a.newton_tol = 1e-9;
  }
  for (int e = 0; e < ivpb::MAX_EVENTS_FN; ++e) { a.ev_dir[e] = 0; a.ev_term[e] = -1; }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Enqueue one shard on one device.  All pointers in `d` are device pointers on dev.
// Fills dev.sort_vals with the locality order of the shard; *perm = null when the shard is not worth ordering.
static int locality_order(ivpb_ctx* ctx, Device& dev, const ProblemInfo& pi, int64_t N, const double* d_y0,
                          const double* d_params, cudaStream_t stream, const unsigned** perm) {
  *perm = nullptr;
  if (N < 4096 || N >= (int64_t)1 << 31) return 0;
  const int nscan = std::min(8, pi.n + (d_params ? pi.p : 0));     // coordinates examined: y0[0..], then the parameters
  if (nscan < 1) return 0;
  CK(dev.sort_keys.ensure(sizeof(unsigned) * 2 * N));
  CK(dev.sort_vals.ensure(sizeof(unsigned) * 2 * N));
  CK(dev.sort_minmax.ensure(sizeof(unsigned long long) * 16));
  unsigned *k_in = (unsigned*)dev.sort_keys.p, *k_out = k_in + N, *v_in = (unsigned*)dev.sort_vals.p, *v_out = v_in + N;
  unsigned long long* mm = (unsigned long long*)dev.sort_minmax.p;
  unsigned long long init[16];
  for (int f = 0; f < 8; ++f) { init[2 * f] = ~0ull; init[2 * f + 1] = 0ull; }
  CK(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  const int blk = 256, grid = (int)((N + blk - 1) / blk);
  sort_minmax_kernel<<<grid, blk, 0, stream>>>(d_y0, d_params, pi.n, pi.p, nscan, (long long)N, mm);
  sort_keys_kernel<<<grid, blk, 0, stream>>>(d_y0, d_params, pi.n, pi.p, nscan, (long long)N, mm, k_in, v_in);
  CK(cudaGetLastError());
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, (int)N, 0, 32, stream));
  CK(dev.sort_tmp.ensure(tmp));
  CK(cub::DeviceRadixSort::SortPairs(dev.sort_tmp.p, tmp, k_in, k_out, v_in, v_out, (int)N, 0, 32, stream));
  ctx->launches += 3;
  *perm = v_out;
  return 0;
}

// Per-component tolerances of a warp-per-trajectory launch (explicit n > 32, RADAU / BDF n > 8) with Tolerance::Vector:
// what the step code expects in KArgs::rtol/atol, i.e. for RADAU the transformed tolerances of radau.rs:188-196.
static bool wants_tol_ext(const ProblemInfo& pi, const ivpb_options* o) {
  const bool implicit_m = o->method == IVPB_RADAU || o->method == IVPB_BDF;
  return (pi.n > ivpb::MAX_N || (implicit_m && pi.n > 8)) && (o->n_rtol > 1 || o->n_atol > 1);
}
// Column groups of a Jacobian structure: the reference's greedy first-fit rule (group_columns, src/python/sparsity.rs:109-154)
// -- columns in index order, each into the first group none of whose columns shares a structural non-zero row with it.
// One bitset of "rows taken" per group.
static int group_columns(int n, const int32_t* colptr, const int32_t* rows, std::vector<int32_t>& group) {
  const size_t words = ((size_t)n + 63) / 64;
  std::vector<std::vector<uint64_t>> taken;
  group.assign((size_t)n, 0);
  for (int c = 0; c < n; ++c) {
    size_t g = 0;
    for (; g < taken.size(); ++g) {
      bool clash = false;
      for (int k = colptr[c]; k < colptr[c + 1] && !clash; ++k) clash = (taken[g][(size_t)rows[k] >> 6] >> (rows[k] & 63)) & 1u;
      if (!clash) break;
    }
    if (g == taken.size()) taken.emplace_back(words, 0ull);
    for (int k = colptr[c]; k < colptr[c + 1]; ++k) taken[g][(size_t)rows[k] >> 6] |= 1ull << (rows[k] & 63);
    group[(size_t)c] = (int32_t)g;
  }
  return (int)taken.size();
}
// jac_sparsity is used by the warp-cooperative RADAU / BDF kernels (n > 8) with finite differences.  The thread-per-
// trajectory kernels (n <= 8) keep dense differences: for a structure that covers the true dependencies both give the same
// bits (a column's rows do not depend on the other columns of its group; entries outside the structure are (f - f) / h = 0),
// and at n <= 8 there is nothing to save.
static bool wants_sparsity(const ProblemInfo& pi, const ivpb_options* o) {
  return o->has_jac_sparsity && (o->method == IVPB_RADAU || o->method == IVPB_BDF) && pi.n > 8 &&
         !(o->jac_mode == 1 && (pi.has_jac & 1));
}
// Upload the small per-call arrays every launch of the call reads (t_eval, per-component tolerances) on `stream`.
static int stage_shared(ivpb_ctx* ctx, Device& dev, const ProblemInfo& pi, const ivpb_options* o, cudaStream_t stream) {
  if (o->has_t_eval && o->n_t_eval > 0) {
    CK(dev.t_eval.ensure(sizeof(double) * o->n_t_eval));
    CK(cudaMemcpyAsync(dev.t_eval.p, o->t_eval, sizeof(double) * o->n_t_eval, cudaMemcpyHostToDevice, stream));
  }
  if (wants_tol_ext(pi, o)) {
    std::vector<double> tol(2 * (size_t)pi.n);
    for (int i = 0; i < pi.n; ++i) {
      double rt = o->rtol[o->n_rtol == 1 ? 0 : i], at = o->atol[o->n_atol == 1 ? 0 : i];
      if (o->method == IVPB_RADAU) {        // host libm pow, as in fill_args
        const double quot = at / rt;
        rt = 0.1 * std::pow(rt, 2.0 / 3.0);
        at = rt * quot;
      }
      tol[i] = rt; tol[pi.n + i] = at;
    }
    CK(dev.tol_ext.ensure(sizeof(double) * tol.size()));
    CK(cudaMemcpyAsync(dev.tol_ext.p, tol.data(), sizeof(double) * tol.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));      // `tol` is a stack-lifetime staging buffer
  }
  if (wants_sparsity(pi, o)) {      // [colptr (n + 1) | group (n) | rows (nnz)] as int32
    const int n = pi.n, nnz = o->jac_sparsity_colptr[n];
    std::vector<int32_t> buf((size_t)2 * n + 1 + (size_t)std::max(nnz, 1)), group;
    dev.sp_ngroups = group_columns(n, o->jac_sparsity_colptr, o->jac_sparsity_rows, group);
    std::memcpy(buf.data(), o->jac_sparsity_colptr, sizeof(int32_t) * ((size_t)n + 1));
    std::memcpy(buf.data() + n + 1, group.data(), sizeof(int32_t) * (size_t)n);
    if (nnz > 0) std::memcpy(buf.data() + 2 * n + 1, o->jac_sparsity_rows, sizeof(int32_t) * (size_t)nnz);
    CK(dev.sparsity.ensure(sizeof(int32_t) * buf.size()));
    CK(cudaMemcpyAsync(dev.sparsity.p, buf.data(), sizeof(int32_t) * buf.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
  }
  return 0;
}

// How many warps of a warp-cooperative implicit launch get a slot in KArgs::scratch: every warp the device can hold
// (64 per SM), no more than there are trajectories, and no more than fit 16 GB (large n: 6.4 MB per warp at n = 400).
static long long warp_pool_warps(const ivpb::WarpImplShape& sh, int sms, int64_t N) {
  const long long by_device = (long long)sms * 64, by_bytes = (16ll << 30) / (8 * std::max<long long>(1, sh.scratch_doubles));
  return std::max<long long>(1, std::min<long long>(std::min<long long>(by_device, by_bytes), std::max<int64_t>(N, 1)));
}

// `slot`: which of the device's work-queue heads this launch uses.  `upload_shared`: copy t_eval / per-component
// tolerances to the device's staging buffers on `stream` (false for the later chunks of one call: already there).
static int launch_shard(ivpb_ctx* ctx, Device& dev, int problem, const ProblemInfo& pi, const ivpb_options* o,
                        int64_t N, double t0, double tf, const double* d_y0, const double* d_params,
                        const ivpb_outputs* d, cudaStream_t stream, bool zero_tail, int slot, bool upload_shared,
                        int64_t chunk_size = 0, int64_t in_chunk = 0) {
  if (N == 0) return 0;
  KArgs a;
  fill_args(a, pi, o, N, t0, tf);
  a.zero_tail = zero_tail ? 1 : 0;
  if (in_chunk > 0) { a.in_chunk = in_chunk; a.in_flag = dev.in_flag; }
  if (chunk_size > 0) { a.chunk_size = chunk_size; a.chunk_count = dev.chunk_count; a.chunk_flag = dev.chunk_flag_dev; }
  // Locality order: on by default for RADAU / BDF, where warp divergence is the bottleneck (Robertson BDF 46.5 -> 23.1 ms,
  // VdP mu=1000 RADAU 54.0 -> 50.1, BDF 79.3 -> 73.2 per 2^18 trajectories); opt-in for the explicit methods, where it gains
  // little on the device (north star 15.03 -> 14.98 ms, CR3BP + t_eval 332 -> 310, chaotic Lorenz 7.86 -> 8.22) and costs
  // a lot end to end when the caller's buffers are mapped host memory (rows are then read and written out of order over
  // PCIe: north star e2e 15.3 -> 19.3 ms, CR3BP 392 -> 529 ms).
  const bool implicit_m = o->method == IVPB_RADAU || o->method == IVPB_BDF;
  const bool want_order = (o->flags & IVPB_FLAG_SORT) || (implicit_m && !(o->flags & IVPB_FLAG_NO_SORT));
  if (want_order && std::fabs(tf - t0) >= 1e-15) {
    const unsigned* perm = nullptr;
    if (int rc = locality_order(ctx, dev, pi, N, d_y0, d_params, stream, &perm)) return rc;
    a.perm = perm;
  }
  a.y0 = d_y0; a.params = d_params; a.queue = dev.queue + slot;
  a.status = d->status; a.counters = d->counters; a.t_final = d->t_final; a.y_final = d->y_final;
  a.h_next = d->h_next; a.n_out = d->n_out; a.t_out = d->t_out; a.y_out = d->y_out;
  a.ev_count = d->ev_count; a.ev_t = d->ev_t; a.ev_y = d->ev_y;
  a.seg_n = d->n_seg; a.seg_x = d->seg_x; a.seg_cont = d->seg_cont;
  a.vec_io = ((reinterpret_cast<uintptr_t>(d_y0) | reinterpret_cast<uintptr_t>(d->y_final)) & 15u) == 0 && (pi.n % 2 == 0);
  if (!a.seg_n || !a.seg_x || (!a.seg_cont && pi.n > 0)) a.seg_cap = 0;      // nowhere to log the segments
  if (a.out_cap == 0 || (!a.t_out && !a.y_out && !a.n_out)) {
    if (!o->has_t_eval) a.out_cap = 0;    // nothing to store in step mode
  }
  if (upload_shared)
    if (int rc = stage_shared(ctx, dev, pi, o, stream)) return rc;
  if (o->has_t_eval && o->n_t_eval > 0) a.t_eval = (const double*)dev.t_eval.p;
  if (wants_tol_ext(pi, o)) { a.rtol_ext = (const double*)dev.tol_ext.p; a.atol_ext = a.rtol_ext + pi.n; }
  if (wants_sparsity(pi, o)) {
    a.sp_colptr = (const int*)dev.sparsity.p; a.sp_group = a.sp_colptr + pi.n + 1; a.sp_rows = a.sp_group + pi.n;
    a.sp_ngroups = dev.sp_ngroups;
  }
  const bool implicit_method = o->method == IVPB_RADAU || o->method == IVPB_BDF;
  const int strict = ctx->fp_override >= 0 ? ctx->fp_override
                                           : ((o->flags & IVPB_FLAG_STRICT_FP) || (implicit_method && !(o->flags & IVPB_FLAG_FAST_FP))) ? 1 : 0;
  const int block = 128;
  const bool warp_mode = pi.n > ivpb::MAX_N;      // one trajectory per warp (WarpLayout, ivpb_erk.cuh)

  if (pi.n == 0 && std::fabs(tf - t0) >= 1e-15) {
    const int grid = (int)((N + block - 1) / block);
    if (!a.seg_n || !a.seg_x) a.seg_cap = 0;
    empty_state_kernel<<<grid, block, 0, stream>>>(a, pi.nev);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return 0;
  }
  if (std::fabs(tf - t0) < 1e-15) {
    const int grid = (int)((N + block - 1) / block);
    zero_interval_kernel<<<grid, block, 0, stream>>>(a, pi.n, pi.nev, o->method);
    CK(cudaGetLastError());
    ctx->launches += 1;
    return 0;
  }

  // event configuration: options override, else the problem's IVP::event_config defaults
  if (pi.nev > 0) {
    if (o->n_event_cfg > 0) {
      for (int e = 0; e < pi.nev; ++e) {
        a.ev_dir[e] = o->ev_direction[e] > 0 ? 1 : (o->ev_direction[e] < 0 ? -1 : 0);
        a.ev_term[e] = o->ev_terminal_count[e];
      }
    } else {
      for (int e = 0; e < pi.nev; ++e) { a.ev_dir[e] = pi.ev_dir[e]; a.ev_term[e] = pi.ev_term[e]; }
    }
  }
  const bool want_out = o->has_t_eval || a.out_cap > 0 || a.seg_cap > 0;
  int feat = 0;
  if (o->user_solout) feat = 4;        // K_USER: the problem's own SolOut replaces DefaultSolOut
  else if (pi.nev > 0) feat = 3;       // K_OUT | K_EVENTS: events always run (they can terminate)
  else if (want_out) feat = 1;         // K_OUT

  CK(cudaMemsetAsync(dev.queue + slot, 0, sizeof(u64), stream));

  if (pi.user) {
    long long max_warps = 0;
    if (pi.n > 8 && (o->method == IVPB_RADAU || o->method == IVPB_BDF)) {
      // warp-cooperative implicit kernels keep one Jacobian per resident warp in global memory; the launch shape is
      // chosen inside ivpb_nvrtc_launch, so size the pool for the most warps an SM can hold (64 warps x SMs)
      // (+ the mass matrix next to it for RADAU problems that have one)
      const ivpb::WarpImplShape sh = ivpb::warp_impl_shape(pi.n, o->method, (pi.has_jac & 2) != 0);
      max_warps = warp_pool_warps(sh, dev.sms, N);
      CK(dev.scratch.ensure(sizeof(double) * (size_t)sh.scratch_doubles * (size_t)max_warps));
      a.scratch = (double*)dev.scratch.p;
    }
    int rc = ivpb_nvrtc_launch(ctx, ctx->user[pi.uidx], dev.id, dev.sms, o->method, feat, strict, &a, sizeof(a), N,
                               a.static_sched, max_warps, stream);
    if (rc) return rc;
    ctx->launches += 1;
    return 0;
  }

  // flavour: 0 FMA build, 1 strict build, 2 strict build with deferred guards.  `probe`: only report whether the kernel exists.
  auto launch_flavour = [&](int flavour, KArgs& a, bool probe) -> int {
  const void* kern = nullptr;
  int kblock = block, ksmem = 0, kunits = 0;
  if (o->method == IVPB_RADAU || o->method == IVPB_BDF) {
    kern = BUILTIN_IMPL[problem][flavour](o->method, feat, &kblock, &ksmem, &kunits);
    if (probe) return kern ? 0 : 1;
    if (!kern) return fail(ctx, IVPB_ERR_CONFIG, "implicit methods: the per-warp vectors of this state size do not fit shared memory");
    if (ksmem > 0) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ksmem));
  } else {
    ivpb_pinfo kinfo;
    kinfo.block = block;
    // no arrival / completion flags in this launch (device-resident entry point, or a host-buffer solve too small to
    // pipeline): take the twin compiled without them (K_NOPIPE, ivpb_kernels.cuh)
    const int nopipe = (a.chunk_size <= 0 && a.in_chunk <= 0 && !(feat & 4)) ? 0x100 : 0;
    kern = BUILTIN[problem][flavour](o->method, feat | nopipe, &kinfo);
    if (probe) return kern ? 0 : 1;
    if (!kern) return fail(ctx, IVPB_ERR_CONFIG, "no kernel for this problem/method/feature combination");
    kblock = kinfo.block;
    if (warp_mode) {
      ksmem = (kblock / 32) * 2 * pi.n * (int)sizeof(double);
      if (ksmem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ksmem));
    }
  }
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kblock, ksmem));
  if (occ < 1) occ = 1;
  static const int occ_cap = std::getenv("IVPB_GRID_OCC") ? std::atoi(std::getenv("IVPB_GRID_OCC")) : 0;      // A/B measurements
  if (occ_cap > 0 && occ > occ_cap) occ = occ_cap;
  // Small shards (strong scaling): when the full grid holds a thread for every trajectory the queue has nothing to refill
  // and every warp runs as long as its slowest lane.  Two resident blocks fewer per SM give each thread >= 1.35
  // trajectories, so refill evens the step counts out again.  Measured on the north star at 131072 trajectories (one of
  // eight shards of the 2^20 ensemble): 7 / 6 / 5 / 4 / 3 blocks per SM -> 2.036 / 2.078 / 1.976 / 2.136 / 2.369 ms;
  // at 262144 and above the full grid wins (4.12 vs 4.19 ms, 7.47 vs 7.72 ms).  Only the kernels that keep >= 7 blocks.
  if (occ_cap == 0 && occ >= 7 && !a.static_sched && !warp_mode && kunits == 0 && N <= (int64_t)dev.sms * occ * kblock &&
      (double)N >= 1.35 * (double)((int64_t)dev.sms * (occ - 2) * kblock))
    occ -= 2;
  int64_t grid = (int64_t)dev.sms * occ;
  const bool impl_warp = kunits < 0;
  if (impl_warp) kunits = -kunits;
  const int64_t units_per_block = kunits > 0 ? kunits : (warp_mode ? kblock / 32 : kblock);
  const int64_t need = (N + units_per_block - 1) / units_per_block;
  if (a.static_sched || need < grid) grid = need;
  if (impl_warp) {       // per-warp slots in global memory (L2-resident): Jacobian (+ mass, + the iteration matrices for large n)
    const ivpb::WarpImplShape sh = ivpb::warp_impl_shape(pi.n, o->method, (pi.has_jac & 2) != 0);
    const int64_t wpb = kblock / 32;
    grid = std::min<int64_t>(grid, std::max<int64_t>(1, warp_pool_warps(sh, dev.sms, N) / wpb));
    CK(dev.scratch.ensure(sizeof(double) * (size_t)sh.scratch_doubles * (size_t)grid * (size_t)wpb));
    a.scratch = (double*)dev.scratch.p;
  }
  void* kargs[] = {&a};
  CK(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(kblock), kargs, ksmem, stream));
  ctx->launches += 1;
  return 0;
  };
  // A strict solve takes two launches when the problem has a `strictd` kernel (thread-per-trajectory kernels of the
  // built-in problems, ivpb_exact.cuh): the first runs every division / square root on its fast path and lists the
  // trajectories in which a guard failed; the second -- the guarded build, one persistent grid that reads the count on
  // the device and normally finds it zero -- integrates those from the start.  No host synchronisation in between.
  static const bool no_defer = std::getenv("IVPB_NO_DEFER") != nullptr;      // A/B measurements
  if (strict == 1 && !no_defer && N < ((int64_t)1 << 32) && launch_flavour(2, a, true) == 0) {
    CK(dev.rerun.ensure(sizeof(unsigned) * ((size_t)N + 1)));
    unsigned* count = (unsigned*)dev.rerun.p;
    CK(cudaMemsetAsync(count, 0, sizeof(unsigned), stream));
    KArgs a1 = a;
    a1.rerun_count = count; a1.rerun_list = count + 1;
    if (const char* dbg = std::getenv("IVPB_DEBUG_RERUN")) a1.debug_rerun_mod = std::atoi(dbg);
    if (int rc = launch_flavour(2, a1, false)) return rc;
    KArgs a2 = a;
    a2.perm = count + 1; a2.n_dev = count; a2.static_sched = 0;
    CK(cudaMemsetAsync(dev.queue + slot, 0, sizeof(u64), stream));
    return launch_flavour(1, a2, false);
  }
  return launch_flavour(strict, a, false);
}

// ------------------------------------------------------------------------------------------------
// Floating-point mode of a solve (which of the two kernel builds runs).
//
//   strict  every reference operation one for one (no FMA contraction, correctly rounded div / sqrt, glibc's pow):
//           bit-identical to the reference's operation sequence on every workload, ~2x the instructions.
//   fma     contracted multiply-adds and slow-path-free controller arithmetic: a few ulp per step away from it.
//
// A few ulp are harmless for a well-conditioned ensemble (north star: all values inside max(10 rtol |y|, 10 atol), step
// counts equal) and fatal for an ill-conditioned one (CR3BP at rtol 1e-10: orbits amplify a last-bit difference by 1e7,
// 19 % of the trajectories stay inside the tolerance).  Which of the two a given ensemble is cannot be read off the
// options, so unless the caller decides (IVPB_FLAG_STRICT_FP / IVPB_FLAG_FAST_FP) the runtime MEASURES it: the first solve
// of a configuration integrates a sample of the ensemble (up to 8192 trajectories, evenly strided) with both builds and
// compares them on the device; the FMA build is used only if every sampled trajectory ends with the same status,
// inside the north-star tolerance of the strict result, and >= 99 % of them with identical accepted / rejected step
// counts.  The verdict is cached per configuration (problem, method, tolerances, span, step limits, t_eval, events);
// both pilot kernels fit in one wave, so the pilot costs about two single-trajectory latencies, once.
// RADAU / BDF stay strict by default as before (IVPB_FLAG_FAST_FP opts out): stiff ensembles lose step-count parity under
// any perturbation and gain only 17-29 % from the FMA build.
namespace {

__global__ void pilot_gather_kernel(const double* src, int w, long long N, int S, double* dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  const long long i = (long long)k * N / S;
  for (int c = 0; c < w; ++c) dst[(long long)k * w + c] = src[i * w + c];
}

struct PilotTol { double rtol[ivpb::MAX_N], atol[ivpb::MAX_N]; };

// res[0] status mismatches, res[1] accepted / rejected step-count mismatches, res[2] trajectories outside
// max(10 rtol |y|, 10 atol) (final state, elementwise; final time likewise with the first component's tolerances)
// The sample has to stay inside HALF the north-star tolerance max(10 rtol |y|, 10 atol): the margin is for the trajectories
// the pilot does not see (VdP DOPRI5 at 1e-6: all 8192 sampled trajectories inside the full tolerance, 2 of 32768 outside).
#define PILOT_TOL_FACTOR 5.0
__global__ void pilot_compare_kernel(int S, int n, PilotTol tol, const int* st_a, const unsigned* cn_a, const double* tf_a,
                                     const double* yf_a, const int* st_b, const unsigned* cn_b, const double* tf_b,
                                     const double* yf_b, unsigned* res) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= S) return;
  if (st_a[k] != st_b[k]) atomicAdd(res + 0, 1u);
  if (cn_a[k * 6 + 4] != cn_b[k * 6 + 4] || cn_a[k * 6 + 5] != cn_b[k * 6 + 5]) atomicAdd(res + 1, 1u);
  bool bad = false;
  for (int c = 0; c < n; ++c) {
    const int tc = c < ivpb::MAX_N ? c : 0;
    const double ref = yf_a[(long long)k * n + c], d = fabs(yf_b[(long long)k * n + c] - ref);
    if (!(d <= fmax(PILOT_TOL_FACTOR * tol.rtol[tc] * fabs(ref), PILOT_TOL_FACTOR * tol.atol[tc]))) bad = true;
  }
  if (!(fabs(tf_b[k] - tf_a[k]) <= fmax(PILOT_TOL_FACTOR * tol.rtol[0] * fabs(tf_a[k]), PILOT_TOL_FACTOR * tol.atol[0]))) bad = true;
  if (bad) atomicAdd(res + 2, 1u);
}

uint64_t fnv(uint64_t h, const void* p, size_t nbytes) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < nbytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

uint64_t fp_key(int problem, const ProblemInfo& pi, const ivpb_options* o, double t0, double tf) {
  uint64_t h = 1469598103934665603ull;
  h = fnv(h, &problem, sizeof(problem));
  h = fnv(h, &o->method, sizeof(o->method));
  h = fnv(h, o->rtol, sizeof(double) * (o->n_rtol == 1 ? 1 : pi.n));
  h = fnv(h, o->atol, sizeof(double) * (o->n_atol == 1 ? 1 : pi.n));
  h = fnv(h, &t0, 8); h = fnv(h, &tf, 8);
  const int has[4] = {o->has_first_step, o->has_max_step, o->has_max_steps, o->has_t_eval};
  h = fnv(h, has, sizeof(has));
  if (o->has_first_step) h = fnv(h, &o->first_step, 8);
  if (o->has_max_step) h = fnv(h, &o->max_step, 8);
  if (o->has_max_steps) h = fnv(h, &o->max_steps, 8);
  if (o->has_t_eval && o->n_t_eval > 0) h = fnv(h, o->t_eval, sizeof(double) * o->n_t_eval);
  if (o->n_event_cfg > 0) {
    h = fnv(h, o->ev_direction, sizeof(int32_t) * o->n_event_cfg);
    h = fnv(h, o->ev_terminal_count, sizeof(int64_t) * o->n_event_cfg);
  }
  const int misc[3] = {o->user_solout, o->max_out > 0 ? 1 : 0, o->max_events};
  h = fnv(h, misc, sizeof(misc));
  return h;
}

}  // namespace

// Decide the floating-point mode of this solve; *strict_out = 1 (strict build) or 0 (FMA build).  `on_device`: y0 / params
// are device-accessible pointers (device 0) valid in `stream` order; otherwise they are host pointers.
static int resolve_fp(ivpb_ctx* ctx, int problem, const ProblemInfo& pi, const ivpb_options* o, int64_t N, double t0,
                      double tf, const double* y0, const double* params, bool on_device, cudaStream_t stream,
                      int* strict_out) {
  const bool implicit_m = o->method == IVPB_RADAU || o->method == IVPB_BDF;
  int32_t* L = ctx->last_fp;
  for (int i = 0; i < 6; ++i) L[i] = 0;
  if (o->flags & IVPB_FLAG_STRICT_FP) { L[0] = 1; L[1] = 0; *strict_out = 1; return 0; }
  if (o->flags & IVPB_FLAG_FAST_FP) { L[0] = 0; L[1] = 0; *strict_out = 0; return 0; }
  if (implicit_m) { L[0] = 1; L[1] = 1; *strict_out = 1; return 0; }           // method default
  if (N == 0 || pi.n == 0 || std::fabs(tf - t0) < 1e-15) { L[1] = 1; *strict_out = 0; return 0; }
  const uint64_t key = fp_key(problem, pi, o, t0, tf);
  for (const auto& c : ctx->fp_cache)
    if (c.key == key) {
      L[0] = c.strict; L[1] = 2;
      for (int i = 0; i < 4; ++i) L[2 + i] = c.info[i];
      *strict_out = c.strict;
      return 0;
    }
  // ---- pilot: the same sample with both builds, on device 0 ----
  Device& dev = ctx->devs[0];
  CK(cudaSetDevice(dev.id));
  const int S = (int)std::min<int64_t>(N, 8192);
  const int n = pi.n, p = pi.p;
  CK(ctx->pilot_in.ensure(sizeof(double) * (size_t)S * (n + p)));
  double* s_y0 = (double*)ctx->pilot_in.p;
  double* s_par = s_y0 + (size_t)S * n;
  if (on_device) {
    pilot_gather_kernel<<<(S + 127) / 128, 128, 0, stream>>>(y0, n, (long long)N, S, s_y0);
    if (p > 0) pilot_gather_kernel<<<(S + 127) / 128, 128, 0, stream>>>(params, p, (long long)N, S, s_par);
    CK(cudaGetLastError());
  } else {
    std::vector<double> h((size_t)S * (n + p));
    for (int k = 0; k < S; ++k) {
      const int64_t i = (int64_t)k * N / S;
      std::memcpy(&h[(size_t)k * n], y0 + i * n, sizeof(double) * n);
      if (p > 0) std::memcpy(&h[(size_t)S * n + (size_t)k * p], params + i * p, sizeof(double) * p);
    }
    CK(cudaMemcpyAsync(s_y0, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
  }
  const size_t per_traj = 4 + 24 + 8 + 8 * (size_t)n;
  ivpb_outputs po[2];
  for (int m = 0; m < 2; ++m) {
    CK(ctx->pilot_out[m].ensure(per_traj * S + 64));
    char* b = (char*)ctx->pilot_out[m].p;
    std::memset(&po[m], 0, sizeof(po[m]));
    po[m].y_final = (double*)b; b += 8 * (size_t)n * S;
    po[m].t_final = (double*)b; b += 8 * (size_t)S;
    po[m].counters = (uint32_t*)b; b += 24 * (size_t)S;
    po[m].status = (int32_t*)b;
  }
  ivpb_options po_opt = *o;
  po_opt.dense_output = 0;
  for (int m = 0; m < 2; ++m) {       // m = 0: strict (the reference's arithmetic), m = 1: FMA
    ctx->fp_override = m == 0 ? 1 : 0;
    const int rc = launch_shard(ctx, dev, problem, pi, &po_opt, S, t0, tf, s_y0, p > 0 ? s_par : nullptr, &po[m], stream,
                                false, 0, true);
    ctx->fp_override = -1;
    if (rc) return rc;
  }
  CK(ctx->pilot_res.ensure(16));
  CK(cudaMemsetAsync(ctx->pilot_res.p, 0, 16, stream));
  PilotTol tol;
  for (int i = 0; i < ivpb::MAX_N; ++i) {
    tol.rtol[i] = o->rtol[o->n_rtol == 1 ? 0 : (i < n ? i : 0)];
    tol.atol[i] = o->atol[o->n_atol == 1 ? 0 : (i < n ? i : 0)];
  }
  if (n > ivpb::MAX_N) {      // warp-per-trajectory systems: compare with the tightest tolerance
    for (int i = 0; i < n; ++i) {
      tol.rtol[0] = std::fmin(tol.rtol[0], o->rtol[o->n_rtol == 1 ? 0 : i]);
      tol.atol[0] = std::fmin(tol.atol[0], o->atol[o->n_atol == 1 ? 0 : i]);
    }
  }
  pilot_compare_kernel<<<(S + 127) / 128, 128, 0, stream>>>(S, n, tol, po[0].status, po[0].counters, po[0].t_final,
                                                           po[0].y_final, po[1].status, po[1].counters, po[1].t_final,
                                                           po[1].y_final, (unsigned*)ctx->pilot_res.p);
  CK(cudaGetLastError());
  unsigned res[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(res, ctx->pilot_res.p, 12, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  ctx->launches += 2 + (on_device ? (p > 0 ? 2 : 1) : 0);
  const bool fma_ok = res[0] == 0 && res[2] == 0 && (uint64_t)res[1] * 100 <= (uint64_t)S;
  ivpb_ctx::FpChoice c;
  c.key = key; c.strict = fma_ok ? 0 : 1;
  c.info[0] = S; c.info[1] = (int32_t)res[0]; c.info[2] = (int32_t)res[1]; c.info[3] = (int32_t)res[2]; c.info[4] = 0;
  if (ctx->fp_cache.size() >= 64) ctx->fp_cache.erase(ctx->fp_cache.begin());
  ctx->fp_cache.push_back(c);
  L[0] = c.strict; L[1] = 3;
  for (int i = 0; i < 4; ++i) L[2 + i] = c.info[i];
  *strict_out = c.strict;
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

int ivpb_create(ivpb_ctx** out, const int* device_ids, int n_devices) {
  if (!out) return IVPB_ERR_CONFIG;
  *out = nullptr;
  ivpb_ctx* ctx = nullptr;   // for CK(): errors before the context exists go to g_create_err
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, IVPB_ERR_CUDA, std::string("no CUDA device available (libivpb has no CPU fallback): ") +
                                            cudaGetErrorString(e));
  std::vector<int> ids;
  if (device_ids && n_devices > 0) ids.assign(device_ids, device_ids + n_devices);
  else { int cur = 0; CK(cudaGetDevice(&cur)); ids.push_back(cur); }
  ivpb_ctx* c = new ivpb_ctx();
  for (int id : ids) {
    if (id < 0 || id >= count) { ivpb_destroy(c); return fail(nullptr, IVPB_ERR_CONFIG, "device id out of range"); }
    Device d;
    d.id = id;
    cudaError_t err = cudaSetDevice(id);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, id);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&d.stream2, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&d.ev_shared, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d.queue, sizeof(u64) * MAX_CHUNKS);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d.chunk_count, sizeof(unsigned) * MAX_CHUNKS);
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&d.chunk_flag, sizeof(int) * MAX_CHUNKS, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err == cudaSuccess) err = cudaHostGetDevicePointer((void**)&d.chunk_flag_dev, d.chunk_flag, 0);
    if (err == cudaSuccess) err = cudaMalloc((void**)&d.in_flag, sizeof(int) * MAX_CHUNKS);
    if (err == cudaSuccess) err = cudaHostAlloc((void**)&d.ones, sizeof(int) * MAX_CHUNKS, cudaHostAllocPortable);
    if (err == cudaSuccess) for (int k = 0; k < MAX_CHUNKS; ++k) d.ones[k] = 1;
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&d.ev_ready, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&d.ev_done, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&d.ev_last, cudaEventDisableTiming);
    c->devs.push_back(d);       // pushed first so that a failed set-up is torn down by ivpb_destroy like everything else
    if (err != cudaSuccess) {
      const std::string msg = std::string("device setup: ") + cudaGetErrorString(err);
      ivpb_destroy(c);
      return fail(nullptr, IVPB_ERR_CUDA, msg);
    }
  }
  // peer access for the multi-device device-resident path (results gathered to device 0 over NVLink)
  for (size_t i = 1; i < c->devs.size(); ++i) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, c->devs[0].id, c->devs[i].id);
    if (can) {
      cudaSetDevice(c->devs[0].id); cudaDeviceEnablePeerAccess(c->devs[i].id, 0);
      cudaSetDevice(c->devs[i].id); cudaDeviceEnablePeerAccess(c->devs[0].id, 0);
      cudaGetLastError();
    }
  }
  cudaSetDevice(c->devs[0].id);
  *out = c;
  return 0;
}

void ivpb_destroy(ivpb_ctx* ctx) {
  if (!ctx) return;
  for (auto& d : ctx->devs) {
    cudaSetDevice(d.id);
    if (d.stream) cudaStreamSynchronize(d.stream);
    d.y0.release(); d.params.release(); d.t_eval.release(); d.tol_ext.release(); d.sparsity.release(); d.scratch.release(); d.rerun.release();
    d.q_traj.release(); d.q_ts.release(); d.q_y.release(); d.q_ok.release();
    d.sort_keys.release(); d.sort_vals.release(); d.sort_tmp.release(); d.sort_minmax.release();
    d.dense_nseg.release(); d.dense_segx.release(); d.dense_segc.release();
    for (auto& b : d.out) b.release();
    if (d.queue) cudaFree(d.queue);
    if (d.chunk_count) cudaFree(d.chunk_count);
    if (d.chunk_flag) cudaFreeHost(d.chunk_flag);
    if (d.in_flag) cudaFree(d.in_flag);
    if (d.ones) cudaFreeHost(d.ones);
    if (d.ev_ready) cudaEventDestroy(d.ev_ready);
    if (d.ev_done) cudaEventDestroy(d.ev_done);
    if (d.ev_last) cudaEventDestroy(d.ev_last);
    if (d.ev_shared) cudaEventDestroy(d.ev_shared);
    if (d.stream2) { cudaStreamSynchronize(d.stream2); cudaStreamDestroy(d.stream2); }
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  for (auto& u : ctx->user) ivpb_nvrtc_release(u);
  if (!ctx->devs.empty()) {
    cudaSetDevice(ctx->devs[0].id);
    ctx->pilot_in.release(); ctx->pilot_out[0].release(); ctx->pilot_out[1].release(); ctx->pilot_res.release();
  }
  delete ctx;
}

const char* ivpb_last_error(const ivpb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int ivpb_device_count(const ivpb_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
uint64_t ivpb_launch_count(const ivpb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int ivpb_last_fp_mode(const ivpb_ctx* ctx, int32_t info[6]) {
  if (!ctx) return -1;
  if (info) for (int i = 0; i < 6; ++i) info[i] = ctx->last_fp[i];
  return ctx->last_fp[0];
}
const char* ivpb_version(void) { return "ivp-b200 0.1.0 (sm_100a)"; }

int ivpb_builtin_problem(ivpb_ctx* ctx, int builtin_id, int* n, int* p, int* n_events) {
  if (builtin_id < 0 || builtin_id >= IVPB_P_BUILTIN_COUNT) return fail(ctx, IVPB_ERR_CONFIG, "unknown built-in problem id");
  ivpb_pinfo info;
  BUILTIN[builtin_id][0](-1, 0, &info);
  if (n) *n = info.n;
  if (p) *p = info.p;
  if (n_events) *n_events = info.nev;
  return 0;
}

int ivpb_nvrtc_problem(ivpb_ctx* ctx, const char* cuda_src, int n, int p, int n_events, int has_jac, int* handle) {
  if (!ctx || (!cuda_src && n != 0) || !handle) return fail(ctx, IVPB_ERR_CONFIG, "null argument");
  // n == 0 is the reference's empty state vector (solve_ivp.rs:148-176): nothing is compiled or integrated
  if (n < 0 || n > 1024) return fail(ctx, IVPB_ERR_CONFIG, "n must be in 0..1024 (n > 32 runs one trajectory per warp and needs ivp_ode_i)");
  if (p < 0 || n_events < 0 || n_events > ivpb::MAX_EVENTS_FN) return fail(ctx, IVPB_ERR_CONFIG, "bad p / n_events");
  ivpb_user_problem up;
  up.n = n; up.p = p; up.n_events = n_events; up.has_jac = has_jac; up.src = cuda_src ? cuda_src : "";
  ctx->user.push_back(up);
  *handle = IVPB_USER_HANDLE_BASE + (int)ctx->user.size() - 1;
  return 0;
}

void* ivpb_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void ivpb_host_free(void* p) { if (p) cudaFreeHost(p); }

int ivpb_solve_batch_device(ivpb_ctx* ctx, int problem, const ivpb_options* opt, int64_t N, double t0, double tf,
                            const double* d_y0, const double* d_params, const ivpb_outputs* d_out, void* stream) {
  if (!ctx) return IVPB_ERR_CONFIG;
  ProblemInfo pi;
  if (int rc = problem_info(ctx, problem, &pi)) return rc;
  if (int rc = validate(ctx, pi, opt, N, t0, tf)) return rc;
  if ((!d_y0 && pi.n > 0) || !d_out) return fail(ctx, IVPB_ERR_CONFIG, "null device buffer");
  if (pi.p > 0 && !d_params) return fail(ctx, IVPB_ERR_CONFIG, "params is null but the problem has parameters");
  Device& dev0 = ctx->devs[0];
  cudaStream_t s0 = (cudaStream_t)stream;
  const int G = (int)ctx->devs.size();
  CK(cudaSetDevice(dev0.id));
  // One solve in flight per context: this call's work on device 0 starts after the previous call's (another stream may
  // have been handed in), because the queue counter and the small staging buffers of the device are shared.
  CK(cudaStreamWaitEvent(s0, dev0.ev_last, 0));
  int fp_strict = 0;
  if (int rc = resolve_fp(ctx, problem, pi, opt, N, t0, tf, d_y0, d_params, true, s0, &fp_strict)) return rc;
  struct FpScope { ivpb_ctx* c; ~FpScope() { c->fp_override = -1; } } fp_scope{ctx};
  ctx->fp_override = fp_strict;
  if (G == 1) {
    const int rc = launch_shard(ctx, dev0, problem, pi, opt, N, t0, tf, d_y0, d_params, d_out, s0, false, 0, true);
    if (rc == 0) CK(cudaEventRecord(dev0.ev_last, s0));
    return rc;
  }

  // Multi-device, device-resident: inputs/outputs live on the first device.  Shard g > 0 receives its
  // slice of y0/params by a peer copy over NVLink, integrates it locally, and returns its results by peer
  // copies into the caller's arrays -- no collective, trajectories never interact.
  const int64_t cap = opt->has_t_eval ? (int64_t)opt->n_t_eval + 1 : (int64_t)opt->max_out;
  const int n = pi.n;
  size_t per[OUT_FIELDS];
  const int64_t seg_cap = opt->dense_output ? opt->max_segments : 0;
  field_bytes(n, pi.nev, cap, opt->max_events, seg_cap, coeffs_per_state(opt->method) * n, per);
  void* dst[OUT_FIELDS];
  out_to_array(d_out, dst);
  CK(cudaEventRecord(dev0.ev_ready, s0));          // inputs on device 0 are ready once s0 reaches here
  for (int g = 1; g < G; ++g) {
    Device& dev = ctx->devs[g];
    const int64_t lo = N * g / G, hi = N * (g + 1) / G, Ng = hi - lo;
    if (Ng == 0) continue;
    CK(cudaSetDevice(dev.id));
    CK(cudaStreamWaitEvent(dev.stream, dev0.ev_ready, 0));
    if (n > 0) {
      CK(dev.y0.ensure(sizeof(double) * n * Ng));
      CK(cudaMemcpyPeerAsync(dev.y0.p, dev.id, d_y0 + lo * n, dev0.id, sizeof(double) * n * Ng, dev.stream));
    }
    if (pi.p > 0) {
      CK(dev.params.ensure(sizeof(double) * pi.p * Ng));
      CK(cudaMemcpyPeerAsync(dev.params.p, dev.id, d_params + lo * pi.p, dev0.id, sizeof(double) * pi.p * Ng, dev.stream));
    }
    void* loc[OUT_FIELDS];
    for (int f = 0; f < OUT_FIELDS; ++f) {
      loc[f] = nullptr;
      if (!dst[f] || per[f] == 0) continue;
      CK(dev.out[f].ensure(per[f] * Ng));
      loc[f] = dev.out[f].p;
      if ((f >= OUT_NOUT && f <= OUT_EVY) || f == OUT_NSEG) CK(cudaMemsetAsync(loc[f], 0, per[f] * Ng, dev.stream));
    }
    ivpb_outputs d;
    array_to_out(loc, &d);
    if (int rc = launch_shard(ctx, dev, problem, pi, opt, Ng, t0, tf, (const double*)dev.y0.p,
                              pi.p > 0 ? (const double*)dev.params.p : nullptr, &d, dev.stream, false, 0, true))
      return rc;
    for (int f = 0; f < OUT_FIELDS; ++f)
      if (loc[f])
        CK(cudaMemcpyPeerAsync((char*)dst[f] + per[f] * lo, dev0.id, loc[f], dev.id, per[f] * Ng, dev.stream));
    CK(cudaEventRecord(dev.ev_done, dev.stream));
  }
  CK(cudaSetDevice(dev0.id));
  {
    const int64_t N0 = N / G;      // shard 0 works in place on the caller's arrays
    if (int rc = launch_shard(ctx, dev0, problem, pi, opt, N0, t0, tf, d_y0, d_params, d_out, s0, false, 0, true)) return rc;
  }
  for (int g = 1; g < G; ++g)
    if (N * (g + 1) / G - N * g / G > 0) CK(cudaStreamWaitEvent(s0, ctx->devs[g].ev_done, 0));
  CK(cudaEventRecord(dev0.ev_last, s0));
  return 0;
}

int ivpb_solve_batch(ivpb_ctx* ctx, int problem, const ivpb_options* opt, int64_t N, double t0, double tf,
                     const double* y0, const double* params, const ivpb_outputs* out) {
  if (!ctx) return IVPB_ERR_CONFIG;
  ProblemInfo pi;
  if (int rc = problem_info(ctx, problem, &pi)) return rc;
  if (int rc = validate(ctx, pi, opt, N, t0, tf)) return rc;
  if ((!y0 && pi.n > 0) || !out) return fail(ctx, IVPB_ERR_CONFIG, "null host buffer");
  if (pi.p > 0 && !params) return fail(ctx, IVPB_ERR_CONFIG, "params is null but the problem has parameters");
  const int G = (int)ctx->devs.size();
  const int64_t cap = opt->has_t_eval ? (int64_t)opt->n_t_eval + 1 : (int64_t)opt->max_out;
  const int n = pi.n, nev = pi.nev, me = opt->max_events;
  // bytes per trajectory of each output field, and the host base pointers
  size_t per[OUT_FIELDS];
  const int64_t seg_cap = opt->dense_output ? opt->max_segments : 0;
  const int n_cont = coeffs_per_state(opt->method) * n;
  field_bytes(n, nev, cap, me, seg_cap, n_cont, per);
  void* host[OUT_FIELDS];
  out_to_array(out, host);
  if (seg_cap > 0) {       // a new dense log replaces the retained one; solves without dense_output leave it alone
    ctx->dense.valid = false;
    ctx->dense.generation += 1;
    ctx->dense.method = opt->method; ctx->dense.n = n; ctx->dense.n_cont = n_cont; ctx->dense.cap = (int)seg_cap;
    ctx->dense.N = N; ctx->dense.lo.assign(G, 0); ctx->dense.count.assign(G, 0);
  }
  // ---- how the bytes cross PCIe -------------------------------------------------------------------------------------
  // Inputs.  A caller buffer that is page-locked host memory (ivpb_host_alloc, cudaHostAlloc/Register, torch pin_memory)
  // is mapped into the device address space under UVA; the kernel reads each trajectory's y0 / params row from it at
  // `init` (24 bytes per trajectory for the north star), so no H2D copy brackets the integration.  Pageable inputs
  // are staged with cudaMemcpyAsync.
  // Outputs.  Results are written to device memory and leave by bulk DMA WHILE the kernel integrates: the shard is one
  // persistent launch (the work queue keeps every lane busy to the end -- cutting it into several launches was measured
  // at 2x the device time for CR3BP, whose chunks were under two waves each), but it is divided into chunks of
  // consecutive trajectories with a completion counter each; the thread that retires a chunk's last trajectory raises
  // a flag in mapped host memory (chunk_signal, ivpb_erk.cuh), and this thread, polling the flags, enqueues that chunk's
  // device-to-host copies on a second stream.  Trajectories are handed out in index order, so chunks complete roughly
  // in order and only the last chunk's copy is exposed.  Round 1 instead let the kernel store results straight into
  // mapped host memory (IVPB_FLAG_ZEROCOPY_OUT keeps that route): every trajectory then issues four small PCIe writes,
  // which move at ~8 GB/s -- hidden under a 15 ms kernel, but the bottleneck of every short one (ball + events: 3.5 ms
  // on the device, 11.0 ms end to end) and of 8 ranks sharing one root complex (N = 8 e2e 18.3 ms against 15.2 ms).
  const bool ordered = (opt->flags & IVPB_FLAG_SORT) ||
                       ((opt->method == IVPB_RADAU || opt->method == IVPB_BDF) && !(opt->flags & IVPB_FLAG_NO_SORT));
  // the mapped (zero-copy) routes of round 1, inputs and outputs alike, are opt-in now
  const bool allow_zc = !(opt->flags & IVPB_FLAG_NO_ZEROCOPY) && !ordered && (opt->flags & IVPB_FLAG_ZEROCOPY_OUT);
  const bool zc_outputs = allow_zc;
  auto mapped = [&](const void* hp) -> char* {
    if (!allow_zc || !hp) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, hp) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? (char*)at.devicePointer : nullptr;
  };
  char* zc_y0 = mapped(y0);
  char* zc_par = pi.p > 0 ? mapped(params) : nullptr;
  char* zc_out[OUT_FIELDS];
  for (int f = 0; f < OUT_FIELDS; ++f) {
    // fields every trajectory writes in full (ErkTraj::finish); the sample blocks too, because `finish` zero-fills the
    // slots a trajectory left empty (KArgs::zero_tail)
    const bool whole = f <= OUT_NOUT || f == OUT_EVCOUNT || f == OUT_TOUT || f == OUT_YOUT;
    zc_out[f] = (whole && zc_outputs) ? mapped(host[f]) : nullptr;
  }
  const bool zero_interval = std::fabs(tf - t0) < 1e-15;       // that kernel writes only the matching samples: stage it
  if (zero_interval || n == 0 || (host[OUT_TOUT] && !zc_out[OUT_TOUT]) || (host[OUT_YOUT] && !zc_out[OUT_YOUT]))   // both or neither
    zc_out[OUT_TOUT] = zc_out[OUT_YOUT] = nullptr;
  size_t out_bytes_per_traj = 0;
  for (int f = 0; f < OUT_FIELDS; ++f) if (host[f] && !zc_out[f]) out_bytes_per_traj += per[f];
  int fp_strict = 0;
  if (int rc = resolve_fp(ctx, problem, pi, opt, N, t0, tf, y0, params, false, ctx->devs[0].stream, &fp_strict)) return rc;
  struct FpScope { ivpb_ctx* c; ~FpScope() { c->fp_override = -1; } } fp_scope{ctx};
  ctx->fp_override = fp_strict;
  // static contiguous split [g*N/G, (g+1)*N/G) -- trajectories are independent, no exchange step
  struct Shard {
    int64_t lo = 0, Ng = 0, chunk = 0;
    int C = 0, copied = 0;
    bool done[MAX_CHUNKS];
    char* dbase[OUT_FIELDS];
    bool direct[OUT_FIELDS];
  };
  std::vector<Shard> shards(G);
  for (int g = 0; g < G; ++g) {
    Device& dev = ctx->devs[g];
    Shard& S = shards[g];
    const int64_t lo = N * g / G, hi = N * (g + 1) / G, Ng = hi - lo;
    S.lo = lo; S.Ng = Ng;
    if (Ng == 0) continue;
    CK(cudaSetDevice(dev.id));
    if (g == 0) CK(cudaStreamWaitEvent(dev.stream, dev.ev_last, 0));    // after a device-buffer solve still in flight on a caller stream
    // chunks of >= 2 MB of results (smaller copies are latency-bound) and >= 4096 trajectories, at most MAX_CHUNKS
    int C = 1;
    if (!zero_interval && n > 0 && out_bytes_per_traj > 0 && !(opt->flags & IVPB_FLAG_NO_PIPELINE)) {
      const int64_t by_bytes = (int64_t)((out_bytes_per_traj * (size_t)Ng) >> 21), by_count = Ng >> 12;
      C = (int)std::max<int64_t>(1, std::min<int64_t>(MAX_CHUNKS, std::min(by_bytes, by_count)));
    }
    S.C = C; S.chunk = (Ng + C - 1) / C;
    S.C = (int)((Ng + S.chunk - 1) / S.chunk);
    for (int c = 0; c < MAX_CHUNKS; ++c) S.done[c] = false;
    // Inputs: staged in device memory.  Unless the shard is sorted first (locality order needs every row before the
    // launch), they arrive in chunks on the copy stream WHILE the kernel runs: the kernel starts after the first chunk
    // (1 MB) instead of after the whole shard, and the scheduler waits on a chunk's arrival flag before it reads a row.
    const double* d_y0 = zc_y0 ? (const double*)zc_y0 + lo * n : nullptr;
    const double* d_par = (pi.p > 0 && zc_par) ? (const double*)zc_par + lo * pi.p : nullptr;
    const bool stage_y0 = !d_y0 && n > 0, stage_par = pi.p > 0 && !d_par;
    const size_t in_bytes = (stage_y0 ? sizeof(double) * n : 0) + (stage_par ? sizeof(double) * pi.p : 0);
    int Cin = 0;
    if (in_bytes > 0 && !ordered && !zero_interval && !(opt->flags & IVPB_FLAG_NO_PIPELINE))
      Cin = (int)std::max<int64_t>(1, std::min<int64_t>(MAX_CHUNKS, std::min<int64_t>((int64_t)((in_bytes * (size_t)Ng) >> 20), Ng >> 12)));
    if (Cin < 2) Cin = 0;
    // chunk boundaries on multiples of 16 rows = whole 128-byte lines of every input array: a line fetched into L1 for the
    // last rows of one chunk must not contain rows of the next chunk that are still in flight
    const int64_t in_chunk = Cin ? (((Ng + Cin - 1) / Cin + 15) / 16) * 16 : 0;
    if (Cin) Cin = (int)((Ng + in_chunk - 1) / in_chunk);
    if (stage_y0) { CK(dev.y0.ensure(sizeof(double) * n * Ng)); d_y0 = (const double*)dev.y0.p; }
    if (stage_par) { CK(dev.params.ensure(sizeof(double) * pi.p * Ng)); d_par = (const double*)dev.params.p; }
    if (Cin) {
      CK(cudaMemsetAsync(dev.in_flag, 0, sizeof(int) * MAX_CHUNKS, dev.stream2));
      CK(cudaEventRecord(dev.ev_shared, dev.stream2));
      CK(cudaStreamWaitEvent(dev.stream, dev.ev_shared, 0));        // the kernel must not see flags of the previous call
    } else {
      if (stage_y0) CK(cudaMemcpyAsync(dev.y0.p, y0 + lo * n, sizeof(double) * n * Ng, cudaMemcpyHostToDevice, dev.stream));
      if (stage_par) CK(cudaMemcpyAsync(dev.params.p, params + lo * pi.p, sizeof(double) * pi.p * Ng, cudaMemcpyHostToDevice, dev.stream));
    }
    void* dptr[OUT_FIELDS];
    for (int f = 0; f < OUT_FIELDS; ++f) {
      S.dbase[f] = nullptr; S.direct[f] = false; dptr[f] = nullptr;
      const bool seg_field = f >= OUT_NSEG && seg_cap > 0;      // the dense log stays on the device even if
      if ((!host[f] && !seg_field) || per[f] == 0) continue;    // the caller wants no host copy of it
      if (zc_out[f]) { S.dbase[f] = zc_out[f] + per[f] * lo; S.direct[f] = true; dptr[f] = S.dbase[f]; continue; }
      Buf& b = seg_field ? (f == OUT_NSEG ? dev.dense_nseg : f == OUT_SEGX ? dev.dense_segx : dev.dense_segc) : dev.out[f];
      CK(b.ensure(per[f] * Ng));
      S.dbase[f] = (char*)b.p; dptr[f] = b.p;
    }
    ivpb_outputs d;
    array_to_out(dptr, &d);
    if (seg_cap > 0) {
      ctx->dense.lo[g] = lo; ctx->dense.count[g] = Ng;
      CK(cudaMemsetAsync(d.n_seg, 0, per[OUT_NSEG] * Ng, dev.stream));
    }
    // sample / event slots the kernel does not touch must read as zero on the host
    if (d.t_out && !S.direct[OUT_TOUT]) CK(cudaMemsetAsync(d.t_out, 0, per[OUT_TOUT] * Ng, dev.stream));
    if (d.y_out && !S.direct[OUT_YOUT]) CK(cudaMemsetAsync(d.y_out, 0, per[OUT_YOUT] * Ng, dev.stream));
    if (d.ev_t) CK(cudaMemsetAsync(d.ev_t, 0, per[OUT_EVT] * Ng, dev.stream));
    if (d.ev_y) CK(cudaMemsetAsync(d.ev_y, 0, per[OUT_EVY] * Ng, dev.stream));
    if (d.ev_count && !S.direct[OUT_EVCOUNT]) CK(cudaMemsetAsync(d.ev_count, 0, per[OUT_EVCOUNT] * Ng, dev.stream));
    if (d.n_out && !S.direct[OUT_NOUT]) CK(cudaMemsetAsync(d.n_out, 0, per[OUT_NOUT] * Ng, dev.stream));
    const bool flags_on = S.C > 1;
    if (flags_on) {
      for (int c = 0; c < S.C; ++c) dev.chunk_flag[c] = 0;
      CK(cudaMemsetAsync(dev.chunk_count, 0, sizeof(unsigned) * S.C, dev.stream));
    }
    if (int rc = launch_shard(ctx, dev, problem, pi, opt, Ng, t0, tf, d_y0, d_par, &d, dev.stream,
                              S.direct[OUT_TOUT] || S.direct[OUT_YOUT], 0, true, flags_on ? S.chunk : 0, in_chunk))
      return rc;
    CK(cudaEventRecord(dev.ev_done, dev.stream));       // "kernel finished": end of the polling loop below
    for (int c = 0; c < Cin; ++c) {                      // input chunks + their arrival flags, in index order, on the copy stream
      const int64_t clo = in_chunk * c, Nc = std::min<int64_t>(in_chunk, Ng - clo);
      if (Nc <= 0) break;
      if (stage_y0) CK(cudaMemcpyAsync((char*)dev.y0.p + sizeof(double) * n * clo, y0 + (lo + clo) * n, sizeof(double) * n * Nc, cudaMemcpyHostToDevice, dev.stream2));
      if (stage_par) CK(cudaMemcpyAsync((char*)dev.params.p + sizeof(double) * pi.p * clo, params + (lo + clo) * pi.p, sizeof(double) * pi.p * Nc, cudaMemcpyHostToDevice, dev.stream2));
      CK(cudaMemcpyAsync(dev.in_flag + c, dev.ones + c, sizeof(int), cudaMemcpyHostToDevice, dev.stream2));
    }
  }
  if (seg_cap > 0) ctx->dense.valid = true;
  // ---- drain: copy chunks out as their flags come up; what is left when a device's kernel has finished goes last ----
  auto copy_chunk = [&](int g, int c, cudaStream_t st) -> int {
    Shard& S = shards[g];
    const int64_t clo = S.chunk * c, Nc = std::min<int64_t>(S.chunk, S.Ng - clo);
    for (int f = 0; f < OUT_FIELDS; ++f) {
      if (!S.dbase[f] || !host[f] || S.direct[f]) continue;
      CK(cudaMemcpyAsync((char*)host[f] + per[f] * (S.lo + clo), S.dbase[f] + per[f] * clo, per[f] * Nc, cudaMemcpyDeviceToHost, st));
    }
    S.done[c] = true; S.copied += 1;
    return 0;
  };
  for (;;) {
    bool pending = false;
    for (int g = 0; g < G; ++g) {
      Shard& S = shards[g];
      if (S.Ng == 0 || S.copied == S.C) continue;
      Device& dev = ctx->devs[g];
      CK(cudaSetDevice(dev.id));
      const bool kernel_done = S.C == 1 || cudaEventQuery(dev.ev_done) == cudaSuccess;
      for (int c = 0; c < S.C; ++c) {
        if (S.done[c]) continue;
        if (kernel_done) { if (int rc = copy_chunk(g, c, dev.stream)) return rc; }
        else if (*((volatile int*)dev.chunk_flag + c) != 0) { if (int rc = copy_chunk(g, c, dev.stream2)) return rc; }
      }
      if (S.copied < S.C) pending = true;
    }
    if (!pending) break;
  }
  cudaGetLastError();      // cudaEventQuery's cudaErrorNotReady is not an error
  for (int g = 0; g < G; ++g) {
    if (shards[g].Ng == 0) continue;
    CK(cudaSetDevice(ctx->devs[g].id));
    CK(cudaStreamSynchronize(ctx->devs[g].stream));
    CK(cudaStreamSynchronize(ctx->devs[g].stream2));
  }
  CK(cudaSetDevice(ctx->devs[0].id));
  return 0;
}

static int dense_check(ivpb_ctx* ctx, uint64_t generation, int n, bool check_n) {
  const DenseLog& L = ctx->dense;
  if (!L.valid) return fail(ctx, IVPB_ERR_CONFIG, "no dense output retained: solve with dense_output = 1 first (InterpolationError::NotEnabled)");
  if (generation != 0 && generation != L.generation)
    return fail(ctx, IVPB_ERR_CONFIG, "this Solution's dense output is no longer retained: a later dense_output solve on the same context replaced it");
  if (check_n && n != L.n) return fail(ctx, IVPB_ERR_CONFIG, "state size does not match the retained dense output");
  return 0;
}

static int dense_eval_impl(ivpb_ctx* ctx, uint64_t generation, int n, int64_t n_query, const int64_t* traj, const double* ts,
                           double* y, int32_t* ok, int extrapolate) {
  if (!ctx) return IVPB_ERR_CONFIG;
  if (int rc = dense_check(ctx, generation, n, true)) return rc;
  const DenseLog& L = ctx->dense;
  if (n_query < 0 || (n_query > 0 && (!traj || !ts || !y || !ok))) return fail(ctx, IVPB_ERR_CONFIG, "null argument");
  for (int64_t q = 0; q < n_query; ++q)
    if (traj[q] < 0 || traj[q] >= L.N) return fail(ctx, IVPB_ERR_CONFIG, "trajectory index out of range");
  if (n_query == 0) return 0;
  const size_t M = (size_t)n_query;
  std::vector<double> yg;
  std::vector<int32_t> okg;
  for (size_t g = 0; g < ctx->devs.size(); ++g) {
    if (L.count[g] == 0) continue;
    Device& dev = ctx->devs[g];
    CK(cudaSetDevice(dev.id));
    Buf &q_traj = dev.q_traj, &q_ts = dev.q_ts, &q_y = dev.q_y, &q_ok = dev.q_ok;
    cudaError_t e = q_traj.ensure(8 * M);
    if (e == cudaSuccess) e = q_ts.ensure(8 * M);
    if (e == cudaSuccess) e = q_y.ensure(8 * M * L.n);
    if (e == cudaSuccess) e = q_ok.ensure(4 * M);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_traj.p, traj, 8 * M, cudaMemcpyHostToDevice, dev.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_ts.p, ts, 8 * M, cudaMemcpyHostToDevice, dev.stream);
    if (e == cudaSuccess)
      e = ivpb_launch_dense_eval(L.method, L.n, L.n_cont, L.cap, (const int*)dev.dense_nseg.p,
                                 (const double*)dev.dense_segx.p, (const double*)dev.dense_segc.p, (long long)M,
                                 (const long long*)q_traj.p, L.lo[g], L.count[g], (const double*)q_ts.p, (double*)q_y.p,
                                 (int*)q_ok.p, extrapolate, dev.stream);
    ctx->launches += 1;
    yg.resize(M * L.n); okg.resize(M);
    if (e == cudaSuccess) e = cudaMemcpyAsync(yg.data(), q_y.p, 8 * M * L.n, cudaMemcpyDeviceToHost, dev.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(okg.data(), q_ok.p, 4 * M, cudaMemcpyDeviceToHost, dev.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev.stream);
    if (e != cudaSuccess) return fail(ctx, IVPB_ERR_CUDA, std::string("ivpb_dense_eval: ") + cudaGetErrorString(e));
    for (size_t q = 0; q < M; ++q) {
      if (traj[q] < L.lo[g] || traj[q] >= L.lo[g] + L.count[g]) continue;
      ok[q] = okg[q];
      if (okg[q]) std::memcpy(y + q * L.n, yg.data() + q * L.n, 8 * (size_t)L.n);
    }
  }
  CK(cudaSetDevice(ctx->devs[0].id));
  return 0;
}

int ivpb_dense_eval(ivpb_ctx* ctx, uint64_t generation, int n, int64_t n_query, const int64_t* traj, const double* ts,
                    double* y, int32_t* ok) {
  return dense_eval_impl(ctx, generation, n, n_query, traj, ts, y, ok, 0);
}
int ivpb_dense_eval_extrapolate(ivpb_ctx* ctx, uint64_t generation, int n, int64_t n_query, const int64_t* traj,
                                const double* ts, double* y, int32_t* ok) {
  return dense_eval_impl(ctx, generation, n, n_query, traj, ts, y, ok, 1);
}
uint64_t ivpb_dense_generation(const ivpb_ctx* ctx) { return (ctx && ctx->dense.valid) ? ctx->dense.generation : 0; }

int ivpb_dense_span(ivpb_ctx* ctx, uint64_t generation, int64_t first, int64_t count, double* t_start, double* t_end,
                    int32_t* n_seg) {
  if (!ctx) return IVPB_ERR_CONFIG;
  if (int rc = dense_check(ctx, generation, 0, false)) return rc;
  const DenseLog& L = ctx->dense;
  if (first < 0 || count < 0 || first + count > L.N || !t_start || !t_end || !n_seg) return fail(ctx, IVPB_ERR_CONFIG, "bad range / null argument");
  for (size_t g = 0; g < ctx->devs.size(); ++g) {
    const int64_t a = std::max<int64_t>(first, L.lo[g]), b = std::min<int64_t>(first + count, L.lo[g] + L.count[g]);
    if (b <= a) continue;
    Device& dev = ctx->devs[g];
    CK(cudaSetDevice(dev.id));
    const size_t m = (size_t)(b - a);
    Buf t0b, t1b, nb;
    cudaError_t e = t0b.ensure(8 * m);
    if (e == cudaSuccess) e = t1b.ensure(8 * m);
    if (e == cudaSuccess) e = nb.ensure(4 * m);
    if (e == cudaSuccess)
      e = ivpb_launch_dense_span(L.cap, (const int*)dev.dense_nseg.p, (const double*)dev.dense_segx.p, a - L.lo[g],
                                 (long long)m, (double*)t0b.p, (double*)t1b.p, (int*)nb.p, dev.stream);
    ctx->launches += 1;
    if (e == cudaSuccess) e = cudaMemcpyAsync(t_start + (a - first), t0b.p, 8 * m, cudaMemcpyDeviceToHost, dev.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t_end + (a - first), t1b.p, 8 * m, cudaMemcpyDeviceToHost, dev.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_seg + (a - first), nb.p, 4 * m, cudaMemcpyDeviceToHost, dev.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev.stream);
    t0b.release(); t1b.release(); nb.release();
    if (e != cudaSuccess) return fail(ctx, IVPB_ERR_CUDA, std::string("ivpb_dense_span: ") + cudaGetErrorString(e));
  }
  CK(cudaSetDevice(ctx->devs[0].id));
  return 0;
}

int ivpb_measure_fp64_peak(ivpb_ctx* ctx, double* tflops) {
  if (!ctx || !tflops) return IVPB_ERR_CONFIG;
  Device& dev = ctx->devs[0];
  CK(cudaSetDevice(dev.id));
  const int block = 256, grid = dev.sms * 8, iters = 1 << 16;
  double* buf = nullptr;
  CK(cudaMalloc((void**)&buf, sizeof(double) * (size_t)block * grid));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, dev.stream));
    dfma_peak_kernel<<<grid, block, 0, dev.stream>>>(buf, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(e1, dev.stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * (double)block * (double)grid;
    if (rep > 0) best = std::fmax(best, fl / (ms * 1e-3) / 1e12);
    ctx->launches += 1;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  *tflops = best;
  return 0;
}

}  // extern "C"

// Debug hook (not part of include/ivpb.h): how many trajectories the last strict solve on device 0 handed to the guarded
// second pass (ivpb_exact.cuh, "deferred guards").  Synchronises the device.  -1: no two-pass solve has run.
extern "C" long long ivpb_debug_last_reruns(ivpb_ctx* ctx) {
  if (!ctx || ctx->devs.empty() || !ctx->devs[0].rerun.p) return -1;
  unsigned n = 0;
  cudaSetDevice(ctx->devs[0].id);
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  if (cudaMemcpy(&n, ctx->devs[0].rerun.p, sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return (long long)n;
}

// Debug hook (not part of include/ivpb.h): the runtime's column grouping of a jac_sparsity structure, so the CPU suite can
// compare it with the reference's greedy rule without a GPU.  Returns the number of groups.
extern "C" int ivpb_debug_group_columns(int n, const int32_t* colptr, const int32_t* rows, int32_t* groups) {
  std::vector<int32_t> g;
  const int ng = group_columns(n, colptr, rows, g);
  for (int c = 0; c < n; ++c) groups[c] = g[(size_t)c];
  return ng;
}

// ------------------------------------------------------------------------------------------------
// Debug hook (not part of include/ivpb.h): evaluates the fast-mode controller helpers of
// ivpb_fastmath.cuh on the device so tests can bound their error against libm.
#include "ivpb_fastmath.cuh"
#include "ivpb_libm_pow.cuh"
namespace {
__global__ void fastmath_kernel(const double* x, int n, double* r_rcp, double* r_rsqrt, double* r_rroot8) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  r_rcp[i] = ivpb::fm::rcp(x[i]);
  r_rsqrt[i] = ivpb::fm::rsqrt(x[i]);
  r_rroot8[i] = ivpb::fm::rroot8(x[i]);
}
}  // namespace
namespace {
__global__ void libm_pow_kernel(const double* x, const double* y, int n, double* r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = ivpb_libm_pow(x[i], y[i]);
}
}  // namespace
// Debug hook: ivpb_libm_pow (ivpb_libm_pow.cuh) evaluated on the device, so tests can compare it bit for bit
// with the host libm.
extern "C" int ivpb_debug_pow(const double* x, const double* y, int n, double* r) {
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * 3 * (size_t)n) != cudaSuccess) return IVPB_ERR_CUDA;
  cudaMemcpy(d, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice);
  libm_pow_kernel<<<(n + 127) / 128, 128>>>(d, d + n, n, d + 2 * (size_t)n);
  cudaError_t e = cudaMemcpy(r, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : IVPB_ERR_CUDA;
}

namespace {
__global__ void fastmath2_kernel(const double* x, int n, double* r_log2, double* r_exp2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  r_log2[i] = ivpb::fm::log2_fast(x[i]);
  r_exp2[i] = ivpb::fm::exp2_fast(x[i]);
}
}  // namespace
extern "C" int ivpb_debug_fastmath2(const double* x, int n, double* r_log2, double* r_exp2) {
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * 3 * (size_t)n) != cudaSuccess) return IVPB_ERR_CUDA;
  cudaMemcpy(d, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  fastmath2_kernel<<<(n + 127) / 128, 128>>>(d, n, d + n, d + 2 * (size_t)n);
  cudaMemcpy(r_log2, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(r_exp2, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : IVPB_ERR_CUDA;
}

extern "C" int ivpb_debug_fastmath(const double* x, int n, double* r_rcp, double* r_rsqrt, double* r_rroot8) {
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * 4 * (size_t)n) != cudaSuccess) return IVPB_ERR_CUDA;
  cudaMemcpy(d, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  fastmath_kernel<<<(n + 127) / 128, 128>>>(d, n, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n);
  cudaMemcpy(r_rcp, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(r_rsqrt, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaMemcpy(r_rroot8, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : IVPB_ERR_CUDA;
}

// Debug hook: ivpb::ex::div / ex::sqrt (ivpb_exact.cuh) next to the plain operators on the device, so a test can
// compare them bit for bit.  r_div_ex / r_div_ref = a / b, r_sqrt_ex / r_sqrt_ref = sqrt(a); the shared-reciprocal form
// is exercised by dividing a and a * 0.75 by the same refined reciprocal (r_div2_ex / r_div2_ref = 0.75 a / b).
#include "ivpb_exact.cuh"
namespace {
__global__ void exact_kernel(const double* a, const double* b, long long n, double* r_div_ex, double* r_div_ref,
                             double* r_div2_ex, double* r_div2_ref, double* r_sqrt_ex, double* r_sqrt_ref) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b[i];
  const ivpb::ex::Recip r = ivpb::ex::recip(y);
  r_div_ex[i] = ivpb::ex::div(x, r);
  r_div2_ex[i] = ivpb::ex::div(__dmul_rn(x, 0.75), r);
  r_div_ref[i] = __ddiv_rn(x, y);
  r_div2_ref[i] = __ddiv_rn(__dmul_rn(x, 0.75), y);
  r_sqrt_ex[i] = ivpb::ex::sqrt(x);
  r_sqrt_ref[i] = __dsqrt_rn(x);
}
// Fully on-device variant for large sweeps: operands from a counter-based generator, only the mismatch count returns.
__global__ void exact_sweep_kernel(unsigned long long seed, long long n, int mode, unsigned long long* mismatches) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto mix = [](unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  unsigned long long ua = mix(seed + 2ull * (unsigned long long)i), ub = mix(seed + 2ull * (unsigned long long)i + 1ull);
  if (mode == 0) {          // moderate exponents (what the solvers see): 2^-64 .. 2^64, random signs and mantissas
    ua = (ua & 0x800fffffffffffffull) | ((unsigned long long)(959 + (ua >> 52) % 128) << 52);
    ub = (ub & 0x800fffffffffffffull) | ((unsigned long long)(959 + (ub >> 52) % 128) << 52);
  } else if (mode == 2) {   // mantissas with long runs of ones / zeros (the hard cases of Markstein-type corrections)
    const int sa = (int)(ua >> 58) % 52, sb = (int)(ub >> 58) % 52;
    ua = (ua & 0x8000000000000000ull) | (1023ull << 52) | ((0x000fffffffffffffull >> sa) ^ ((ua >> 3) & 7ull));
    ub = (ub & 0x8000000000000000ull) | (1023ull << 52) | ((0x000fffffffffffffull << sb) & 0x000fffffffffffffull) | ((ub >> 3) & 7ull);
  }                         // mode 1: raw bit patterns (every class: zeros, subnormals, inf, NaN, huge)
  const double x = __longlong_as_double((long long)ua), y = __longlong_as_double((long long)ub);
  const double q = ivpb::ex::div(x, y), qr = __ddiv_rn(x, y);
  const double s = ivpb::ex::sqrt(x), sr = __dsqrt_rn(x);
  const bool dq = !(__double_as_longlong(q) == __double_as_longlong(qr) || (q != q && qr != qr));
  const bool ds = !(__double_as_longlong(s) == __double_as_longlong(sr) || (s != s && sr != sr));
  if (dq) atomicAdd(mismatches, 1ull);
  if (ds) atomicAdd(mismatches + 1, 1ull);
}
}  // namespace
extern "C" int ivpb_debug_exact(const double* a, const double* b, long long n, double* out6) {
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * 8 * (size_t)n) != cudaSuccess) return IVPB_ERR_CUDA;
  cudaMemcpy(d, a, sizeof(double) * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d + n, b, sizeof(double) * n, cudaMemcpyHostToDevice);
  exact_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d, d + n, n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, d + 6 * n, d + 7 * n);
  cudaError_t e = cudaMemcpy(out6, d + 2 * n, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : IVPB_ERR_CUDA;
}
extern "C" int ivpb_debug_exact_sweep(unsigned long long seed, long long n, int mode, unsigned long long* mismatches2) {
  unsigned long long* d = nullptr;
  if (cudaMalloc((void**)&d, 16) != cudaSuccess) return IVPB_ERR_CUDA;
  cudaMemset(d, 0, 16);
  exact_sweep_kernel<<<(unsigned)((n + 255) / 256), 256>>>(seed, n, mode, d);
  cudaError_t e = cudaMemcpy(mismatches2, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : IVPB_ERR_CUDA;
}
