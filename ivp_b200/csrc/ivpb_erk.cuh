// ivpb_erk.cuh -- persistent explicit Runge-Kutta ensemble kernel for sm_100a (fp64).
//
// One trajectory per thread (n <= 32; ThreadLayout) or per warp (n > 32; WarpLayout) -- the step code is written
// once against a layout policy.  All stage vectors live in registers (fully unrolled, table-driven stage sums
// whose slot indices fold at compile time), the Butcher coefficients become constant-bank operands of DFMA, and
// every trajectory runs the reference's own controller:
//   DOP853  reference src/methods/dop853.rs:114-670      DOPRI5  src/methods/dopri5.rs:122-478
//   RK23    reference src/methods/rk23.rs:81-321         RK4     src/methods/rk4.rs:64-244
//   hinit   reference src/methods/mod.rs:217-281
// followed, when outputs are requested, by the device form of DefaultSolOut::solout
// (reference src/solve/solout.rs:128-431): t_eval sampling, step-mode capture, the first_step rule,
// event sign tests + Brent root finding on the step interpolant, terminal counts.
//
// Scheduling: a persistent grid; lanes whose trajectory finished pull the next index from a global
// atomic work queue (warp-aggregated), so step-count divergence between trajectories does not idle
// the SM.  `static_sched` keeps the plain one-thread-one-trajectory mapping for A/B measurements.
#pragma once
#include "ivpb_common.cuh"
#include "dop853_tableau.cuh"
#include "ivpb_fastmath.cuh"
#include "ivpb_libm_pow.cuh"
#include "ivpb_exact.cuh"

namespace ivpb {

enum { K_OUT = 1, K_EVENTS = 2, K_USER = 4 };
enum { K_NOPIPE = 0x100 };   // lookup flag (not a template feature bit): the twin compiled without arrival / completion flags   // kernel feature bits (template parameter FEAT); K_USER: the problem's own SolOut

// Multiply-add of the solver core.  Default build: an explicit fma, so the result does not depend on which
// products a particular compiler run chooses to contract -- the static, work-queue and NVRTC instances of a
// kernel agree bit for bit.  Strict build (-fmad=false): the reference's separate multiply and add.
#ifdef IVPB_STRICT
#define IVPB_MA(a, b, c) ((a) * (b) + (c))
#else
#define IVPB_MA(a, b, c) fma((a), (b), (c))
#endif

// DOPRI5 tableau values (reference src/methods/dopri5.rs:482-520; rational expressions evaluated in fp64
// at compile time exactly as the reference's consts).  Structure tables are constexpr in the step function.
static __constant__ double D5_S_COEF[6][5] = {
    {0.2, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
    {35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};
static __constant__ double D5_S_C[6] = {0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0};
static __constant__ double D5_E_COEF[6] = {71.0 / 57600.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0,
                                           22.0 / 525.0, -1.0 / 40.0};
static __constant__ double D5_D_COEF[6] = {-12715105075.0 / 11282082432.0, 87487479700.0 / 32700410799.0,
                                           -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,
                                           -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0};

template <int METHOD> struct MethodTraits;
template <> struct MethodTraits<M_RK23>   { static constexpr int NC = 4, IORD = 3; };
template <> struct MethodTraits<M_DOPRI5> { static constexpr int NC = 5, IORD = 5; };
template <> struct MethodTraits<M_DOP853> { static constexpr int NC = 8, IORD = 8; };
template <> struct MethodTraits<M_RK4>    { static constexpr int NC = 4, IORD = 4; };
template <> struct MethodTraits<M_RADAU>  { static constexpr int NC = 4, IORD = 5; };   // cont: radau.rs:697-705
template <> struct MethodTraits<M_BDF>    { static constexpr int NC = 7, IORD = 1; };   // D0..D5 + order marker

// ---------------------------------------------------------------------------------------------
// Step interpolants (Method::interpolate).  cont is coefficient-major: c[coef][state].
template <int METHOD, int N>
__device__ __forceinline__ void erk_interp(double xi, double* yi, const double (&c)[MethodTraits<METHOD>::NC][N],
                                           double xold, double h) {
  if constexpr (METHOD == M_DOP853) {          // dop853.rs:659-670
    const double s = ex::gdiv(xi - xold, h), s1 = 1.0 - s;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double conpar = IVPB_MA(s, IVPB_MA(s1, IVPB_MA(s, c[7][i], c[6][i]), c[5][i]), c[4][i]);
      yi[i] = IVPB_MA(s, IVPB_MA(s1, IVPB_MA(s, IVPB_MA(s1, conpar, c[3][i]), c[2][i]), c[1][i]), c[0][i]);
    }
  } else if constexpr (METHOD == M_DOPRI5) {   // dopri5.rs:467-478
    const double th = ex::gdiv(xi - xold, h), th1 = 1.0 - th;
#pragma unroll
    for (int i = 0; i < N; ++i)
      yi[i] = IVPB_MA(th, IVPB_MA(th1, IVPB_MA(th, IVPB_MA(th1, c[4][i], c[3][i]), c[2][i]), c[1][i]), c[0][i]);
  } else if constexpr (METHOD == M_RK23) {     // rk23.rs:313-321
    const double xc = ex::gdiv(xi - xold, h), x2 = xc * xc, x3 = x2 * xc;
#pragma unroll
    for (int i = 0; i < N; ++i)
      yi[i] = IVPB_MA(h, IVPB_MA(c[3][i], x3, IVPB_MA(c[2][i], x2, c[1][i] * xc)), c[0][i]);
  } else if constexpr (METHOD == M_RADAU) {    // radau.rs:798-809
    const double s = ex::gdiv(xi - (xold + h), h);
    const double C1M1 = -0.8449489742783178, C2M1 = -0.3550510257216822;
#pragma unroll
    for (int i = 0; i < N; ++i) yi[i] = c[0][i] + s * (c[1][i] + (s - C2M1) * (c[2][i] + (s - C1M1) * c[3][i]));
  } else if constexpr (METHOD == M_BDF) {      // bdf.rs:618-656 (device cont is coefficient-major; c[6][0] = order)
    if (h == 0.0) return;
    const int order = (int)fmin(fmax(round(c[6][0]), 1.0), 5.0);
    const double x_new = xold + h;
    double pk[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (k < order) {
        const double denom = h * ((double)k + 1.0);
        const double t_shift = x_new - h * (double)k;
        const double xf = (xi - t_shift) / denom;
        pk[k] = (k == 0) ? xf : pk[k > 0 ? k - 1 : 0] * xf;
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double sum = c[0][i];
#pragma unroll
      for (int k = 0; k < 5; ++k) if (k < order) sum += c[1 + k][i] * pk[k];
      yi[i] = sum;
    }
  } else {                                      // rk4.rs:229-244 (cubic Hermite, cont = [y_old, k4 stage, f_new, y_new])
    const double t = ex::gdiv(xi - xold, h), t2 = t * t, t3 = t2 * t;
    const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0, h10 = t3 - 2.0 * t2 + t;
    const double h01 = -2.0 * t3 + 3.0 * t2, h11 = t3 - t2;
#pragma unroll
    for (int i = 0; i < N; ++i)
      yi[i] = IVPB_MA(h11 * h, c[2][i], IVPB_MA(h01, c[3][i], IVPB_MA(h10 * h, c[1][i], h00 * c[0][i])));
  }
}


// ---------------------------------------------------------------------------------------------
// Layout policies: how a state vector is spread over threads.
//   ThreadLayout  one trajectory per thread, every vector a register array of N (n <= 32).
//   WarpLayout    one trajectory per warp (n > 32): component i lives in lane i % 32, slot i / 32; the RHS sees
//                 the full state through a per-warp shared-memory row, norms are reduced with shuffles (fast
//                 build) or summed in index order from shared memory (strict build: the reference's sequential
//                 `err += q * q`, bit for bit).  All control flow is warp-uniform.
// The step code below is written once against this interface; with ThreadLayout everything folds away.
template <class Prob>
struct ThreadLayout {
  static constexpr int N = Prob::N, NL = Prob::N;
  static constexpr bool WARP = false;
  static __device__ __forceinline__ int gi(int i) { return i; }
  static __device__ __forceinline__ bool valid(int) { return true; }
  static __device__ __forceinline__ bool leader() { return true; }
  static __device__ __forceinline__ double rtol(const KArgs& a, int i) { return a.rtol[i]; }
  static __device__ __forceinline__ double atol(const KArgs& a, int i) { return a.atol[i]; }
  static __device__ __forceinline__ double sumsq(const double (&q)[NL]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) acc = IVPB_MA(q[i], q[i], acc);
    return acc;
  }
  static __device__ __forceinline__ double sum(const double (&t)[NL]) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) acc += t[i];
    return acc;
  }
  // IVPB_NOINLINE_ODE: ONE out-of-line copy of the right-hand side per kernel instead of one per stage (DOP853: 12-15
  // sites).  The operands cross in local memory (2 n + p doubles per call, L1-resident); what it buys is code size: the
  // strict CR3BP DOP853 kernel is 140 KB of SASS against a 32 KB instruction cache.
  static __device__ __noinline__ void ode_call(double t, const double* y, const double* p, double* d) { Prob::ode(t, y, p, d); }
  static __device__ __forceinline__ void ode(double t, const double* y, const double* p, double* d) {
#ifdef IVPB_NOINLINE_ODE
    if constexpr (N >= IVPB_NOINLINE_ODE) {
      double yy[N], dd[N];
#pragma unroll
      for (int i = 0; i < N; ++i) yy[i] = y[i];
      ode_call(t, yy, p, dd);
#pragma unroll
      for (int i = 0; i < N; ++i) d[i] = dd[i];
    } else
#endif
    Prob::ode(t, y, p, d);
  }
  static __device__ __forceinline__ void events(double t, const double* y, const double* p, double* g) { Prob::events(t, y, p, g); }
};

template <class Prob, int EXTRA = 0>
struct WarpLayout {
  static constexpr int N = Prob::N, NL = (Prob::N + 31) / 32;
  static constexpr bool WARP = true;
  // full state row + reduction scratch (+ EXTRA doubles the implicit kernels keep per warp behind them)
  static constexpr int SMEM_DOUBLES_PER_WARP = 2 * Prob::N + EXTRA;
  static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
  static __device__ __forceinline__ int gi(int i) { return lane() + 32 * i; }
  static __device__ __forceinline__ bool valid(int i) { return gi(i) < N; }
  static __device__ __forceinline__ bool leader() { return lane() == 0; }
  static __device__ __forceinline__ double* row() {
    extern __shared__ double ivpb_smem[];
    return ivpb_smem + (threadIdx.x >> 5) * SMEM_DOUBLES_PER_WARP;
  }
  static __device__ __forceinline__ double rtol(const KArgs& a, int i) { return a.rtol_ext ? a.rtol_ext[valid(i) ? gi(i) : 0] : a.rtol[0]; }
  static __device__ __forceinline__ double atol(const KArgs& a, int i) { return a.atol_ext ? a.atol_ext[valid(i) ? gi(i) : 0] : a.atol[0]; }
  static __device__ __forceinline__ double sumsq(const double (&q)[NL]) {
#ifdef IVPB_STRICT
    double* sc = row() + N;
#pragma unroll
    for (int i = 0; i < NL; ++i) if (valid(i)) sc[gi(i)] = q[i] * q[i];
    __syncwarp();
    double acc = 0.0;
    for (int i = 0; i < N; ++i) acc = acc + sc[i];
    __syncwarp();
    return acc;
#else
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) if (valid(i)) acc = fma(q[i], q[i], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return acc;
#endif
  }
  // sum of per-component terms in index order (strict) / by shuffle reduction (fast)
  static __device__ __forceinline__ double sum(const double (&t)[NL]) {
#ifdef IVPB_STRICT
    double* sc = row() + N;
#pragma unroll
    for (int i = 0; i < NL; ++i) if (valid(i)) sc[gi(i)] = t[i];
    __syncwarp();
    double acc = 0.0;
    for (int i = 0; i < N; ++i) acc = acc + sc[i];
    __syncwarp();
    return acc;
#else
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) if (valid(i)) acc += t[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    return acc;
#endif
  }
  // component i of the RHS from the full state `ys`: the problem's ode_i, or -- for problems that only define
  // the whole-vector ode (n <= 32) -- the whole RHS evaluated by every lane (correct, but n-fold redundant)
  static __device__ __forceinline__ double ode_comp(double t, const double* ys, const double* p, int i) {
    if constexpr (Prob::HAS_ODE_I) return Prob::ode_i(t, ys, p, i);
    else {
      double tmp[N];
      Prob::ode(t, ys, p, tmp);
      double r = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) r = (k == i) ? tmp[k] : r;
      return r;
    }
  }
  // publish a distributed vector as a full row in shared memory
  static __device__ __forceinline__ const double* full(const double* y) {
    double* r = row();
#pragma unroll
    for (int i = 0; i < NL; ++i) if (valid(i)) r[gi(i)] = y[i];
    __syncwarp();
    return r;
  }
  static __device__ __forceinline__ void ode(double t, const double* y, const double* p, double* d) {
    const double* ys = full(y);
#pragma unroll
    for (int i = 0; i < NL; ++i) d[i] = valid(i) ? ode_comp(t, ys, p, gi(i)) : 0.0;
    __syncwarp();
  }
  static __device__ __forceinline__ void events(double t, const double* y, const double* p, double* g) {
    const double* ys = full(y);
    Prob::events(t, ys, p, g);
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------
// hinit -- reference src/methods/mod.rs:217-281 (initial step guess; one extra RHS call; the final
// min(|h|, 100|h|, h1, hmax) keeps the reference's extra |h| term, mod.rs:279)
template <class Prob, int IORD, class L = ThreadLayout<Prob>>
__device__ __forceinline__ double hinit_dev(const KArgs& a, double x, const double* y, const double* f0,
                                            const double* p, double posneg, double hmax, bool& gbad) {
  constexpr int N = L::NL;
  double qf[N], qy[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double sk = IVPB_MA(L::rtol(a, i), fabs(y[i]), L::atol(a, i));
    const ex::Recip rsk = ex::recip(sk);       // ex::div == `/` bit for bit (ivpb_exact.cuh), one reciprocal for both
    qf[i] = L::valid(i) ? ex::div(f0[i], rsk, gbad) : 0.0;
    qy[i] = L::valid(i) ? ex::div(y[i], rsk, gbad) : 0.0;
  }
  // interleaved in the reference (dnf, dny in one loop); the two sums are independent
  const double dnf = L::sumsq(qf), dny = L::sumsq(qy);
  double hh = (dnf <= 1e-10 || dny <= 1e-10) ? 1.0e-6 : ex::sqrt(ex::div(dny, dnf, gbad), gbad) * 0.01;
  if (hh > fabs(hmax)) hh = fabs(hmax);
  hh = fabs(hh) * signum(posneg);
  double y1[N], f1[N];
#pragma unroll
  for (int i = 0; i < N; ++i) y1[i] = IVPB_MA(hh, f0[i], y[i]);
  L::ode(x + hh, y1, p, f1);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double sk = IVPB_MA(L::rtol(a, i), fabs(y[i]), L::atol(a, i));
    qf[i] = L::valid(i) ? ex::div(f1[i] - f0[i], sk, gbad) : 0.0;
  }
  double der2 = L::sumsq(qf);
  der2 = ex::div(ex::sqrt(der2, gbad), fabs(hh), gbad);
  const double der12 = fmax(fabs(der2), ex::sqrt(dnf, gbad));
  const double h1 = (der12 <= 1.0e-15) ? fmax(1.0e-6, fabs(hh) * 1.0e-3) : ivpb_libm_pow(ex::div(0.01, der12, gbad), 1.0 / (double)IORD);
  const double hf = fmin(fmin(fmin(fabs(hh), 100.0 * fabs(hh)), h1), fabs(hmax));
  return fabs(hf) * signum(posneg);
}

// ---------------------------------------------------------------------------------------------
// Device DefaultSolOut (reference src/solve/solout.rs:15-432), state kept per thread.
template <class Prob, int METHOD, int FEAT, class L = ThreadLayout<Prob>>
struct SolOutDev {
  static constexpr int N = L::NL;          // local slice of the state (== Prob::N for ThreadLayout)
  static constexpr int NG = Prob::N;       // global state size
  static constexpr int NEV = (FEAT & K_EVENTS) ? Prob::NEV : 0;
  static constexpr int NEVS = NEV > 0 ? NEV : 1;
  static constexpr int NC = MethodTraits<METHOD>::NC;
  static constexpr double TOL = 1e-12;      // solout.rs:86

  int next_idx, n_out, n_seg;
  double last_t;                 // self.t.last()
  bool first_output_done, have_prev;
  double prev_g[NEVS];
  int hits[NEVS];
  double yold[(NEV > 0) ? N : 1];

  __device__ __forceinline__ void reset() {
    next_idx = 0; n_out = 0; n_seg = 0; last_t = 0.0; first_output_done = false; have_prev = false;
#pragma unroll
    for (int e = 0; e < NEVS; ++e) { prev_g[e] = 0.0; hits[e] = 0; }
  }

  __device__ __forceinline__ void push(const KArgs& a, i64 idx, double t, const double* yv) {
    if (n_out < a.out_cap) {
      const i64 o = idx * (i64)a.out_cap + n_out;
      if (a.t_out && L::leader()) a.t_out[o] = t;
      if (a.y_out) {
        // 8-byte stores on purpose: 16-byte ones (STG.128) measured the same on device memory and SLOWER when y_out is
        // mapped host memory (CR3BP, 101 samples: e2e 412 ms vs 392 ms; staged copies 425 ms)
#pragma unroll
        for (int i = 0; i < N; ++i) if (L::valid(i)) a.y_out[o * NG + L::gi(i)] = yv[i];
      }
    }
    last_t = t;
    ++n_out;
  }

  // Zero-copy sample output (KArgs::zero_tail): define the slots this trajectory did not fill.
  __device__ __forceinline__ void zero_tail(const KArgs& a, i64 idx) const {
    if (!a.zero_tail) return;
    for (int s = n_out; s < a.out_cap; ++s) {
      const i64 o = idx * (i64)a.out_cap + s;
      if (a.t_out && L::leader()) a.t_out[o] = 0.0;
      if (a.y_out) {
#pragma unroll
        for (int i = 0; i < N; ++i) if (L::valid(i)) a.y_out[o * NG + L::gi(i)] = 0.0;
      }
    }
  }

  static __device__ __forceinline__ bool crossed(double l, double r, int dir) {   // solout.rs:168-176
    if (dir == 0) return (l <= 0.0 && r >= 0.0) || (l >= 0.0 && r <= 0.0);
    if (dir > 0) return l < 0.0 && r >= 0.0;
    return l > 0.0 && r <= 0.0;
  }

  // Returns true for ControlFlag::Interrupt; then (tev, yev) hold the terminal event point.
  // `first` marks the initial callback (interpolant == None, xold == x).  (ixold, hstep) are the
  // StepInterpolant's own xold / h (src/dense.rs:32-97): BDF passes xold = x - h to solout but x_start to the
  // interpolant (bdf.rs:516-519); every other method passes the same value twice.
  __device__ __forceinline__ bool solout(const KArgs& a, i64 idx, const double* p, bool first, double xold, double x,
                                         const double* y, const double (&cont)[NC][N], double hstep, double ixold,
                                         double& tev, double* yev) {
    // Dense-output capture (solout.rs:141-146): every accepted step's (cont, xold, h), before the events.
    if (a.seg_cap > 0 && !first && x != xold && hstep != 0.0) {
      if (n_seg < a.seg_cap) {
        const i64 sg = idx * (i64)a.seg_cap + n_seg;
        if (L::leader()) { a.seg_x[2 * sg] = ixold; a.seg_x[2 * sg + 1] = hstep; }
        double* dst = a.seg_cont + sg * (i64)a.n_cont;
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int i = 0; i < N; ++i) {
            if (!L::valid(i)) continue;
            if constexpr (METHOD == M_BDF) dst[L::gi(i) * NC + c] = cont[c][i];      // state-major (bdf.rs:506-514)
            else dst[c * NG + L::gi(i)] = cont[c][i];                                // coefficient-major
          }
      }
      ++n_seg;
    }
    if constexpr (NEV > 0) {
      double g[NEV];
      L::events(x, y, p, g);
      if (!have_prev) {            // solout.rs:163-164 (yold.is_empty())
#pragma unroll
        for (int e = 0; e < NEV; ++e) prev_g[e] = g[e];
      } else {
        double det_t[NEV], det_y[NEV][N];
        int det_i[NEV], ndet = 0;
#pragma unroll
        for (int e = 0; e < NEV; ++e) {
          const double g_prev = prev_g[e], g_c = g[e];
          if (!crossed(g_prev, g_c, a.ev_dir[e])) continue;
          const double XTOL = 2e-12, RTOL = 2.220446049250313e-16;
          double aa = xold, b = x, fa = g_prev, fb = g_c;
          double et, ey[N];
          if (fabs(fa) <= XTOL) {
            et = aa;
#pragma unroll
            for (int i = 0; i < N; ++i) ey[i] = yold[i];
          } else if (fabs(fb) <= XTOL) {
            et = b;
#pragma unroll
            for (int i = 0; i < N; ++i) ey[i] = y[i];
          } else {                 // Brent, solout.rs:204-291
            double c = aa, fc = fa, d = b - aa, ee = d;
            double gm[NEV];
            for (int it = 0; it < 100; ++it) {
              if (fb * fc > 0.0) { c = aa; fc = fa; d = b - aa; ee = d; }
              if (fabs(fc) < fabs(fb)) { aa = b; b = c; c = aa; fa = fb; fb = fc; fc = fa; }
              const double tol1 = 2.0 * RTOL * fabs(b) + 0.5 * XTOL;
              const double xm = 0.5 * (c - b);
              if (fabs(xm) <= tol1 || fb == 0.0) break;
              if (fabs(ee) >= tol1 && fabs(fa) > fabs(fb)) {
                double s, pp, q;
                if (aa == c) {
                  s = fb / fa; pp = 2.0 * xm * s; q = 1.0 - s;
                } else {
                  const double qv = fa / fc, r = fb / fc;
                  s = fb / fa;
                  pp = s * (2.0 * xm * qv * (qv - r) - (b - aa) * (r - 1.0));
                  q = (qv - 1.0) * (r - 1.0) * (s - 1.0);
                }
                if (q > 0.0) pp = -pp; else q = -q;
                if (2.0 * pp < fmin(3.0 * xm * q - fabs(tol1 * q), fabs(ee * q))) { ee = d; d = pp / q; }
                else { d = xm; ee = d; }
              } else { d = xm; ee = d; }
              aa = b; fa = fb;
              if (fabs(d) > tol1) b += d;
              else b += (xm > 0.0 ? tol1 : -tol1);
              erk_interp<METHOD, N>(b, ey, cont, ixold, hstep);
              L::events(b, ey, p, gm);
              fb = gm[e];
            }
            erk_interp<METHOD, N>(b, ey, cont, ixold, hstep);
            et = b;
          }
          // append at position ndet (compile-time slot selected by predicate)
#pragma unroll
          for (int j = 0; j < NEV; ++j)
            if (j == ndet) {
              det_t[j] = et; det_i[j] = e;
#pragma unroll
              for (int i = 0; i < N; ++i) det_y[j][i] = ey[i];
            }
          ++ndet;
        }
        {  // stable sort by time, descending for backward integration (solout.rs:297-303)
          const bool forward = x > xold;
#pragma unroll
          for (int u = 1; u < NEV; ++u) {
#pragma unroll
            for (int v = u; v > 0; --v) {
              const bool sw = v < ndet && (forward ? (det_t[v] < det_t[v - 1]) : (det_t[v] > det_t[v - 1]));
              if (sw) {
                double tt = det_t[v]; det_t[v] = det_t[v - 1]; det_t[v - 1] = tt;
                int ti = det_i[v]; det_i[v] = det_i[v - 1]; det_i[v - 1] = ti;
#pragma unroll
                for (int i = 0; i < N; ++i) { double ty = det_y[v][i]; det_y[v][i] = det_y[v - 1][i]; det_y[v - 1][i] = ty; }
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < NEV; ++j) {
          if (j >= ndet) break;
          // det_i[j] is a run-time value: select the matching compile-time slot
#pragma unroll
          for (int e = 0; e < NEV; ++e) {
            if (det_i[j] != e) continue;
            if (hits[e] < a.max_events) {
              const i64 o = (idx * NEV + e) * (i64)a.max_events + hits[e];
              if (a.ev_t && L::leader()) a.ev_t[o] = det_t[j];
              if (a.ev_y) {
#pragma unroll
                for (int i = 0; i < N; ++i) if (L::valid(i)) a.ev_y[o * NG + L::gi(i)] = det_y[j][i];
              }
            }
            hits[e] += 1;
            if (a.ev_term[e] >= 0 && (i64)hits[e] >= a.ev_term[e]) {
              push(a, idx, det_t[j], det_y[j]);        // solout.rs:315-324
              tev = det_t[j];
#pragma unroll
              for (int i = 0; i < N; ++i) yev[i] = det_y[j][i];
#pragma unroll
              for (int q = 0; q < NEV; ++q) prev_g[q] = g[q];
              return true;
            }
          }
        }
#pragma unroll
        for (int e = 0; e < NEV; ++e) prev_g[e] = g[e];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) yold[i] = y[i];
      have_prev = true;
    }

    if constexpr ((FEAT & K_OUT) != 0) {
      if (a.n_t_eval >= 0) {       // Mode 1: t_eval (n_t_eval == -1 encodes Option::None)
        int i = next_idx;
        if (fabs(xold - x) <= TOL) {
          while (i < a.n_t_eval && fabs(a.t_eval[i] - x) <= TOL) { push(a, idx, a.t_eval[i], y); ++i; }
        } else {
          double yi[N];
          if (x > xold) {
            while (i < a.n_t_eval) {
              const double te = a.t_eval[i];
              if (!(te <= x + TOL)) break;
              if (te >= xold - TOL) { erk_interp<METHOD, N>(te, yi, cont, ixold, hstep); push(a, idx, te, yi); }
              ++i;
            }
          } else {
            while (i < a.n_t_eval) {
              const double te = a.t_eval[i];
              if (!(te >= x - TOL)) break;
              if (te <= xold + TOL) { erk_interp<METHOD, N>(te, yi, cont, ixold, hstep); push(a, idx, te, yi); }
              ++i;
            }
          }
        }
        next_idx = i;
      } else if (a.out_cap > 0) {  // Mode 2: accepted step endpoints
        if (a.has_first_step && !first_output_done && fabs(xold - x) > TOL) {   // solout.rs:392-421
          const double direction = signum(x - xold);
          const double target = a.t0 + direction * a.first_step;
          if (direction * (x - target) >= -TOL) {
            if (!first) {
              double yi[N];
              erk_interp<METHOD, N>(target, yi, cont, ixold, hstep);
              push(a, idx, target, yi);
              first_output_done = true;
            }
            if (fabs(x - target) > TOL) push(a, idx, x, y);
          }
          return false;
        }
        if (n_out == 0 || fabs(last_t - x) > TOL) push(a, idx, x, y);
      }
    }
    return false;
  }
};

// ---------------------------------------------------------------------------------------------
// The problem's own SolOut (src/solout.rs:55-63) in the warp-per-trajectory kernels (n > 32; RADAU / BDF n > 8).  The
// hook is executed by all 32 lanes with identical arguments (warp-uniform control flow): `y` is the FULL state in the
// warp's shared-memory row (n doubles, writable); `dense.eval(t, yi)` fills a full vector that must be shared by the warp
// -- `dense.buffer()` hands out n doubles for it (the thread-per-trajectory kernels offer the same call, so a hook
// written with it runs in both) -- each lane evaluating its own components; `emit(t, yv)` takes a full vector and
// every lane stores its slice.  Afterwards the lanes re-read their slices of y, so a moved state (ModifiedSolution, or the
// point an Interrupt reports) is picked up.
template <class Prob, int METHOD, class L, class Out>
struct WarpHook {
  static constexpr int N = L::NL, NG = Prob::N, NC = MethodTraits<METHOD>::NC;
  struct Interp {
    const double (&c)[NC][N]; double xold, h; bool ok;
    __device__ __forceinline__ bool valid() const { return ok; }
    __device__ __forceinline__ double* buffer() const { return L::row() + NG; }
    __device__ __forceinline__ void eval(double t, double* yi) const {
      double loc[N];
      erk_interp<METHOD, N>(t, loc, c, xold, h);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < N; ++i) if (L::valid(i)) yi[L::gi(i)] = loc[i];
      __syncwarp();
    }
  };
  struct Emit {
    Out& so; const KArgs& a; i64 idx;
    __device__ __forceinline__ void operator()(double t, const double* yv) {
      double loc[N];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < N; ++i) loc[i] = L::valid(i) ? yv[L::gi(i)] : 0.0;
      so.push(a, idx, t, loc);
    }
  };
  // returns the hook's flag: 0 Continue, 1 Interrupt, 2 ModifiedSolution
  static __device__ __forceinline__ int run(const KArgs& a, i64 idx, Out& so, bool first, double xold, double& x, double (&y)[N],
                                            const double* p, double* ustate, const double (&cont)[NC][N], double hstep, double ixold) {
    double* ys = L::row();
#pragma unroll
    for (int i = 0; i < N; ++i) if (L::valid(i)) ys[L::gi(i)] = y[i];
    __syncwarp();
    const Interp ip{cont, ixold, hstep, !first};
    Emit em{so, a, idx};
    const int fl = Prob::solout(xold, x, ys, p, ustate, ip, em);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = L::valid(i) ? ys[L::gi(i)] : 0.0;
    __syncwarp();
    return fl;
  }
};

// ---------------------------------------------------------------------------------------------
// Per-thread trajectory state + the step loops.
template <class Prob, int METHOD, int FEAT, class L = ThreadLayout<Prob>>
struct ErkTraj {
  static constexpr int N = L::NL, NG = Prob::N, P = Prob::P;     // N: local slice (== NG for ThreadLayout)
  static constexpr int PS = P > 0 ? P : 1;
  static constexpr bool DENSE = FEAT != 0;
  static constexpr bool BATCH_HEAVY = false;     // see run_schedule
  // Block-synchronous trips (run_schedule) for the kernels whose step code is far larger than the 32 KB instruction cache
  // and fetch-bound: the strict build (two instructions per multiply-add) of the wide tableaux / larger systems.
#if defined(IVPB_FORCE_BLOCK_SYNC)
  static constexpr bool BLOCK_SYNC = !L::WARP && (IVPB_FORCE_BLOCK_SYNC != 0);      // A/B builds (tools/build_variant.sh)
#elif defined(IVPB_STRICT)
  static constexpr bool BLOCK_SYNC = !L::WARP && (METHOD == M_DOP853 ? NG >= 3 : NG >= 5);
#else
  static constexpr bool BLOCK_SYNC = false;
#endif
  static constexpr int NC = MethodTraits<METHOD>::NC;
  using Out = SolOutDev<Prob, METHOD, FEAT, L>;

  i64 idx;
  double x, h;
  double y[N], k1[N], p[PS];
  double facold, hlamb;
  u32 nfev, nstep, naccpt, nrejct;
  int iasti, nonstiff, status;
  bool last, reject;
  Out so;
  static constexpr bool USER = (FEAT & K_USER) != 0;
  double ustate[USER ? Prob::NSTATE : 1];      // the user SolOut's own fields (Options.user_solout)
  bool gbad;      // deferred-guard flag of this trajectory (ivpb_exact.cuh); only the strictd kernels ever raise it

  __device__ __forceinline__ void to_event_point(double tev, const double* yev) {
    x = tev;
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = yev[i];
  }
  // What the user SolOut sees of the step (StepInterpolant, src/dense.rs:32-97) and of the output arrays
  struct UserInterp {
    const double (&c)[NC][N]; double xold, h; bool ok;
    mutable double buf[N];
    __device__ __forceinline__ bool valid() const { return ok; }
    __device__ __forceinline__ double* buffer() const { return buf; }      // n doubles for eval(), see WarpHook
    __device__ __forceinline__ void eval(double t, double* yi) const { erk_interp<METHOD, N>(t, yi, c, xold, h); }
  };
  struct UserEmit {
    Out& so; const KArgs& a; i64 idx;
    __device__ __forceinline__ void operator()(double t, const double* yv) { so.push(a, idx, t, yv); }
  };
  // The solver's callback slot (e.g. dop853.rs:246-268,596-624): DefaultSolOut, or the problem's own SolOut.
  // Returns 0 Continue, 1 Interrupt (status set, (x, y) at the point to report), 2 ModifiedSolution (k1 re-evaluated).
  __device__ __forceinline__ int callback(const KArgs& a, bool first, double xold, const double (&cont)[NC][N], double hstep) {
    if (gbad) { status = ST_RERUN; return 1; }      // strictd kernels: abandon before emitting (ivpb_exact.cuh)
    if constexpr (USER) {
      int fl;
      if constexpr (L::WARP) fl = WarpHook<Prob, METHOD, L, Out>::run(a, idx, so, first, xold, x, y, p, ustate, cont, hstep, xold);
      else {
        const UserInterp ip{cont, xold, hstep, !first};
        UserEmit em{so, a, idx};
        fl = Prob::solout(xold, x, y, p, ustate, ip, em);
      }
      if (fl == 1) { status = ST_INTERRUPT; return 1; }
      if (fl == 2) { L::ode(x, y, p, k1); nfev += 1; return 2; }
      return 0;
    } else {
      double tev, yev[N];
      if (so.solout(a, idx, p, first, xold, x, y, cont, hstep, xold, tev, yev)) { status = ST_INTERRUPT; to_event_point(tev, yev); return 1; }
      return 0;
    }
  }
  __device__ __forceinline__ double rt(const KArgs& a, int i) const { return L::rtol(a, i); }
  __device__ __forceinline__ double at(const KArgs& a, int i) const { return L::atol(a, i); }

  __device__ __forceinline__ double hinit(const KArgs& a, double posneg, double hmax) {
    return hinit_dev<Prob, MethodTraits<METHOD>::IORD, L>(a, x, y, k1, p, posneg, hmax, gbad);
  }

  __device__ __forceinline__ double hmax_of(const KArgs& a) const {
    if (!a.has_max_step) return fabs(a.tf - a.t0);
    if constexpr (METHOD == M_DOPRI5) return a.max_step;      // dopri5.rs:180 (no abs)
    return fabs(a.max_step);                                   // dop853.rs:172-175, rk23.rs:135
  }

  // Everything the reference does before its main loop; returns true if the trajectory is already done.
  __device__ __forceinline__ bool init(const KArgs& a, i64 index) {
    idx = index;
    gbad = false;
    x = a.t0;
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = 0.0;
    if constexpr (!L::WARP && (NG % 2 == 0)) {
      if (a.vec_io) {                       // one row = NG/2 aligned 16-byte words
        const double2* src = reinterpret_cast<const double2*>(a.y0 + index * NG);
#pragma unroll
        for (int i = 0; i < NG / 2; ++i) { const double2 v = __ldg(src + i); y[2 * i] = v.x; y[2 * i + 1] = v.y; }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = a.y0[index * NG + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = L::valid(i) ? a.y0[index * NG + L::gi(i)] : 0.0;
    }
    if constexpr (P > 0) {
#pragma unroll
      for (int i = 0; i < P; ++i) p[i] = a.params[index * P + i];
    }
#ifdef IVPB_STRICT
    facold = 1e-4;
#else
    facold = (METHOD == M_DOPRI5) ? -13.287712379549449 : 1e-4;   // fast DOPRI5 carries log2(facold), facold = 1e-4
#endif
    hlamb = 0.0;
    nfev = 0; nstep = 0; naccpt = 0; nrejct = 0;
    iasti = 0; nonstiff = 0; status = ST_SUCCESS;
    last = false; reject = false;
    so.reset();
    const double posneg = signum(a.tf - a.t0);
    L::ode(x, y, p, k1);
    if constexpr (METHOD == M_RK4) {
      h = a.has_first_step ? a.first_step : (a.tf - a.t0) / 100.0;    // solve_ivp.rs:185; rk4.rs:116 (not counted)
    } else {
      nfev = 1;
      if (a.has_first_step) h = fabs(a.first_step) * posneg;
      else { nfev = 2; h = hinit(a, posneg, hmax_of(a)); }
    }
    if constexpr (FEAT != 0) {
      double cont[NC][N];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) cont[c][i] = 0.0;
      if constexpr (USER) {
#pragma unroll
        for (int s = 0; s < Prob::NSTATE; ++s) ustate[s] = 0.0;
      }
      if (callback(a, true, x, cont, 0.0) == 1) return true;
    }
    return false;
  }

  // Write the per-trajectory results.  (x, y) is the integrator's last accepted point, or the event point
  // when a terminal event interrupted the integration (step() moves it there).
  __device__ __forceinline__ void finish(const KArgs& a) {
    if constexpr (FEAT != 0) so.zero_tail(a, idx);
    if (a.y_final) {
      if constexpr (!L::WARP && (NG % 2 == 0)) {
        if (a.vec_io) {
          double2* dst = reinterpret_cast<double2*>(a.y_final + idx * NG);
#pragma unroll
          for (int i = 0; i < NG / 2; ++i) dst[i] = make_double2(y[2 * i], y[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < N; ++i) a.y_final[idx * NG + i] = y[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) if (L::valid(i)) a.y_final[idx * NG + L::gi(i)] = y[i];
      }
    }
    if (!L::leader()) return;                 // per-trajectory scalars: one writer
    if (a.status) a.status[idx] = status;
    if (a.counters) {                         // 24-byte row as three 8-byte words
      uint2* c = reinterpret_cast<uint2*>(a.counters + idx * 6);
      c[0] = make_uint2(nfev, 0u); c[1] = make_uint2(0u, nstep); c[2] = make_uint2(naccpt, nrejct);
    }
    if (a.t_final) a.t_final[idx] = x;
    if (a.h_next) a.h_next[idx] = h;
    if (a.n_out) a.n_out[idx] = a.out_cap > 0 ? so.n_out : 0;
    if (a.seg_n) a.seg_n[idx] = so.n_seg;
    if constexpr (Out::NEV > 0) {
      if (a.ev_count) {
#pragma unroll
        for (int e = 0; e < Out::NEV; ++e) a.ev_count[idx * Out::NEV + e] = so.hits[e];
      }
    }
  }

  // One attempted step.  Returns true when the trajectory has ended (status set; the caller runs finish()).
  __device__ __forceinline__ bool step(const KArgs& a);
};

// Table-driven linear combination: acc = sum_j COEF[j] * k[SLOT[j]][i], left to right.
#define IVPB_LINCOMB(acc, LEN, SLOT, COEF, KARR, i)                               \
  double acc = (COEF)[0] * (KARR)[(SLOT)[0]][i];                                  \
  _Pragma("unroll") for (int j_ = 1; j_ < 9; ++j_) if (j_ < (LEN)) acc = IVPB_MA((COEF)[j_], (KARR)[(SLOT)[j_]][i], acc);

template <class Prob, int METHOD, int FEAT, class L>
__device__ __forceinline__ bool ErkTraj<Prob, METHOD, FEAT, L>::step(const KArgs& a) {
  const double xend = a.tf;
  const double posneg = signum(a.tf - a.t0);
  double tev = 0.0, yev[N];
  bool term = false;

  if constexpr (METHOD == M_DOP853) {
    // ---- reference src/methods/dop853.rs:272-653 ----
    IVPB_DOP853_TABLES
    const double uround = 2.3e-16, safe = 0.9;
    const double facc1 = 1.0 / 0.333, facc2 = 1.0 / 6.0;      // beta = 0 => facold^beta == 1 exactly
    const double h_max = hmax_of(a);
    if ((u64)nstep > a.max_steps) { status = ST_NMAX; return true; }
#ifdef IVPB_STRICT
    if (0.1 * fabs(h) <= fabs(x) * uround) { status = ST_SMALL; return true; }
    if ((IVPB_MA(1.01, h, x) - xend) * posneg > 0.0) { h = xend - x; last = true; }
#else
    // same two tests with fewer fp64 operations: (a - xend) * (+-1) > 0 is a comparison of a with xend, and the
    // factor 0.1 moves into the constant
    if (fabs(h) <= fabs(x) * (10.0 * uround)) { status = ST_SMALL; return true; }
    {
      const double xa = fma(1.01, h, x);
      if (posneg > 0.0 ? (xa > xend) : (xa < xend)) { h = xend - x; last = true; }
    }
#endif
    nstep += 1;

    double k[10][N], y1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) k[0][i] = k1[i];
#pragma unroll
    for (int s = 0; s < 11; ++s) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (STG_LEN[s] == 1) {
          y1[i] = IVPB_MA(h * D853_STG_COEF[s][0], k[STG_SLOT[s][0]][i], y[i]);   // (h*a21)*k1, dop853.rs:296
        } else {
          IVPB_LINCOMB(acc, STG_LEN[s], STG_SLOT[s], D853_STG_COEF[s], k, i)
          y1[i] = IVPB_MA(h, acc, y[i]);
        }
      }
      const double ts = (s == 10) ? (x + h) : IVPB_MA(D853_STG_C[s], h, x);
      L::ode(ts, y1, p, k[STG_OUT[s]]);
    }
    const double xph = x + h;
    nfev += 11;
    // k4 = sum b_j k_j ; k5 = y + h k4      (slots 3 and 4)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      IVPB_LINCOMB(acc, LIN_LEN[0], LIN_SLOT[0], D853_LIN_COEF[0], k, i)
      k[3][i] = acc;
      k[4][i] = IVPB_MA(h, acc, y[i]);
    }
    double qe1[N], qe2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#ifdef IVPB_STRICT
      const double sk = IVPB_MA(rt(a, i), fmax(fabs(y[i]), fabs(k[4][i])), at(a, i));
#else
      const double sk = IVPB_MA(rt(a, i), fm::maxsel(fabs(y[i]), fabs(k[4][i])), at(a, i));
#endif
      // k4 - bh1 k1 - bh2 k9 - bh3 k3 (a - b*c == (-b)*c + a exactly)
      const double erri = IVPB_MA(-D853_BHH[2], k[2][i], IVPB_MA(-D853_BHH[1], k[8][i], IVPB_MA(-D853_BHH[0], k[0][i], k[3][i])));
      IVPB_LINCOMB(e8, LIN_LEN[1], LIN_SLOT[1], D853_LIN_COEF[1], k, i)
#ifdef IVPB_STRICT
      const ex::Recip rsk = ex::recip(sk);               // dop853.rs:412,423: two correctly rounded quotients, one reciprocal
      const double q2 = ex::div(erri, rsk, gbad), q1 = ex::div(e8, rsk, gbad);
#else
      const double rsk = fm::rcp(sk);                      // one reciprocal shared by both norms
      const double q2 = erri * rsk, q1 = e8 * rsk;
#endif
      qe2[i] = L::valid(i) ? q2 : 0.0;
      qe1[i] = L::valid(i) ? q1 : 0.0;
    }
    double err2 = L::sumsq(qe2), err = L::sumsq(qe1);
    double deno = IVPB_MA(0.01, err2, err);
    if (deno <= 0.0) deno = 1.0;
#ifdef IVPB_STRICT
    err = fabs(h) * err * ex::sqrt(ex::div(1.0, (double)NG * deno, gbad), gbad);
    const double fac11 = ivpb_libm_pow(err, 0.125);                     // expo1 = 1/8 - beta*0.2, beta = 0
    // facold^beta == 1 exactly (beta = 0), so fac = fac11 (dop853.rs:434)
    const double fmin1 = fmin(facc1, ex::div(fac11, safe, gbad));
    const double fac = fmax(facc2, fmin1);
    double hnew = ex::div(h, fac, gbad);
    // dop853.rs:645: the rejected step is h / min(facc1, fac11 / safe).  Same quotient whenever the lower clamp is idle
    // (always after a rejection: err > 1 => fac11 >= 1 > facc2 * safe), so the second division is almost never executed.
    const double hrej = (fac == fmin1) ? hnew : ex::div(h, fmin1, gbad);
#else
    // Same controller written with the reciprocal step factor 1/fac = safe * err^(-1/8), which needs
    // multiplications only (see ivpb_fastmath.cuh): hnew = h * clamp(safe/fac11, 1/facc1, 1/facc2).
    err = fabs(h) * err * fm::rsqrt((double)NG * deno);
    const double ratio = safe * fm::rroot8(err);
    double hnew = h * fm::minsel(fm::maxsel(ratio, 0.333), 6.0);
    const double hrej = h * fm::maxsel(ratio, 0.333);
    (void)facc1; (void)facc2;
#endif

    if (err <= 1.0) {
      facold = fmax(err, 1.0e-4);
      naccpt += 1;
      L::ode(xph, k[4], p, k[3]);
      nfev += 1;
      if ((naccpt % 1000u == 0u) || (iasti > 0)) {            // stiffness detection, dop853.rs:447-472
        double sd1[N], sd2[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { sd1[i] = k[3][i] - k[2][i]; sd2[i] = k[4][i] - y1[i]; }
        const double stnum = L::sumsq(sd1), stden = L::sumsq(sd2);
        if (stden > 0.0) hlamb = fabs(h) * ex::sqrt(ex::div(stnum, stden, gbad), gbad);
        if (hlamb > 6.1) {
          nonstiff = 0; iasti += 1;
          if (iasti == 15) { status = ST_STIFF; return true; }
        } else {
          nonstiff += 1;
          if (nonstiff == 6) iasti = 0;
        }
      }
      // The reference always builds the dense coefficients (3 extra RHS calls).  They cannot influence
      // y, h or err, so FEAT == 0 skips the arithmetic but still counts the evaluations (dop853.rs:560).
      nfev += 3;
      double cont[NC][N];
      if constexpr (DENSE) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          cont[0][i] = y[i];
          const double ydiff = k[4][i] - y[i];
          cont[1][i] = ydiff;
          const double bspl = IVPB_MA(h, k[0][i], -ydiff);
          cont[2][i] = bspl;
          cont[3][i] = IVPB_MA(-h, k[3][i], ydiff) - bspl;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            IVPB_LINCOMB(acc, DF_LEN[r], DF_SLOT[r], D853_DF_COEF[r], k, i)
            cont[4 + r][i] = acc;
          }
        }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            IVPB_LINCOMB(acc, DSTG_LEN[s], DSTG_SLOT[s], D853_DSTG_COEF[s], k, i)
            y1[i] = IVPB_MA(h, acc, y[i]);
          }
          L::ode(IVPB_MA(D853_DSTG_C[s], h, x), y1, p, k[DSTG_OUT[s]]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            double acc = cont[4 + r][i];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = IVPB_MA(D853_DS_COEF[r][j], k[DS_SLOT[r][j]][i], acc);
            cont[4 + r][i] = h * acc;
          }
        }
      }
      const double xold = x;
#pragma unroll
      for (int i = 0; i < N; ++i) { k1[i] = k[3][i]; y[i] = k[4][i]; }
      x = xph;
      if constexpr (FEAT != 0) {
        if (callback(a, false, xold, cont, h) == 1) return true;
      }
      if (last) { h = hnew; status = ST_SUCCESS; return true; }
      if (fabs(hnew) > fabs(h_max)) hnew = posneg * fabs(h_max);
      if (reject) { hnew = posneg * fmin(fabs(hnew), fabs(h)); reject = false; }
    } else {
      hnew = hrej;
      reject = true;
      if (naccpt > 1u) nrejct += 1;
      last = false;
    }
    h = hnew;
    (void)term;
    return false;

  } else if constexpr (METHOD == M_DOPRI5) {
    // ---- reference src/methods/dopri5.rs:266-461 ----
    constexpr int S_LEN[6] = {1, 2, 3, 4, 5, 5};
    constexpr int S_SLOT[6][5] = {{0, 0, 0, 0, 0}, {0, 1, 0, 0, 0}, {0, 1, 2, 0, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 4}, {0, 2, 3, 4, 5}};
    constexpr int S_OUT[6] = {1, 2, 3, 4, 5, 1};
    constexpr int ED_SLOT[6] = {0, 2, 3, 4, 5, 1};            // e / d rows over {k1,k3,k4,k5,k6,k2(new)}
    const double uround = 2.3e-16, safe = 0.9, beta = 0.04;
    const double facc1 = 1.0 / 0.2, facc2 = 1.0 / 10.0;
    const double expo1 = 0.2 - beta * 0.75;
    const double h_max = hmax_of(a);
    if ((u64)nstep > a.max_steps) { status = ST_NMAX; return true; }
#ifdef IVPB_STRICT
    if (0.1 * fabs(h) <= fabs(x) * uround) { status = ST_SMALL; return true; }
    if ((IVPB_MA(1.01, h, x) - xend) * posneg > 0.0) { h = xend - x; last = true; }
#else
    // same two tests with fewer fp64 operations: (a - xend) * (+-1) > 0 is a comparison of a with xend, and the
    // factor 0.1 moves into the constant
    if (fabs(h) <= fabs(x) * (10.0 * uround)) { status = ST_SMALL; return true; }
    {
      const double xa = fma(1.01, h, x);
      if (posneg > 0.0 ? (xa > xend) : (xa < xend)) { h = xend - x; last = true; }
    }
#endif
    nstep += 1;

    double k[6][N], y1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) k[0][i] = k1[i];
    const double xph = x + h;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (S_LEN[s] == 1) {
          y1[i] = IVPB_MA(h * D5_S_COEF[s][0], k[S_SLOT[s][0]][i], y[i]);
        } else {
          double acc = D5_S_COEF[s][0] * k[S_SLOT[s][0]][i];
#pragma unroll
          for (int j = 1; j < 5; ++j) if (j < S_LEN[s]) acc = IVPB_MA(D5_S_COEF[s][j], k[S_SLOT[s][j]][i], acc);
          y1[i] = IVPB_MA(h, acc, y[i]);
        }
      }
      const double ts = (s >= 4) ? xph : IVPB_MA(D5_S_C[s], h, x);
      L::ode(ts, y1, p, k[S_OUT[s]]);
    }
    nfev += 6;
    double cont[NC][N];
    if constexpr (DENSE) {                                       // dopri5.rs:328-334
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = D5_D_COEF[0] * k[ED_SLOT[0]][i];
#pragma unroll
        for (int j = 1; j < 6; ++j) acc = IVPB_MA(D5_D_COEF[j], k[ED_SLOT[j]][i], acc);
        cont[4][i] = h * acc;
      }
    }
    double qe[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {                                // k4 <- scaled error vector, dopri5.rs:337-340
      double acc = D5_E_COEF[0] * k[ED_SLOT[0]][i];
#pragma unroll
      for (int j = 1; j < 6; ++j) acc = IVPB_MA(D5_E_COEF[j], k[ED_SLOT[j]][i], acc);
      k[3][i] = acc * h;
#ifdef IVPB_STRICT
      const double sk = at(a, i) + rt(a, i) * fmax(fabs(y[i]), fabs(y1[i]));
      const double q = ex::div(k[3][i], sk, gbad);
#else
      const double sk = fma(rt(a, i), fm::maxsel(fabs(y[i]), fabs(y1[i])), at(a, i));
      const double q = k[3][i] * fm::rcp(sk);
#endif
      qe[i] = L::valid(i) ? q : 0.0;
    }
    double err = L::sumsq(qe);
#ifdef IVPB_STRICT
    err = ex::sqrt(ex::div(err, (double)NG, gbad), gbad);
    const double fac11 = ivpb_libm_pow(err, expo1);
    double fac = ex::div(fac11, ivpb_libm_pow(facold, beta), gbad);
    fac = fmax(facc2, fmin(facc1, ex::div(fac, safe, gbad)));
    double hnew = ex::div(h, fac, gbad);
    const bool accept = err <= 1.0;
#else
    // Same PI controller in log2 space: err = sqrt(e2), so log2(err) = log2(e2)/2 needs no square root, and
    // log2(facold) is carried from the previous accepted step (facold holds log2(max(err, 1e-4)) in this build).
    const double e2 = err * (1.0 / (double)NG);
    const double lerr = 0.5 * fm::log2_fast(e2);
    double hnew = h * fm::minsel(fm::maxsel(safe * fm::exp2_fast(fma(beta, facold, -(expo1 * lerr))), 0.2), 10.0);
    const bool accept = e2 <= 1.0;
    (void)facc1; (void)facc2;
#endif

    if (accept) {
#ifdef IVPB_STRICT
      facold = fmax(err, 1.0e-4);
#else
      facold = fmax(lerr, -13.287712379549449);             // log2(1e-4)
#endif
      naccpt += 1;
      if ((naccpt % 1000u == 0u) || (iasti > 0)) {              // dopri5.rs:364-391 (uses overwritten k4: quirk kept)
        double sd1[N], sd2[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          sd1[i] = k[1][i] - k[5][i];
          double sacc = D5_S_COEF[4][0] * k[0][i];
#pragma unroll
          for (int j = 1; j < 5; ++j) sacc = IVPB_MA(D5_S_COEF[4][j], k[j][i], sacc);
          const double ysti = IVPB_MA(h, sacc, y[i]);
          sd2[i] = y1[i] - ysti;
        }
        const double stnum = L::sumsq(sd1), stden = L::sumsq(sd2);
        if (stden > 0.0) hlamb = fabs(h) * ex::sqrt(ex::div(stnum, stden, gbad), gbad);
        if (hlamb > 3.25) {
          nonstiff = 0; iasti += 1;
          if (iasti == 15) { status = ST_STIFF; return true; }
        } else {
          nonstiff += 1;
          if (nonstiff == 6) iasti = 0;
        }
      }
      if constexpr (DENSE) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double ydiff = y1[i] - y[i];
          const double bspl = IVPB_MA(h, k[0][i], -ydiff);
          cont[0][i] = y[i];
          cont[1][i] = ydiff;
          cont[2][i] = bspl;
          cont[3][i] = IVPB_MA(-h, k[1][i], ydiff) - bspl;
        }
      }
      const double xold = x;
#pragma unroll
      for (int i = 0; i < N; ++i) { k1[i] = k[1][i]; y[i] = y1[i]; }
      x = xph;
      if constexpr (FEAT != 0) {
        if (callback(a, false, xold, cont, h) == 1) return true;
      }
      if (last) { h = hnew; status = ST_SUCCESS; return true; }
      if (fabs(hnew) > fabs(h_max)) hnew = posneg * fabs(h_max);
      if (reject) { hnew = posneg * fmin(fabs(hnew), fabs(h)); reject = false; }
    } else {
#ifdef IVPB_STRICT
      hnew = ex::div(h, fmin(facc1, ex::div(fac11, safe, gbad)), gbad);
#else
      hnew = h * fm::maxsel(safe * fm::exp2_fast(-expo1 * lerr), 0.2);
#endif
      reject = true;
      if (naccpt > 1u) nrejct += 1;
      last = false;
    }
    h = hnew;
    return false;

  } else if constexpr (METHOD == M_RK23) {
    // ---- reference src/methods/rk23.rs:189-307 ----
    constexpr double b1 = 2.0 / 9.0, b2 = 1.0 / 3.0, b3 = 4.0 / 9.0;
    constexpr double e1 = 5.0 / 72.0, e2 = -1.0 / 12.0, e3 = -1.0 / 9.0, e4 = 1.0 / 8.0;
    constexpr double d21 = -4.0 / 3.0, d22 = 1.0, d23 = 4.0 / 3.0, d24 = -1.0;
    constexpr double d31 = 5.0 / 9.0, d32 = -2.0 / 3.0, d33 = -8.0 / 9.0, d34 = 1.0;
    const double safe = 0.9, scale_min = 0.2, scale_max = 10.0, expo = -1.0 / 3.0;
    const double hmax = hmax_of(a);
    if ((u64)nstep >= a.max_steps) { status = ST_NMAX; return true; }
    if ((x + h - xend) * posneg > 0.0) h = xend - x;
    double k2[N], k3[N], k4[N], yt[N];
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h * 0.5, k1[i], y[i]);
    L::ode(IVPB_MA(0.5, h, x), yt, p, k2);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h * 0.75, k2[i], y[i]);
    L::ode(IVPB_MA(0.75, h, x), yt, p, k3);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h, IVPB_MA(b3, k3[i], IVPB_MA(b2, k2[i], b1 * k1[i])), y[i]);
    L::ode(x + h, yt, p, k4);
    nfev += 3;
    double qe[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double ye = h * IVPB_MA(e4, k4[i], IVPB_MA(e3, k3[i], IVPB_MA(e2, k2[i], e1 * k1[i])));
      const double tol = IVPB_MA(rt(a, i), fmax(fabs(yt[i]), fabs(y[i])), at(a, i));
#ifdef IVPB_STRICT
      const double q = ex::div(ye, tol, gbad);
#else
      const double q = ye * fm::rcp(tol);
#endif
      qe[i] = L::valid(i) ? q : 0.0;
    }
    double err = L::sumsq(qe);
#ifdef IVPB_STRICT
    err = ex::sqrt(ex::div(err, (double)NG, gbad), gbad);
    const double sfac = safe * ivpb_libm_pow(err, expo);                 // 0.9 * err^(-1/3), rk23.rs:289,303
    const bool accept = err <= 1.0;
#else
    const double errsq = err * (1.0 / (double)NG);              // err^2; err^(-1/3) = cbrt(1/sqrt(err^2))
    const double sfac = safe * cbrt(fm::rsqrt(errsq));
    const bool accept = errsq <= 1.0;
    (void)expo;
#endif
    if (accept) {
      nstep += 1; naccpt += 1;
      const double xold = x;
      double cont[NC][N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if constexpr (DENSE) {
          cont[0][i] = y[i];
          cont[1][i] = k1[i];
          cont[2][i] = IVPB_MA(d24, k4[i], IVPB_MA(d23, k3[i], IVPB_MA(d22, k2[i], d21 * k1[i])));
          cont[3][i] = IVPB_MA(d34, k4[i], IVPB_MA(d33, k3[i], IVPB_MA(d32, k2[i], d31 * k1[i])));
        }
        y[i] = yt[i];
      }
      x += h;
      int fl = 0;
      if constexpr (FEAT != 0) {
        fl = callback(a, false, xold, cont, h);
        if (fl == 1) return true;
      }
      if (fl != 2) {                     // rk23.rs:270-284: ModifiedSolution re-evaluated k1, every other flag reuses k4
#pragma unroll
        for (int i = 0; i < N; ++i) k1[i] = k4[i];
      }
      h *= fmax(fmin(sfac, scale_max), scale_min);
      if (fabs(h) > hmax) h = hmax * posneg;
      if (x == xend) { status = ST_SUCCESS; return true; }
    } else {
      nrejct += 1;
      h *= fmax(fmin(sfac, 1.0), scale_min);
    }
    return false;

  } else {
    // ---- RK4: reference src/methods/rk4.rs:141-222 ----
    if ((u64)nstep >= a.max_steps) { status = ST_NMAX; return true; }
    const bool lst = (IVPB_MA(1.01, h, x) - xend) * signum(h) > 0.0;
    double k2[N], k3[N], k4[N], yt[N];
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h * 0.5, k1[i], y[i]);
    L::ode(IVPB_MA(0.5, h, x), yt, p, k2);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h * 0.5, k2[i], y[i]);
    L::ode(IVPB_MA(0.5, h, x), yt, p, k3);
#pragma unroll
    for (int i = 0; i < N; ++i) yt[i] = IVPB_MA(h * 1.0, k3[i], y[i]);
    L::ode(IVPB_MA(1.0, h, x), yt, p, k4);
    const double xold = x;
    double cont[NC][N];
    x += h;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if constexpr (DENSE) { cont[0][i] = y[i]; cont[1][i] = k4[i]; }
      y[i] = IVPB_MA(h, IVPB_MA(1.0 / 6.0, k4[i], IVPB_MA(1.0 / 3.0, k3[i], IVPB_MA(1.0 / 3.0, k2[i], (1.0 / 6.0) * k1[i]))), y[i]);
    }
    L::ode(x, y, p, k1);
    nfev += 4;
    nstep += 1;
    if constexpr (DENSE) {
#pragma unroll
      for (int i = 0; i < N; ++i) { cont[2][i] = k1[i]; cont[3][i] = y[i]; }
    }
    if constexpr (FEAT != 0) {
      if (callback(a, false, xold, cont, h) == 1) return true;
    }
    if (lst) { status = ST_SUCCESS; return true; }
    return false;
  }
}

// ---------------------------------------------------------------------------------------------
#ifndef IVPB_BLOCK
#define IVPB_BLOCK 128
#endif

#ifndef IVPB_BATCH_NUM
#define IVPB_BATCH_NUM 1      // heavy lanes wait until NUM/DEN of the active lanes are heavy; measured on the BDF
#define IVPB_BATCH_DEN 1      // ensembles: 1/2 -> 89 ms, 3/4 -> 80 ms, 1/1 -> 62 ms (Robertson, 2^18 trajectories)
#endif
// Arrival flags (KArgs::in_*): block until the copy engine has delivered the input rows of trajectory `idx`.  The flag is
// written by a DMA that is stream-ordered after the chunk's data, so a set flag means the rows are in device memory; the
// rows are read for the first time after this point (nothing stale can sit in L1).
__device__ __forceinline__ void wait_input(const KArgs& a, i64 idx) {
  if (a.in_chunk <= 0) return;                         // uniform over the grid
  const volatile int* f = a.in_flag + idx / a.in_chunk;
  while (*f == 0) __nanosleep(200);
  __threadfence();
}

// Completion flags (KArgs::chunk_*): called by all lanes of a converged warp after `finish`; `fin` marks the lanes whose
// trajectory `idx` has just written its results.  Lanes retiring trajectories of the same chunk are counted with one
// atomic; every result store is fenced before the count that covers it, so when the count is complete the chunk's
// results are visible device-wide (the "last block" pattern), and the flag store tells the host it may copy them.
static __device__ __noinline__ void chunk_signal(const KArgs& a, bool fin, i64 idx) {
#ifdef IVPB_NO_CHUNK_SIGNAL
  return;
#endif
  if (a.chunk_size <= 0) return;                       // uniform over the grid
  const unsigned m = __ballot_sync(0xffffffffu, fin);
  if (!fin) return;
  const int c = (int)(idx / a.chunk_size);
  const unsigned peers = __match_any_sync(m, c);
  __threadfence();
  __syncwarp(peers);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
    const unsigned n = (unsigned)__popc(peers);
    const unsigned old = atomicAdd(a.chunk_count + c, n);
    const i64 lo = (i64)c * a.chunk_size;
    const i64 size = (lo + a.chunk_size <= a.N ? a.chunk_size : a.N - lo);
    if ((i64)old + (i64)n == size) {
      __threadfence_system();
      *((volatile int*)a.chunk_flag + c) = 1;
    }
  }
}
// one trajectory per warp: every lane has written its slice of the results
__device__ __forceinline__ void chunk_signal_warp(const KArgs& a, i64 idx) {
  if (a.chunk_size <= 0) return;
  __threadfence();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    const int c = (int)(idx / a.chunk_size);
    const unsigned old = atomicAdd(a.chunk_count + c, 1u);
    const i64 lo = (i64)c * a.chunk_size;
    const i64 size = (lo + a.chunk_size <= a.N ? a.chunk_size : a.N - lo);
    if ((i64)old + 1 == size) {
      __threadfence_system();
      *((volatile int*)a.chunk_flag + c) = 1;
    }
  }
}

// Generic persistent scheduler: Traj provides init(a, idx) -> done, step(a) -> done, finish(a).
//
// Traj::BLOCK_SYNC (block-synchronous trips, a compile-time property of the kernel): the warps of a block start every
// attempted step together (__syncthreads_or doubles as the loop test).  The step code of the larger systems is far bigger than the
// instruction caches (CR3BP DOP853 with dense output: 50 KB hot loop in the FMA build, 110 KB in the strict one, against a
// 32 KB L1.5 I$), so warps that drift apart each stream the whole loop from L2 on their own (ncu: "no instruction"
// is the top stall, 5.4 per issue in the strict kernel); in lock step one warp's fetch serves all of them.  Measured on
// CR3BP DOP853 + 101 samples, 2^18 trajectories, strict build: 384 ms -> 313 ms with 128-thread blocks, 238 ms with one
// 256-thread block per SM (all eight resident warps in step).  It does nothing for code that fits the cache or is not
// fetch-bound (FMA build of the same kernel 103.5 -> 104.4 ms; north star, Lorenz, RADAU / BDF: 0.5-15 % slower), hence a
// per-kernel switch.  Results are unaffected: only the interleaving of independent trajectories changes.
//
// PIPE = false compiles the arrival / completion flags out altogether: the device-resident entry point
// (ivpb_solve_batch_device) never uses them, and even the uniform early-outs cost the north-star kernel 3 % (15.0 -> 15.46 ms
// per 2^20 trajectories, measured A/B on one box).
// A trajectory has ended: write its results -- or, in a strictd kernel, put it on the re-run list if one of its divisions /
// square roots left the fast-path range (ivpb_exact.cuh).  Returns true when the results were written.
template <class Traj>
__device__ __forceinline__ bool retire(const KArgs& a, Traj& T) {
#ifdef IVPB_DEFER_GUARDS
  // a.debug_rerun_mod > 0 (IVPB_DEBUG_RERUN=k): hand every k-th trajectory to the second pass, to test it
  if (T.status == ST_RERUN || T.gbad || (a.debug_rerun_mod > 0 && T.idx % a.debug_rerun_mod == 0)) {
    const unsigned k = atomicAdd(a.rerun_count, 1u);
    a.rerun_list[k] = (unsigned)T.idx;
    return false;
  }
#endif
  T.finish(a);
  return true;
}

template <class Traj, bool PIPE = true>
__device__ __forceinline__ void run_schedule(const KArgs& a) {
  Traj T;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  constexpr bool bsync = Traj::BLOCK_SYNC;
  bool active = false, exhausted = false;
  // Second pass of a strictd launch (guarded strict build only): the trajectories are the ones on the first pass's re-run
  // list (KArgs::n_dev, perm).  Every other build reads N straight from the constant bank -- a register for it costs the
  // north-star kernel, which sits at its 72-register cap, 17 % (15.0 -> 17.5 ms).
#if defined(IVPB_STRICT) && !defined(IVPB_DEFER_GUARDS)
  const i64 NQ = a.n_dev ? (i64)*a.n_dev : a.N;
#else
#define NQ a.N
#endif
  for (;;) {
    // ---- refill: lanes without a trajectory pull the next index from the global queue (one atomic per warp).
    // The static schedule is the same loop with a "queue" that hands every thread exactly its own index, so
    // both schedules execute one and the same copy of the step code (bit-identical results, half the code).
    const unsigned need = __ballot_sync(FULL, !active);
    if (need && !exhausted) {
      i64 idx;
      if (a.static_sched) {
        idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
        exhausted = true;
      } else {
        const int leader = __ffs(need) - 1;
        u64 base = 0;
        if (lane == leader) base = atomicAdd(a.queue, (u64)__popc(need));
        base = __shfl_sync(FULL, base, leader);
        idx = (i64)base + __popc(need & ((1u << lane) - 1u));
        if ((i64)base + __popc(need) >= NQ) exhausted = true;
      }
      bool fin0 = false;
      if (!active && idx < NQ) {
        active = true;
        const i64 tidx = a.perm ? (i64)a.perm[idx] : idx;
        if constexpr (PIPE) wait_input(a, tidx);
        if (T.init(a, tidx)) { fin0 = retire(a, T); active = false; }
      }
      if constexpr (PIPE) chunk_signal(a, fin0, T.idx);
    }
    if constexpr (bsync) {
      // block-uniform decisions: leave when no lane of the block has work and every warp has seen the end of the queue
      if (!__syncthreads_or(active ? 1 : 0)) {
        if (!__syncthreads_or(exhausted ? 0 : 1)) break;
        continue;
      }
    } else {
      if (__ballot_sync(FULL, active) == 0u) {
        if (exhausted) break;
        continue;
      }
    }
    // ---- hot loop: every lane that owns a trajectory attempts steps until one of them finishes.  Nothing of
    // the refill logic is live in here.
    bool done = false;
    do {
      bool run = active;
      if constexpr (Traj::BATCH_HEAVY) {
        // Trajectories whose next trip starts with rarely needed, expensive work (BDF: order selection, change_d,
        // Jacobian, refactorisation) wait until all of the warp's active lanes are in the same situation, so that code
        // runs once for many lanes instead of on every trip for a few.  Each lane gets there within order + 2 trips,
        // so nobody waits long; results are unaffected (only the interleaving of independent lanes changes).
        const bool hv = active && T.heavy();
        const int na = __popc(__ballot_sync(FULL, active)), nh = __popc(__ballot_sync(FULL, hv));
        run = active && (!hv || nh * IVPB_BATCH_DEN >= na * IVPB_BATCH_NUM);
      }
      if (run) done = T.step(a);
#ifdef IVPB_DEFER_GUARDS
      if (run && T.gbad) done = true;      // abandon at once: garbage must not keep a trajectory stepping
#endif
    } while (!(bsync ? (__syncthreads_or(done ? 1 : 0) != 0) : (__any_sync(FULL, done) != 0)));
    bool fin = false;
    if (done) { fin = retire(a, T); active = false; }
    if constexpr (PIPE) chunk_signal(a, fin, T.idx);
  }
#ifdef NQ
#undef NQ
#endif
}

template <class Prob, int METHOD, int FEAT, bool PIPE = true>
__device__ __forceinline__ void erk_body(const KArgs& a) {
  run_schedule<ErkTraj<Prob, METHOD, FEAT>, PIPE>(a);
}

// Warp-per-trajectory scheduler: every warp owns one trajectory at a time and pulls the next index from the
// same global queue (one atomic per trajectory, by the leader lane); `static_sched` maps warp w of the grid to
// trajectory w.  Control flow is warp-uniform throughout.
template <class Traj>
__device__ __forceinline__ void run_schedule_warp(const KArgs& a) {
  Traj T;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const i64 warp_global = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  bool first = true;
  for (;;) {
    i64 idx;
    if (a.static_sched) {
      if (!first) break;
      idx = warp_global;
    } else {
      u64 base = 0;
      if (lane == 0) base = atomicAdd(a.queue, 1ull);
      idx = (i64)__shfl_sync(FULL, base, 0);
    }
    first = false;
    if (idx >= a.N) break;
    const i64 tidx = a.perm ? (i64)a.perm[idx] : idx;
    wait_input(a, tidx);
    if (!T.init(a, tidx)) {
      while (!T.step(a)) {}
    }
    T.finish(a);
    chunk_signal_warp(a, T.idx);
  }
}

template <class Prob, int METHOD, int FEAT>
__device__ __forceinline__ void erk_warp_body(const KArgs& a) {
  run_schedule_warp<ErkTraj<Prob, METHOD, FEAT, WarpLayout<Prob>>>(a);
}

}  // namespace ivpb
