// ivpb_fastmath.cuh -- branch-light fp64 special functions for the step-size controller (fast mode only).
//
// The reference computes, per attempted step, 2n divisions, a sqrt, a division and one or two libm `pow`s
// (dop853.rs:408-437, dopri5.rs:343-356, rk23.rs:229-305).  CUDA's IEEE-exact versions of those cost more
// issue slots than the 12 Runge-Kutta stages because each carries a slow-path check and call.  The
// helpers below are accurate to a few ulp (NOT correctly rounded), have no slow path inside their stated
// range, and fall back to the exact library function outside it.  They are compiled only into the
// default (FMA) kernels; the IVPB_FLAG_STRICT_FP kernels keep the reference's operations one for one.
#pragma once

namespace ivpb {
namespace fm {

// true when |x| is comfortably inside the normal range (exponent check on the high word)
__device__ __forceinline__ bool in_range(double x) {
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
  return hi - 0x03000000u < 0x7a000000u;      // roughly 1e-293 < |x| < 1e280
}

// Out-of-range arguments (zero, subnormal, huge, inf, NaN) take the exact library routines; one shared,
// non-inlined copy per kernel keeps the call sites small (the implicit kernels are instruction-cache bound).
static __device__ __noinline__ double rcp_slow(double x) { return 1.0 / x; }
static __device__ __noinline__ double rsqrt_slow(double x) { return ::rsqrt(x); }

// 1/x: MUFU.RCP64H seed r (relative error e ~ 2^-20), then one third-order step r (1 + e + e^2) with
// e = 1 - x r: remaining error e^3 ~ 2^-60, i.e. rounding-limited, in 3 dependent DFMAs.
__device__ __forceinline__ double rcp(double x) {
  if (!in_range(x)) return rcp_slow(x);
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// 1/sqrt(x), x > 0: MUFU.RSQ64H seed r (e = 1 - x r^2 ~ 2^-19), then one third-order step
// r (1 + e/2 + 3 e^2 / 8): remaining error ~ (5/16) e^3 ~ 2^-59.
__device__ __forceinline__ double rsqrt(double x) {
  if (!in_range(x) || x < 0.0) return rsqrt_slow(x);
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-(x * r), r, 1.0);
  return fma(r * e, fma(0.375, e, 0.5), r);
}

// Comparison-select max/min: 3 instructions instead of the ~7 of IEEE fmax/fmin on doubles.  A NaN in `a`
// yields `b` (the reference's NaN-ignoring f64::max/min would do the same for its first operand).
__device__ __forceinline__ double maxsel(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double minsel(double a, double b) { return a < b ? a : b; }

// x^(-1/8) for x >= 0.  fp32 seed u through MUFU lg2/ex2 on the argument clamped to [1e-30, 1e30] (relative error
// below ~1e-5), then ONE step of the binomial series of (1 - r)^(-1/8) with r = 1 - x u^8:
//   u (1 + r/8 + 9 r^2/128 + 51 r^3/1024),   truncation error ~ 0.04 r^4 < 1e-17 for |r| < 1e-4,
// which needs multiplications and FMAs only.  Below 1e-30 (and for 0) the step just grows u a little from the
// clamped seed 5.6e3; at or above 1e30, and for inf / NaN, the result is forced to 0.  Callers clamp the step
// factor to [0.333, 6] (or [0.2, 10]) with maxsel/minsel, so both ends land on the bound the reference's
// controller reaches for such an error (growth by scale_max for err -> 0, shrink by scale_min for err -> inf / NaN).
__device__ __forceinline__ double rroot8(double x) {
  const float xr = (float)x;
  float xf = fminf(fmaxf(xr, 1e-30f), 1e30f), lf, uf;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lf) : "f"(xf));
  lf *= -0.125f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(uf) : "f"(lf));
  const double u = (double)uf;
  const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
  const double r = fma(-x, u8, 1.0);
  const double pl = fma(fma(0.0498046875, r, 0.0703125), r, 0.125) * r;
  const double v = fma(u, pl, u);
  return (xr < 1e30f) ? v : 0.0;
}

// 1/x for arguments known to be comfortably normal (no range check)
__device__ __forceinline__ double rcp_nc(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

// Polynomial coefficients live in the constant bank: a literal double costs two UMOVs (one issue slot per 32-bit half) on
// sm_100a, a constant-bank entry one LDCU.64 -- or half of an LDCU.128 -- and the DOPRI5 / RK23 kernels are issue-bound
// (Lorenz DOPRI5 hot loop: 682 instructions, 304 of them on the fp64 pipe, 95 UMOVs before this change).
static __constant__ double LOG2_C[12] = {0.12545174268599682, 0.1373995277037108, 0.15186263588304877, 0.16972882833987804,
                                         0.19235933878519512, 0.22195308321368667, 0.2623081892525388, 0.3205988979753252,
                                         0.4121985831111324, 0.5770780163555853, 0.9617966939259756, 2.8853900817779268};
static __constant__ double EXP2_C[14] = {1.3691488853904128e-12, 2.5678435993488206e-11, 4.4455382718708116e-10,
                                         7.054911620801123e-09, 1.01780860092397e-07, 1.321548679014431e-06,
                                         1.5252733804059841e-05, 0.0001540353039338161, 0.0013333558146428443,
                                         0.009618129107628477, 0.05550410866482158, 0.24022650695910072, 0.6931471805599453, 1.0};
static __device__ __noinline__ double log2_slow(double x) { return ::log2(x); }
static __device__ __noinline__ double exp2_slow(double x) { return ::exp2(x); }

// log2(x), x > 0 normal: x = 2^e m with m in [sqrt(1/2), sqrt(2)), log2(m) = (2 / ln 2) atanh(s), s = (m - 1) / (m + 1),
// |s| <= 0.1716: twelve terms of the odd series (truncation 5e-20), branch-free; anything else -> library routine.
// Used by the DOPRI5 / PI controller of the default build, where CUDA's log + exp cost more than the six stages.
__device__ __forceinline__ double log2_fast(double x) {
  if (!in_range(x) || x <= 0.0) return log2_slow(x);
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, lo);
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double s = (m - 1.0) * rcp_nc(m + 1.0), s2 = s * s;
  double p = LOG2_C[0];
#pragma unroll
  for (int j = 1; j < 12; ++j) p = fma(p, s2, LOG2_C[j]);
  return fma(p, s, (double)e);
}

// 2^t for |t| < 1000: t = k + f, |f| <= 1/2, 2^f by fourteen terms of the exponential series (truncation 4e-18),
// scaled by constructing 2^k; anything else (inf, NaN, huge) -> library routine.
__device__ __forceinline__ double exp2_fast(double t) {
  if (!(fabs(t) < 1000.0)) return exp2_slow(t);
  const double kd0 = t + 6755399441055744.0;          // 1.5 * 2^52: the low word now holds rint(t)
  const int k = __double2loint(kd0);
  const double f = t - (kd0 - 6755399441055744.0);
  double p = EXP2_C[0];
#pragma unroll
  for (int j = 1; j < 14; ++j) p = fma(p, f, EXP2_C[j]);
  return p * __hiloint2double((k + 1023) << 20, 0);
}

}  // namespace fm
}  // namespace ivpb
