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

// 1/x: MUFU.RCP64H seed (~2^-20) + two Newton steps
__device__ __forceinline__ double rcp(double x) {
  if (!in_range(x)) return 1.0 / x;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// 1/sqrt(x), x > 0: MUFU.RSQ64H seed + two Newton steps
__device__ __forceinline__ double rsqrt(double x) {
  if (!in_range(x) || x < 0.0) return ::rsqrt(x);
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  return fma(r, e, r);
}

// Comparison-select max/min: 3 instructions instead of the ~7 of IEEE fmax/fmin on doubles.  A NaN in `a`
// yields `b` (the reference's NaN-ignoring f64::max/min would do the same for its first operand).
__device__ __forceinline__ double maxsel(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double minsel(double a, double b) { return a < b ? a : b; }

// x^(-1/8) for x >= 0.  fp32 seed through MUFU lg2/ex2 on the argument clamped to [1e-30, 1e30] (rel. error
// ~2e-6), then two Newton steps on f(u) = u^-8 - x, u <- u + u (1 - x u^8) / 8, which need multiplications only
// (error 4.5 e^2 per step).  Below 1e-30 (and for 0) the iteration just grows u a little from the clamped seed
// 5.6e3; at or above 1e30, and for inf / NaN, the result is forced to 0.  Callers clamp the step factor to
// [0.333, 6] (or [0.2, 10]) with maxsel/minsel, so both ends land on the bound the reference's controller
// reaches for such an error (growth by scale_max for err -> 0, shrink by scale_min for err -> inf / NaN).
__device__ __forceinline__ double rroot8(double x) {
  const float xr = (float)x;
  float xf = fminf(fmaxf(xr, 1e-30f), 1e30f), lf, uf;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lf) : "f"(xf));
  lf *= -0.125f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(uf) : "f"(lf));
  double u = (double)uf;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double u2 = u * u, u4 = u2 * u2, u8 = u4 * u4;
    const double r = fma(-x, u8, 1.0);
    u = fma(0.125 * u, r, u);
  }
  return (xr < 1e30f) ? u : 0.0;
}

}  // namespace fm
}  // namespace ivpb
