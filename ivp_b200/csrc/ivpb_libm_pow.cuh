// ivpb_libm_pow.cuh -- glibc's pow, operation for operation, for the strict (-fmad=false) kernels.
//
// The reference's controllers call f64::powf == the host libm's pow (dop853.rs:429-434, dopri5.rs:348-353,
// rk23.rs:234-303, radau.rs:478-719, bdf.rs:413-575, methods/mod.rs:276).  CUDA's pow is accurate to 1-2 ulp
// but not bit-identical to glibc's (they disagree in the last bit for roughly one call in five), and for
// stiff / ill-conditioned ensembles a single differing ulp in h is enough to flip a Newton-convergence or
// accept/reject decision a few hundred steps later (measured: RADAU on Van der Pol mu=1000 keeps only 77 % of
// the step counts with CUDA's pow).  glibc >= 2.28 uses the table-driven algorithm of ARM's optimized-routines
// (sysdeps/ieee754/dbl-64/e_pow.c): log(x) as hi+lo from a 128-entry table and a degree-7 polynomial,
// exp(y*log x) from a 128-entry 2^(i/128) table and a degree-5 polynomial, every step a plain IEEE fp64
// operation.  The sequence below is the x86-64 FMA variant (__pow_fma, the one the ifunc resolver picks on
// every CPU with FMA+AVX2, i.e. on the GPU hosts), transcribed from its machine code so that fused and unfused
// operations match: the result is bit-identical for positive normal x and 2^-65 <= |y| < 2^63 with
// 2^-54 <= |y log x| < 512; anything else (zeros, infinities, NaNs, negative or subnormal x, overflow /
// underflow range) is delegated to the platform pow, where the results are exact or saturated anyway.
// Tables: ivpb_libm_pow_tables.cuh (generated from the installed libm by tools/extract_glibc_pow_tables.py).
//
// PROVENANCE / LICENCE: the operation sequence below was read off the machine code of glibc 2.39's libm (x86-64,
// __pow_fma); glibc is LGPL-2.1-or-later (its pow is derived from ARM optimized-routines, MIT).  This file and the
// generated tables are therefore LGPL-derived material and must be distributed under terms compatible with that licence.
// Strict-build parity is parity with a reference linked against THIS libm variant ("Floating-point mode").
#pragma once

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define IVPB_LIBM_TABLE static __device__ const
// The polynomial / reduction constants (compile-time indices only) come from the constant bank: as `__device__ const` the
// compiler folds them into immediates, and every 64-bit immediate costs two moves per use (IVPB_PLAIN_POW_HEAD: A/B switch)
#ifdef IVPB_PLAIN_POW_HEAD
#define IVPB_LIBM_HEAD static __device__ const
#else
#define IVPB_LIBM_HEAD static __constant__
#endif
#define IVPB_LIBM_FN __device__ __forceinline__
#define IVPB_LIBM_FMA(a, b, c) fma((a), (b), (c))
#define IVPB_LIBM_MUL(a, b) __dmul_rn((a), (b))
#define IVPB_LIBM_ADD(a, b) __dadd_rn((a), (b))
#define IVPB_LIBM_D2U(x) ((unsigned long long)__double_as_longlong(x))
#define IVPB_LIBM_U2D(u) __longlong_as_double((long long)(u))
#define IVPB_LIBM_FALLBACK(x, y) pow((x), (y))
#else
#include <cmath>
#include <cstring>
#define IVPB_LIBM_TABLE static const
#define IVPB_LIBM_HEAD static const
#define IVPB_LIBM_FN static inline
#define IVPB_LIBM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define IVPB_LIBM_MUL(a, b) ((a) * (b))
#define IVPB_LIBM_ADD(a, b) ((a) + (b))
static inline unsigned long long ivpb_libm_d2u(double x) { unsigned long long u; std::memcpy(&u, &x, 8); return u; }
static inline double ivpb_libm_u2d(unsigned long long u) { double x; std::memcpy(&x, &u, 8); return x; }
#define IVPB_LIBM_D2U(x) ivpb_libm_d2u(x)
#define IVPB_LIBM_U2D(u) ivpb_libm_u2d(u)
#define IVPB_LIBM_FALLBACK(x, y) std::pow((x), (y))
#endif

#include "ivpb_libm_pow_tables.cuh"

IVPB_LIBM_FN double ivpb_libm_pow(double x, double y) {
  typedef unsigned long long u64_t;
  u64_t ix = IVPB_LIBM_D2U(x);
  const u64_t iy = IVPB_LIBM_D2U(y);
  const unsigned topx = (unsigned)(ix >> 52), topy = (unsigned)(iy >> 52) & 0x7ffu;
  // e_pow.c: x must be a positive finite number and 2^-65 <= |y| < 2^63 for the main path
  if (topy - 0x3beu > 0x7fu) return IVPB_LIBM_FALLBACK(x, y);
  if (topx - 1u > 0x7fdu) {
    if (topx != 0u || ix == 0ULL) return IVPB_LIBM_FALLBACK(x, y);      // zero, negative, inf, nan
    ix = IVPB_LIBM_D2U(IVPB_LIBM_MUL(x, 4503599627370496.0)) - (52ULL << 52);   // positive subnormal: normalise
  }

  // ---- log_inline: log(x) = hi + lo ----
  const u64_t tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int)((tmp >> 45) & 0x7f);
  const int k = (int)((long long)tmp >> 52);
  const double z = IVPB_LIBM_U2D(ix - (tmp & 0xfff0000000000000ULL));
  const double kd = (double)k;
  const double ln2hi = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[0]), ln2lo = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[1]);
  const double A0 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[2]), A1 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[3]),
               A2 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[4]), A3 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[5]),
               A4 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[6]), A5 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[7]),
               A6 = IVPB_LIBM_U2D(IVPB_POWLOG_HEAD[8]);
  const double invc = IVPB_LIBM_U2D(IVPB_POWLOG_TAB[3 * i]), logc = IVPB_LIBM_U2D(IVPB_POWLOG_TAB[3 * i + 1]),
               logctail = IVPB_LIBM_U2D(IVPB_POWLOG_TAB[3 * i + 2]);
  const double t1 = IVPB_LIBM_FMA(kd, ln2hi, logc);
  const double lo1 = IVPB_LIBM_FMA(kd, ln2lo, logctail);
  const double r = IVPB_LIBM_FMA(z, invc, -1.0);
  const double ar = IVPB_LIBM_MUL(r, A0);
  const double p12 = IVPB_LIBM_FMA(r, A2, A1);
  const double p34 = IVPB_LIBM_FMA(r, A4, A3);
  const double t2 = IVPB_LIBM_ADD(r, t1);
  const double lo2 = IVPB_LIBM_ADD(IVPB_LIBM_ADD(t1, -t2), r);
  const double ar2 = IVPB_LIBM_MUL(r, ar);
  const double ar3 = IVPB_LIBM_MUL(r, ar2);
  const double lo3 = IVPB_LIBM_FMA(ar, r, -ar2);
  const double hi = IVPB_LIBM_ADD(t2, ar2);
  const double p56 = IVPB_LIBM_FMA(r, A6, A5);
  const double lo4 = IVPB_LIBM_ADD(IVPB_LIBM_ADD(t2, -hi), ar2);
  const double q = IVPB_LIBM_FMA(ar2, IVPB_LIBM_FMA(p56, ar2, p34), p12);
  const double lsum = IVPB_LIBM_ADD(IVPB_LIBM_ADD(IVPB_LIBM_ADD(lo1, lo2), lo3), lo4);
  const double lo = IVPB_LIBM_FMA(ar3, q, lsum);
  const double lx = IVPB_LIBM_ADD(hi, lo);
  const double ltail = IVPB_LIBM_ADD(IVPB_LIBM_ADD(hi, -lx), lo);

  // ---- y * log(x) as ehi + elo ----
  const double ehi = IVPB_LIBM_MUL(y, lx);
  const double elo = IVPB_LIBM_FMA(y, ltail, IVPB_LIBM_FMA(lx, y, -ehi));

  // ---- exp_inline ----
  const unsigned abstop = (unsigned)(IVPB_LIBM_D2U(ehi) >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if (abstop < 0x3c9u) return IVPB_LIBM_ADD(1.0, ehi);   // |y log x| < 2^-54: glibc returns 1 + x
    return IVPB_LIBM_FALLBACK(x, y);                        // |y log x| >= 512: overflow / underflow handling
  }
  const double InvLn2N = IVPB_LIBM_U2D(IVPB_EXP_HEAD[0]), Shift = IVPB_LIBM_U2D(IVPB_EXP_HEAD[1]),
               NegLn2hiN = IVPB_LIBM_U2D(IVPB_EXP_HEAD[2]), NegLn2loN = IVPB_LIBM_U2D(IVPB_EXP_HEAD[3]),
               C2 = IVPB_LIBM_U2D(IVPB_EXP_HEAD[4]), C3 = IVPB_LIBM_U2D(IVPB_EXP_HEAD[5]),
               C4 = IVPB_LIBM_U2D(IVPB_EXP_HEAD[6]), C5 = IVPB_LIBM_U2D(IVPB_EXP_HEAD[7]);
  const double zs = IVPB_LIBM_FMA(ehi, InvLn2N, Shift);
  const u64_t ki = IVPB_LIBM_D2U(zs);
  const double kde = IVPB_LIBM_ADD(zs, -Shift);
  double rr = IVPB_LIBM_FMA(kde, NegLn2loN, IVPB_LIBM_FMA(kde, NegLn2hiN, ehi));
  rr = IVPB_LIBM_ADD(elo, rr);
  const int idx = 2 * (int)(ki & 0x7f);
  const u64_t sbits = IVPB_EXP_TAB[idx + 1] + (ki << 45);
  const double tail = IVPB_LIBM_U2D(IVPB_EXP_TAB[idx]);
  const double c23 = IVPB_LIBM_FMA(rr, C3, C2);
  const double tr = IVPB_LIBM_ADD(rr, tail);
  const double r2 = IVPB_LIBM_MUL(rr, rr);
  const double c45 = IVPB_LIBM_FMA(rr, C5, C4);
  const double s1 = IVPB_LIBM_FMA(c23, r2, tr);
  const double r4 = IVPB_LIBM_MUL(r2, r2);
  const double tmpv = IVPB_LIBM_FMA(c45, r4, s1);
  const double scale = IVPB_LIBM_U2D(sbits);
  return IVPB_LIBM_FMA(tmpv, scale, scale);
}
