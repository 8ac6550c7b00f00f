// ivpb_dense.cu -- Solution::sol / sol_many / sol_span on the device (reference src/solve/solution.rs:25-72,
// ContinuousOutput src/solve/cont.rs:9-154, DenseSegment src/dense.rs:104-147).
//
// The solver kernels log one segment (xold, h, cont) per accepted step into device memory (SolOutDev,
// ivpb_erk.cuh); these kernels answer (trajectory, time) queries against that log: find the first segment
// covering t within 1e-12 -- the reference scans linearly, here a binary search over the monotone segment
// edges lands on the same segment -- and evaluate the method's step interpolant.  Compiled with -fmad=false:
// the interpolation formulas are the reference's, operation for operation (dop853.rs:659-670,
// dopri5.rs:467-478, rk23.rs:313-321, rk4.rs:229-244, radau.rs:798-809, bdf.rs:618-656), so a query on a
// strict-mode solve returns the bits the reference's own Solution::sol would.
#include <cuda_runtime.h>

#include "ivpb_common.cuh"

namespace {

using ivpb::i64;

__device__ void interp_rt(int method, int n, double xi, double* yi, const double* c, double xold, double h) {
  switch (method) {
    case ivpb::M_DOP853: {
      const double s = (xi - xold) / h, s1 = 1.0 - s;
      for (int i = 0; i < n; ++i) {
        const double conpar = c[4 * n + i] + s * (c[5 * n + i] + s1 * (c[6 * n + i] + s * c[7 * n + i]));
        yi[i] = c[i] + s * (c[n + i] + s1 * (c[2 * n + i] + s * (c[3 * n + i] + s1 * conpar)));
      }
      break;
    }
    case ivpb::M_DOPRI5: {
      const double th = (xi - xold) / h, th1 = 1.0 - th;
      for (int i = 0; i < n; ++i)
        yi[i] = c[i] + th * (c[n + i] + th1 * (c[2 * n + i] + th * (c[3 * n + i] + th1 * c[4 * n + i])));
      break;
    }
    case ivpb::M_RK23: {
      const double xc = (xi - xold) / h, x2 = xc * xc, x3 = x2 * xc;
      for (int i = 0; i < n; ++i) yi[i] = c[i] + h * (c[n + i] * xc + c[2 * n + i] * x2 + c[3 * n + i] * x3);
      break;
    }
    case ivpb::M_RK4: {
      const double t = (xi - xold) / h, t2 = t * t, t3 = t2 * t;
      const double h00 = 2.0 * t3 - 3.0 * t2 + 1.0;
      const double h10 = t3 - 2.0 * t2 + t;
      const double h01 = -2.0 * t3 + 3.0 * t2;
      const double h11 = t3 - t2;
      for (int i = 0; i < n; ++i)
        yi[i] = h00 * c[i] + h10 * h * c[n + i] + h01 * c[3 * n + i] + h11 * h * c[2 * n + i];
      break;
    }
    case ivpb::M_RADAU: {
      const double C1M1 = -0.8449489742783178, C2M1 = -0.3550510257216822;
      const double s = (xi - (xold + h)) / h;
      for (int i = 0; i < n; ++i)
        yi[i] = c[i] + s * (c[n + i] + (s - C2M1) * (c[2 * n + i] + (s - C1M1) * c[3 * n + i]));
      break;
    }
    default: {   // BDF, state-major blocks of 7: D0, D1..D5, order
      if (h == 0.0 || n == 0) return;
      const int order = (int)fmin(fmax(round(c[6]), 1.0), 5.0);
      const double x_new = xold + h;
      double p[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
      for (int k = 0; k < order; ++k) {
        const double denom = h * ((double)k + 1.0);
        const double t_shift = x_new - h * (double)k;
        const double xf = (xi - t_shift) / denom;
        p[k] = (k == 0) ? xf : p[k - 1] * xf;
      }
      for (int i = 0; i < n; ++i) {
        double sum = c[i * 7];
        for (int k = 0; k < order; ++k) sum += c[i * 7 + 1 + k] * p[k];
        yi[i] = sum;
      }
    }
  }
}

// One thread per query.  [lo, lo + Ng) is the slice of the batch whose segments live on this device.
__global__ void dense_eval_kernel(int method, int n, int n_cont, int cap, const int* seg_n, const double* seg_x,
                                  const double* seg_cont, i64 M, const i64* traj, i64 lo, i64 Ng, const double* ts,
                                  double* y, int* ok, int extrapolate) {
  const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= M) return;
  const i64 tr = traj[q] - lo;
  if (tr < 0 || tr >= Ng) return;              // another device's shard answers this one
  ok[q] = 0;
  int m = seg_n[tr];
  if (m > cap) m = cap;
  if (m <= 0) return;
  const double t = ts[q], tol = 1e-12;         // cont.rs:105
  const double* sx = seg_x + 2 * tr * (i64)cap;
  const bool fwd = sx[1] >= 0.0;
  // smallest s whose far edge (in the direction of integration) is not yet passed by t
  int a = 0, b = m;
  while (a < b) {
    const int mid = (a + b) >> 1;
    const double xo = sx[2 * mid], hh = sx[2 * mid + 1];
    const double left = fmin(xo, xo + hh), right = fmax(xo, xo + hh);
    const bool reached = fwd ? (t <= right + tol) : (t >= left - tol);
    if (reached) b = mid; else a = mid + 1;
  }
  bool inside = false;
  if (a < m) {
    const double xa = sx[2 * a], ha = sx[2 * a + 1];
    inside = t >= fmin(xa, xa + ha) - tol && t <= fmax(xa, xa + ha) + tol;
  }
  if (!inside) {
    // ContinuousOutput::evaluate_extrapolate (cont.rs:91-150): outside every step, the FIRST segment answers
    // t below its lower edge and the LAST one t above its upper edge -- whatever the direction of integration
    if (!extrapolate) return;
    const double first_left = fmin(sx[0], sx[0] + sx[1]);
    const double last_right = fmax(sx[2 * (m - 1)], sx[2 * (m - 1)] + sx[2 * (m - 1) + 1]);
    if (t < first_left) a = 0;
    else if (t > last_right) a = m - 1;
    else return;
  }
  const double xo = sx[2 * a], hh = sx[2 * a + 1];
  interp_rt(method, n, t, y + q * (i64)n, seg_cont + (tr * (i64)cap + a) * (i64)n_cont, xo, hh);
  ok[q] = 1;
}

__global__ void dense_span_kernel(int cap, const int* seg_n, const double* seg_x, i64 first, i64 count, double* t_start,
                                  double* t_end, int* n_out) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const i64 tr = first + i;
  const int total = seg_n[tr];
  const int m = total < cap ? total : cap;
  n_out[i] = m;
  if (m <= 0) { t_start[i] = 0.0; t_end[i] = 0.0; return; }
  const double* sx = seg_x + 2 * tr * (i64)cap;
  t_start[i] = sx[0];
  t_end[i] = sx[2 * (m - 1)] + sx[2 * (m - 1) + 1];     // cont.rs:67-76
}

}  // namespace

extern "C" cudaError_t ivpb_launch_dense_eval(int method, int n, int n_cont, int cap, const int* seg_n,
                                              const double* seg_x, const double* seg_cont, long long M,
                                              const long long* traj, long long lo, long long Ng, const double* ts,
                                              double* y, int* ok, int extrapolate, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  dense_eval_kernel<<<(unsigned)((M + 127) / 128), 128, 0, stream>>>(method, n, n_cont, cap, seg_n, seg_x, seg_cont, M, traj,
                                                                     lo, Ng, ts, y, ok, extrapolate);
  return cudaGetLastError();
}

extern "C" cudaError_t ivpb_launch_dense_span(int cap, const int* seg_n, const double* seg_x, long long first,
                                              long long count, double* t_start, double* t_end, int* n_out,
                                              cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  dense_span_kernel<<<(unsigned)((count + 127) / 128), 128, 0, stream>>>(cap, seg_n, seg_x, first, count, t_start, t_end, n_out);
  return cudaGetLastError();
}
