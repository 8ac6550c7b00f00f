// ivpb_common.cuh -- shared device-side types for the sm_100a IVP kernels.
// This header (and everything it is combined with) is compiled twice: ahead of time by nvcc for the
// built-in problems, and at run time by NVRTC together with user CUDA C (`ivp_ode`, ...).  It therefore
// uses no standard headers -- only CUDA built-ins.
#pragma once

namespace ivpb {

typedef unsigned int u32;
typedef long long i64;
typedef unsigned long long u64;

// Method codes == reference enum Method (src/solve/options.rs:14-27)
enum { M_RK23 = 0, M_DOPRI5 = 1, M_DOP853 = 2, M_RK4 = 3, M_RADAU = 4, M_BDF = 5 };
// Status codes == reference enum Status (src/status.rs:4-19)
enum { ST_SUCCESS = 0, ST_INTERRUPT = 1, ST_NMAX = 2, ST_SMALL = 3, ST_STIFF = 4, ST_SINGULAR = 5, ST_POOR = 6 };
enum { ST_RERUN = 100 };   // internal, never written out: a strictd trajectory abandoned for the guarded re-run (ivpb_exact.cuh)
// Output-handler features a kernel instance is specialised for (what DefaultSolOut has to do,
// reference src/solve/solout.rs:128-431).  FEAT == 0: final state + counters only, and the dense-output
// coefficients nobody would read are not computed.
enum { F_TEVAL = 1, F_STEPOUT = 2, F_EVENTS = 4 };

constexpr int MAX_N = 32;       // thread-per-trajectory kernels: state size limit
constexpr int MAX_EVENTS_FN = 8;  // event functions per problem

// Kernel arguments (one struct by value; ~1 KB of the 4 KB parameter space).  Host layout must match:
// the runtime fills this struct for nvcc-built kernels and memcpy's the same bytes for NVRTC kernels.
struct KArgs {
  i64 N;                 // trajectories in this launch (this device's shard)
  double t0, tf;
  const double* y0;      // [N][n] row-major
  const double* params;  // [N][p] row-major (may be null when p == 0)
  u64* queue;            // work-queue head (atomicAdd), zeroed before launch
  const unsigned* perm;  // locality order (null = identity): queue position k integrates trajectory perm[k], so the lanes of a
                         // warp hold neighbouring initial conditions and take (nearly) the same accept / reject / Newton paths
  double rtol[MAX_N], atol[MAX_N];   // scalar tolerances are broadcast by the host
  const double* rtol_ext;            // n > MAX_N with Tolerance::Vector: device arrays [n] (else null: rtol[0])
  const double* atol_ext;
  double* scratch;       // implicit warp kernels: one slot per warp of the grid -- the Jacobian (n x (n|1) doubles), RADAU's
                         // mass matrix, and for large n the iteration matrices as well (warp_impl_shape below)
  // jac_sparsity (src/python/sparsity.rs): compressed columns of the Jacobian's structure and the column groups the
  // runtime built from them; sp_colptr == null => dense forward differences
  const int* sp_colptr;  // [n + 1]
  const int* sp_rows;    // [sp_colptr[n]]
  const int* sp_group;   // [n] group of each column
  int sp_ngroups;
  // strictd kernels (deferred division / square-root guards, ivpb_exact.cuh): trajectories that have to be repeated by
  // the guarded twin are appended here; the twin's launch gets perm = rerun_list and n_dev = rerun_count
  unsigned* rerun_list;      // [N]
  unsigned* rerun_count;     // zeroed before the first pass
  const unsigned* n_dev;     // second pass: number of trajectories, read on the device (null: KArgs::N)
  int debug_rerun_mod;       // > 0: the strictd kernel lists every k-th trajectory as if a guard had failed (tests)
  double first_step, max_step, min_step;
  int has_first_step, has_max_step, has_min_step, static_sched;
  u64 max_steps;         // usize::MAX when Options.max_steps is None
  const double* t_eval;  // device copy, [n_t_eval]
  int n_t_eval, out_cap;
  int max_events, jac_mode;
  int zero_tail;         // t_out / y_out are the caller's mapped host buffers: `finish` zero-fills the slots a trajectory
                         // left unwritten (nobody memsets them), so the whole row is defined when the kernel ends
  int vec_io;            // y0 / y_final rows are 16-byte aligned: move them as double2 (LDG.128 / STG.128)
  int nind1, nind2, nind3;   // RADAU with a mass matrix: resolved DAE partition (radau.rs:210-245); nind1 + nind2 + nind3 == n
  double newton_tol;     // implicit methods: Newton stopping tolerance (radau.rs:198-205, bdf.rs:174-184), host-computed
  int ev_dir[MAX_EVENTS_FN];
  i64 ev_term[MAX_EVENTS_FN];   // < 0: not terminal
  // outputs (any may be null)
  int* status;
  u32* counters;
  double* t_final;
  double* y_final;
  double* h_next;
  int* n_out;
  double* t_out;
  double* y_out;
  int* ev_count;
  double* ev_t;
  double* ev_y;
  // Completion flags of the host-buffer path (ivpb_solve_batch): the shard is cut into chunks of `chunk_size` consecutive
  // trajectories; the thread that retires the LAST trajectory of a chunk raises chunk_flag[c] in mapped host memory, and
  // the host then copies that chunk's results out while the kernel keeps integrating the others.  chunk_size == 0: off.
  i64 chunk_size;
  unsigned* chunk_count;     // device, [chunks], zeroed before the launch: trajectories of the chunk retired so far
  int* chunk_flag;           // page-locked host memory mapped into the device address space, [chunks]
  // Arrival flags of the host-buffer path: y0 / params reach the device in chunks of `in_chunk` consecutive trajectories
  // while the kernel is already running; in_flag[c] (device memory, written by the copy engine after chunk c's data)
  // tells the scheduler that the rows of chunk c may be read.  in_chunk == 0: everything is resident at launch.
  i64 in_chunk;
  const int* in_flag;
  // dense_output: per-trajectory interpolant log (src/solve/solout.rs:141-146); seg_cap == 0 => off
  int seg_cap, n_cont;   // n_cont = coeffs_per_state * n doubles per segment
  int* seg_n;
  double* seg_x;
  double* seg_cont;
};

// Shape of the warp-cooperative RADAU / BDF kernels (ivpb_implicit_warp.cuh), shared by the device templates and the two
// launchers (ivpb_runtime.cu, ivpb_nvrtc.cpp).  Per warp, shared memory holds the layout's 2n doubles, the staged vectors
// (RADAU: b1..b3 + pivots = 4n; BDF: b1, D (8 rows), scratch (6 rows), pivots = 16n) and -- when they fit 227 KB -- the
// iteration matrices (RADAU: E1, E2re, E2im; BDF: the LU of I - cJ).  The Jacobian (and RADAU's mass matrix) always sit in
// the warp's slot of KArgs::scratch.  When the matrices do not fit (RADAU n > 84, BDF n > 118: the reference's own
// MEDAKZO test has n = 400) they move to the same slot in global memory (`gmats`), where they are L2-resident.
struct WarpImplShape {
  long long smem_doubles;      // per warp
  int warps;                   // per block
  long long scratch_doubles;   // per warp, in KArgs::scratch
  bool gmats;
};
__host__ __device__ constexpr WarpImplShape warp_impl_shape(int n, int method, bool mass) {
  const long long matd = (long long)(n | 1) * n;
  const long long vecs = (method == M_RADAU ? 6LL : 18LL) * n;
  const long long mats = method == M_RADAU ? 3 : 1;
  const bool g = (vecs + mats * matd) * 8 > 227 * 1024;
  const long long smem = vecs + (g ? 0 : mats * matd);
  const int warps = g ? 1 : (smem * 8 * 4 <= 200 * 1024 ? 4 : (smem * 8 * 2 <= 200 * 1024 ? 2 : 1));
  const long long scratch = (1 + ((method == M_RADAU && mass) ? 1 : 0) + (g ? mats : 0)) * matd;
  return WarpImplShape{smem, warps, scratch, g};
}

// std::conditional without <type_traits> (NVRTC has no standard headers)
template <bool B, class T, class F> struct std_conditional { typedef T type; };
template <class T, class F> struct std_conditional<false, T, F> { typedef F type; };

__device__ __forceinline__ double signum(double x) {   // f64::signum: +-1 by sign bit, NaN -> NaN
  return (x != x) ? x : copysign(1.0, x);
}

}  // namespace ivpb
