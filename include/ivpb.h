/* ivpb.h -- C ABI of libivpb (ivp-b200): batched IVP solves on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of the Rust crate `ivp` (Ryan-D-Gast/ivp
 * v0.5.1).  The reference has no FFI of its own; its boundary is the generic function
 *     solve_ivp<F: IVP>(f, x0, xend, y0, options) -> Result<Solution, Error>
 *                                              (reference src/solve/solve_ivp.rs:99-108)
 * and the `IVP` trait (src/ivp.rs:27-121).  `ivpb_solve_batch` is the batched form of that
 * call: N independent trajectories of one problem, each with its own y0 row and parameter
 * row, integrated by the per-trajectory step loops of
 *     DOP853::solve  src/methods/dop853.rs:114-656      DOPRI5::solve src/methods/dopri5.rs:122-464
 *     RK23::solve    src/methods/rk23.rs:81-310         RK4::solve    src/methods/rk4.rs:64-226
 *     RADAU::solve   src/methods/radau.rs:114-796       BDF::solve    src/methods/bdf.rs:86-615
 * with the output handler DefaultSolOut::solout (src/solve/solout.rs:128-431) running on the
 * device.  A Rust `ivp-batch` crate binds these symbols (rust/).
 *
 * Conventions: plain pointers and sizes only; all arrays row-major; return 0 on success,
 * non-zero for configuration / CUDA / NVRTC errors (== reference `Error::Config`,
 * src/error.rs:18-60; message via ivpb_last_error).  Numerical failures are NOT errors: they are
 * reported per trajectory in `status[]` (reference src/status.rs:4-19, declaration order).
 * There is no CPU fallback: without a CUDA device every compute entry point fails loudly.
 */
#ifndef IVPB_H
#define IVPB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/solve/options.rs:14-27 (enum Method, declaration order) */
typedef enum {
  IVPB_RK23 = 0, IVPB_DOPRI5 = 1, IVPB_DOP853 = 2, IVPB_RK4 = 3, IVPB_RADAU = 4, IVPB_BDF = 5
} ivpb_method;

/* reference src/status.rs:4-19 (enum Status, declaration order) */
typedef enum {
  IVPB_SUCCESS = 0, IVPB_USER_INTERRUPT = 1, IVPB_NEED_LARGER_NMAX = 2, IVPB_STEP_SIZE_TOO_SMALL = 3,
  IVPB_PROBABLY_STIFF = 4, IVPB_SINGULAR_MATRIX = 5, IVPB_POOR_CONVERGENCE = 6
} ivpb_status;

/* Built-in problems: the RHS / events / Jacobian are __device__ functions compiled with the solver.
 * n = state size, p = parameters per trajectory. */
typedef enum {
  IVPB_P_DECAY = 0,      /* y' = -k y                    n=1 p=1 (k)        reference examples/exponential_decay.rs:10-12 */
  IVPB_P_VDP_EPS = 1,    /* y1' = ((1-y0^2) y1 - y0)/eps n=2 p=1 (eps)      reference examples/van_der_pol.rs:10-13 */
  IVPB_P_VDP_MU = 2,     /* y1' = mu (1-y0^2) y1 - y0    n=2 p=1 (mu)       reference benches/benchmark.py:22-27 */
  IVPB_P_LORENZ = 3,     /* n=3 p=3 (sigma, rho, beta)                      reference benches/benchmark.py:30-37 */
  IVPB_P_CR3BP = 4,      /* n=6 p=1 (mu)                                    reference examples/cr3bp.rs:24-35 */
  IVPB_P_BALL = 5,       /* n=2 p=2 (g, drag); event y[0], terminal, negative  examples/bouncing_ball.rs:11-31 */
  IVPB_P_ROBERTSON = 6,  /* n=3 p=3 (0.04, 1e4, 3e7)                        reference tests/test_stiff.py:104-110 */
  IVPB_P_SHO = 7,        /* y0'=y1, y1'=-y0  n=2 p=0; event y[0] (All, non-terminal)  reference tests/common.rs:3-9, tests/ivp.rs:151-220 */
  IVPB_P_ZERO3 = 8,      /* y' = 0          n=3 p=0                         reference tests/ivp.rs:11-18 */
  IVPB_P_EXP2 = 9,       /* y' = y          n=2 p=0                         reference tests/ivp.rs:291-297 */
  IVPB_P_RATIONAL = 10,  /* n=2 p=0                                         reference tests/test_helpers.py:23-25 */
  IVPB_P_CANNON = 11,    /* y' = [y1, -9.80665] n=2 p=0; event y[0], terminal, negative  reference tests/test_ivp.py:153-160 */
  IVPB_P_LINEAR100 = 12, /* y' = -y         n=100 p=0 (warp-per-trajectory kernels)     reference benches/benchmark.py:39-41,137-146 */
  IVPB_P_MEDAKZO64 = 13, /* MEDAKZO on 32 grid points  n=64 p=0                      reference tests/test_ivp.py:77-101 */
  IVPB_P_ROBERTSON_DAE = 14, /* Robertson as an index-1 DAE: M = diag(1, 1, 0), third row x + y + z - 1 = 0;  n=3 p=3; needs
                                mass_storage = 1 (RADAU).  Exercises src/methods/radau.rs:375-386,525-539,626-634 */
  IVPB_P_MASS_LINEAR3 = 15,  /* M y' = A y with a full (non-diagonal, invertible) constant M; n=3 p=1 (scales A); RADAU + mass */
  IVPB_P_BALL_BOUNCE = 16,   /* bouncing ball that bounces INSIDE the solve: n=2 p=3 (g, drag, restitution); its own SolOut hook
                                (user_solout = 1) finds each impact on the step interpolant, reverses and damps the velocity
                                and returns ControlFlag::ModifiedSolution (src/solout.rs:18-29,73-78; the reference's
                                examples/bouncing_ball.py:14-36 restarts solve_ivp from the host instead) */
  IVPB_P_BUILTIN_COUNT = 17
} ivpb_builtin;

/* Mirrors `Options` (reference src/solve/options.rs:75-123) plus the per-event `EventConfig`
 * (src/solve/event.rs:5-27) that the reference takes from `IVP::event_config`. */
typedef struct {
  int32_t method;              /* ivpb_method; reference default DOPRI5 */
  int32_t n_rtol, n_atol;      /* 1 = Tolerance::Scalar, n = Tolerance::Vector (src/methods/mod.rs:104-107) */
  const double* rtol;          /* reference default 1e-3 */
  const double* atol;          /* reference default 1e-6 */
  int32_t has_first_step, has_max_step, has_min_step, has_max_steps;   /* Option<..> discriminants */
  double first_step, max_step, min_step;
  uint64_t max_steps;          /* None => usize::MAX (src/solve/solve_ivp.rs:187-273) */
  int32_t has_t_eval;          /* Option<Vec<Float>>: Some(empty) is legal */
  int32_t n_t_eval;
  const double* t_eval;        /* shared by all trajectories; ascending (descending for tf < t0) */
  int32_t dense_output;        /* keep every accepted step's interpolant (Solution::sol; src/solve/solout.rs:141-146,
                                  src/solve/cont.rs:9-30) in device memory; needs max_segments >= 1; does not change t/y */
  int32_t n_event_cfg;         /* 0 => problem defaults; else must equal the problem's n_events */
  const int32_t* ev_direction;        /* >0 Positive, <0 Negative, 0 All (src/solve/event.rs:68-77) */
  const int64_t* ev_terminal_count;   /* <0 => None */
  int32_t max_events;          /* capacity of ev_t / ev_y per event function per trajectory (>=1 if events) */
  int32_t max_out;             /* step-mode (no t_eval) capacity of t_out / y_out per trajectory; 0 => samples not stored */
  int32_t jac_mode;            /* 0 finite differences (src/ivp.rs:67-107), 1 analytic ivp_jac */
  int32_t flags;               /* IVPB_FLAG_* */
  int32_t max_segments;        /* dense_output: interpolant segments stored per trajectory (one per accepted step) */
  /* RADAU only (src/solve/solve_ivp.rs:246-258; every other method ignores them, like the reference): */
  int32_t mass_storage;        /* Options.mass_storage: 0 = Identity (default, y' = f), 1 = Full: M y' = f with the problem's
                                  constant mass matrix (IVP::mass, src/ivp.rs:109-120; device hook `mass` / `ivp_mass`) */
  int32_t user_solout;         /* 1: the problem's own SolOut hook (`solout` / `ivp_solout`, src/solout.rs:55-63) replaces
                                  DefaultSolOut -- the reference's low-level `Method::solve(.., Some(&mut solout))` call
                                  (e.g. src/methods/dop853.rs:114-127): no t_eval / events / dense log; samples the hook emits
                                  go to t_out / y_out (capacity max_out).  All six methods, every state size: in the
                                  warp-per-trajectory kernels (n > 32; RADAU / BDF n > 8) the hook runs warp-uniformly on the
                                  full state in shared memory and must take its eval() buffer from dense.buffer(). */
  int32_t nind1, nind2, nind3; /* Options.nind1..3: index-1/2/3 variable counts of a DAE, < 0 => None
                                  (partition rules and Error::Config of src/methods/radau.rs:210-245) */
  /* RADAU / BDF with jac_mode = 0: `jac_sparsity` of the reference's Python front end (src/python/solve.rs, SparsityStructure
   * src/python/sparsity.rs:13-102).  The structural non-zeros of the Jacobian in compressed-column form; the runtime groups
   * structurally orthogonal columns with the reference's greedy first-fit rule (sparsity.rs:109-154) and the kernels take one
   * RHS evaluation per GROUP instead of one per column (sparse_jacobian_fd, sparsity.rs:160-202); entries outside the
   * pattern stay 0.  has_jac_sparsity = 0 => dense forward differences. */
  int32_t has_jac_sparsity;
  const int32_t* jac_sparsity_colptr;   /* [n + 1]  column starts (colptr[n] = number of structural non-zeros) */
  const int32_t* jac_sparsity_rows;     /* [colptr[n]]  row indices, ascending within a column */
} ivpb_options;

#define IVPB_FLAG_STRICT_FP 1u /* run the kernel variant compiled with -fmad=false (operation-for-operation
                                  the reference's rounding, no FMA contraction).  Without this flag and without
                                  IVPB_FLAG_FAST_FP the explicit methods let a parity pilot choose (ivpb_last_fp_mode) */
#define IVPB_FLAG_NO_REFILL 2u /* static one-trajectory-per-thread schedule (debug / A-B measurements) */
#define IVPB_FLAG_FAST_FP 8u   /* run the FMA-contracted kernels unconditionally.  The implicit methods default to the
                                  strict arithmetic (bit-identical to the reference's operation sequence) because stiff
                                  ensembles lose step-count parity under any last-bit perturbation (VdP mu=1000: 73 % RADAU,
                                  97 % BDF with FMA) while the FMA build is only 17-29 % faster there.  The explicit methods
                                  use the FMA build (2x faster) when the parity pilot finds it inside the tolerance of the
                                  strict build on a sample of the ensemble (north star: yes; CR3BP at 1e-10: no). */
#define IVPB_FLAG_NO_SORT 16u  /* RADAU / BDF: hand the trajectories out in index order instead of the locality order (a Morton
                                  curve through the varying coordinates of y0 / params, so that the lanes of a warp hold
                                  neighbouring initial conditions and diverge less; results are identical either way) */
#define IVPB_FLAG_SORT 32u     /* explicit methods: use the locality order too (off by default there: small gain on the device,
                                  a loss end to end with mapped host buffers, see ivpb_runtime.cu launch_shard) */
#define IVPB_FLAG_ZEROCOPY_OUT 64u /* ivpb_solve_batch: let the kernel store results straight into the caller's page-locked
                                    buffers (round 1's route) instead of staging them in device memory and copying them
                                    out chunk by chunk while the kernel integrates the rest (completion flags).  Only
                                    worth it for long kernels with small outputs on an otherwise idle PCIe root complex. */
#define IVPB_FLAG_NO_PIPELINE 128u /* ivpb_solve_batch: copy the results out after the kernel has finished instead of chunk
                                    by chunk while it runs (A-B measurements) */
#define IVPB_FLAG_NO_ZEROCOPY 4u /* ivpb_solve_batch: always stage through device buffers, even when the caller's
                                    buffers are page-locked (default: pinned y0 / params rows are read by the kernel
                                    directly over PCIe when a trajectory starts) */

/* Per-trajectory outputs.  Any pointer may be NULL (= not wanted).  In ivpb_solve_batch they are
 * HOST pointers, in ivpb_solve_batch_device DEVICE pointers.
 * out_cap = has_t_eval ? n_t_eval + 1 : max_out   (+1: a terminal event appends its point,
 * reference src/solve/solout.rs:315-324). */
typedef struct {
  int32_t* status;      /* [N]      ivpb_status                                         */
  uint32_t* counters;   /* [N][6]   nfev, njev, nlu, nstep, naccpt, nrejct (src/solve/solution.rs:12-17) */
  double* t_final;      /* [N]      last accepted x, or the event time if a terminal event fired */
  double* y_final;      /* [N][n]   state at t_final                                    */
  double* h_next;       /* [N]      IntegrationResult.h (src/methods/mod.rs:31-32)      */
  int32_t* n_out;       /* [N]      samples the reference would hold in Solution.t (may exceed out_cap => truncated) */
  double* t_out;        /* [N][out_cap]                                                 */
  double* y_out;        /* [N][out_cap][n]                                              */
  int32_t* ev_count;    /* [N][n_events]   hits per event function (may exceed max_events => truncated) */
  double* ev_t;         /* [N][n_events][max_events]                                    */
  double* ev_y;         /* [N][n_events][max_events][n]                                 */
  /* dense_output: the reference's ContinuousOutput segments (src/dense.rs:104-147).  Optional copies -- the
   * segments always stay resident on the device for ivpb_dense_eval. */
  int32_t* n_seg;       /* [N]      accepted steps with h != 0 (may exceed max_segments => truncated) */
  double* seg_x;        /* [N][max_segments][2]        xold, h of each segment           */
  double* seg_cont;     /* [N][max_segments][c*n]      cont in the reference's layout: coefficient-major cont[k*n+i]
                                                        (c = Method::coeffs_per_state), BDF state-major cont[i*7+k] */
} ivpb_outputs;

typedef struct ivpb_ctx ivpb_ctx;

/* Return codes (0 = ok).  Non-zero mirrors the reference's `Err(Error::Config(..))` -- raised before any
 * stepping -- or reports a CUDA / NVRTC failure. */
#define IVPB_OK 0
#define IVPB_ERR_CONFIG 1
#define IVPB_ERR_CUDA 2
#define IVPB_ERR_NVRTC 3

/* Context: owns the device set, one stream + work queue per device, compiled NVRTC modules.
 * device_ids == NULL => the current device only. */
int ivpb_create(ivpb_ctx** out, const int* device_ids, int n_devices);
void ivpb_destroy(ivpb_ctx* ctx);
const char* ivpb_last_error(const ivpb_ctx* ctx);   /* ctx may be NULL: error of the last failed create */
int ivpb_device_count(const ivpb_ctx* ctx);

/* Problem handles.  Built-ins: handle == ivpb_builtin id.  *n, *p, *n_events receive the sizes. */
int ivpb_builtin_problem(ivpb_ctx* ctx, int builtin_id, int* n, int* p, int* n_events);
/* User problem in CUDA C, compiled with NVRTC together with the solver templates.  `cuda_src` must define
 *   __device__ void ivp_ode(double t, const double* y, const double* p, double* dydt);
 * and, if n_events > 0 / has_jac (bit 0: analytic Jacobian `ivp_jac`; bit 1: constant mass matrix
 * `__device__ void ivp_mass(const double* p, double* M)`, row-major n x n, IVP::mass of src/ivp.rs:109-120; bit 2: own
 * SolOut `template <class Interp, class Emit> __device__ int ivp_solout(double xold, double& x, double* y, const double* p,
 * double* state, const Interp& dense, Emit& emit)` returning 0 Continue / 1 Interrupt / 2 ModifiedSolution, used when
 * ivpb_options.user_solout = 1 -- dense.valid(), dense.eval(t, yi), emit(t, y), state = 4 doubles per trajectory),
 *   __device__ void ivp_events(double t, const double* y, const double* p, double* g);
 *   __device__ void ivp_jac(double t, const double* y, const double* p, double* J);  // row-major n x n
 * (the IVP trait, reference src/ivp.rs:27-121, with the parameter row made explicit). */
int ivpb_nvrtc_problem(ivpb_ctx* ctx, const char* cuda_src, int n, int p, int n_events, int has_jac,
                       int* handle);

/* Batched solve_ivp with HOST buffers (the call a Rust/C++/Python user makes): copies y0/params to the
 * devices of the context (static contiguous split of [0,N)), solves, copies the requested outputs back.
 * y0 [N][n], params [N][p] (NULL if p == 0). */
int ivpb_solve_batch(ivpb_ctx* ctx, int problem, const ivpb_options* opt, int64_t N, double t0, double tf,
                     const double* y0, const double* params, const ivpb_outputs* out);

/* Same solve with DEVICE buffers resident on the context's first device, enqueued on `stream`
 * (a cudaStream_t, may be NULL = default stream); returns after enqueueing. rtol/atol/t_eval/event
 * config in `opt` stay host pointers (small, copied to constant-like device storage).
 * One solve is in flight per context: a call enqueued on another stream starts on the device after the previous call of
 * the same context has finished (the per-device work-queue counter and staging buffers are shared; the runtime orders the
 * calls with an event).  Use one context per stream for solves that should overlap.  A context is not thread-safe. */
int ivpb_solve_batch_device(ivpb_ctx* ctx, int problem, const ivpb_options* opt, int64_t N, double t0,
                            double tf, const double* d_y0, const double* d_params,
                            const ivpb_outputs* d_out, void* stream);

/* Solution::sol / sol_many on the device (reference src/solve/solution.rs:25-72, ContinuousOutput::evaluate
 * src/solve/cont.rs:82-114): evaluates the dense output retained by the most recent ivpb_solve_batch call that
 * had dense_output = 1.  Query q: trajectory traj[q] (index into that batch), time ts[q]; y[q][n] receives the
 * interpolated state, ok[q] = 1 if a stored segment covers ts[q] within 1e-12 (first match in step order, like
 * the reference's linear search), else 0 and y[q] is untouched.  All pointers are HOST pointers. */
int ivpb_dense_eval(ivpb_ctx* ctx, uint64_t generation, int n, int64_t n_query, const int64_t* traj, const double* ts,
                    double* y, int32_t* ok);
/* Which dense log the context retains: a counter incremented by every dense_output solve (0 = none yet).  The three
 * dense calls take the generation the caller's Solution belongs to and its state size n and fail with IVPB_ERR_CONFIG
 * when the retained log is a different one (a newer dense solve replaced it) or has another state size -- so a stale
 * Solution can neither read someone else's trajectory nor overrun its y buffer (n doubles per query).  generation = 0
 * skips the identity check ("whatever is retained"); n is always checked.  The log lives in buffers of its own: solves
 * without dense_output, on either entry point, leave it intact. */
uint64_t ivpb_dense_generation(const ivpb_ctx* ctx);
/* ContinuousOutput::evaluate_extrapolate (src/solve/cont.rs:91-150; what the reference's Python OdeSolution.__call__
 * uses, src/python/solution.rs:41,116): as ivpb_dense_eval, but a time outside every stored step is answered by the
 * FIRST segment when it lies below that segment's lower edge and by the LAST segment when it lies above that one's
 * upper edge (the reference's rule, independent of the direction of integration).  ok[q] = 0 only for a trajectory
 * without segments. */
int ivpb_dense_eval_extrapolate(ivpb_ctx* ctx, uint64_t generation, int n, int64_t n_query, const int64_t* traj,
                                const double* ts, double* y, int32_t* ok);
/* Solution::sol_span for trajectories [first, first + count): t_start = first.xold, t_end = last.xold + last.h
 * (src/solve/cont.rs:67-76); n_seg = segments stored (0 => no span). */
int ivpb_dense_span(ivpb_ctx* ctx, uint64_t generation, int64_t first, int64_t count, double* t_start, double* t_end,
                    int32_t* n_seg);

/* Pinned host memory for y0 / params / outputs: ivpb_solve_batch maps such buffers into the kernel (zero-copy,
 * see IVPB_FLAG_NO_ZEROCOPY) or, for the staged fields, copies at full PCIe rate.  Pageable buffers work too,
 * through staging copies. */
void* ivpb_host_alloc(size_t bytes);
void ivpb_host_free(void* p);

/* Kernels launched by this context so far (for bench.py's gpu_launches). */
uint64_t ivpb_launch_count(const ivpb_ctx* ctx);

/* Which kernel build the most recent solve of this context ran: returns 1 = strict (the reference's arithmetic,
 * operation for operation), 0 = FMA build, and fills info = {strict, source, sample, status mismatches, step-count
 * mismatches, trajectories out of tolerance}.  source: 0 the caller's flag (IVPB_FLAG_STRICT_FP / IVPB_FLAG_FAST_FP),
 * 1 the method's default (RADAU / BDF: strict), 2 a cached verdict of the parity pilot, 3 the parity pilot just ran.
 * The pilot (explicit methods without either flag): the first solve of a configuration integrates `sample` evenly
 * strided trajectories of the ensemble with BOTH builds and picks the FMA build only if all of them end with the same
 * status and inside HALF of max(10 rtol |y|, 10 atol) of the strict result (the margin is for the trajectories the pilot
 * does not see), and >= 99 % with the same accepted / rejected step counts (ivpb_runtime.cu resolve_fp). */
int ivpb_last_fp_mode(const ivpb_ctx* ctx, int32_t info[6]);

/* FP64 FMA-pipe peak of the first device measured with a dependent-chain DFMA microbenchmark
 * (TFLOP/s, FMA = 2 flop).  Used as the roofline denominator (MEASURED_PEAKS.json has no fp64 entry). */
int ivpb_measure_fp64_peak(ivpb_ctx* ctx, double* tflops);

const char* ivpb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* IVPB_H */
