// ivpb_problems_min.cuh -- the problem interface (device form of the reference's `IVP` trait, src/ivp.rs:27-121)
// without the built-in problems; this is what an NVRTC-compiled user problem derives from.
#pragma once
#include "ivpb_common.cuh"

namespace ivpb {

#define IVPB_DEV static __device__ __forceinline__
#define IVPB_HD static __host__ __device__ __forceinline__

// Defaults shared by all problems: no events, no analytic Jacobian.
template <int N_, int P_, int NEV_>
struct ProblemDefaults {
  static constexpr int N = N_, P = P_, NEV = NEV_;
  static constexpr bool HAS_JAC = false;
  static constexpr bool HAS_ODE_I = false;   // per-component RHS `double ode_i(t, y, p, i)` for the warp-per-trajectory kernels
  IVPB_DEV void events(double, const double*, const double*, double*) {}
  IVPB_DEV void jac(double, const double*, const double*, double*) {}
  // IVP::mass (src/ivp.rs:109-120): constant mass matrix of M y' = f, row-major n x n; used by RADAU when
  // Options.mass_storage = Full (src/methods/radau.rs:283,358-359)
  // The problem's own SolOut (src/solout.rs:55-63) for Options.user_solout = 1:
  //   template <class Interp, class Emit>
  //   static __device__ int solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit);
  // returns 0 Continue | 1 Interrupt | 2 ModifiedSolution (ControlFlag, src/solout.rs:73-78; XOut == Continue on the
  // device, where the step interpolant is always built).  `state`: NSTATE doubles per trajectory, zero at the start (the
  // SolOut struct's fields); dense.valid() is false at the initial call; dense.eval(t, yi) interpolates inside the step;
  // emit(t, y) appends a sample to t_out / y_out.
  static constexpr bool HAS_SOLOUT = false;
  static constexpr int NSTATE = 4;
  static constexpr bool HAS_MASS = false;
  IVPB_DEV void mass(const double*, double*) {}
  // IVP::event_config default (src/ivp.rs:51-53 -> EventConfig::new: All, non-terminal)
  IVPB_HD int default_dir(int) { return 0; }
  IVPB_HD i64 default_term(int) { return -1; }
};

}  // namespace ivpb
