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
  static constexpr bool HAS_MASS = false;
  IVPB_DEV void mass(const double*, double*) {}
  // IVP::event_config default (src/ivp.rs:51-53 -> EventConfig::new: All, non-terminal)
  IVPB_HD int default_dir(int) { return 0; }
  IVPB_HD i64 default_term(int) { return -1; }
};

}  // namespace ivpb
