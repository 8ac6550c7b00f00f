// ivpb_kernels.cuh -- __global__ entry points and the per-problem lookup used by the runtime.
// Included by one translation unit per built-in problem (ivpb_inst.cu, compiled with
// -DIVPB_PROBLEM=<struct> -DIVPB_PROBLEM_TAG=<name>), and by the NVRTC program for user problems.
#pragma once
#include "ivpb_erk.cuh"

namespace ivpb {

// Small systems leave registers to spare: asking for 5 resident blocks (96 registers/thread) costs no spills
// for n <= 2 and measured +4 % on the north-star kernel; larger systems keep the whole register file.
template <class Prob, int METHOD, int FEAT>
__global__ void __launch_bounds__(IVPB_BLOCK, (Prob::N <= 2 ? 5 : 1)) erk_kernel(const __grid_constant__ KArgs a) {
  erk_body<Prob, METHOD, FEAT>(a);
}

// Kernel variants per (problem, method): feature 0 (final state only), K_OUT (sampled output),
// K_OUT|K_EVENTS (only for problems that define events).
template <class Prob, int METHOD>
__host__ inline const void* erk_lookup_feat(int feat) {
  switch (feat) {
    case 0: return (const void*)&erk_kernel<Prob, METHOD, 0>;
    case K_OUT: return (const void*)&erk_kernel<Prob, METHOD, K_OUT>;
    case K_OUT | K_EVENTS:
      if constexpr (Prob::NEV > 0) return (const void*)&erk_kernel<Prob, METHOD, K_OUT | K_EVENTS>;
      else return nullptr;
    default: return nullptr;
  }
}

template <class Prob>
__host__ inline const void* erk_lookup(int method, int feat) {
  switch (method) {
    case M_RK23: return erk_lookup_feat<Prob, M_RK23>(feat);
    case M_DOPRI5: return erk_lookup_feat<Prob, M_DOPRI5>(feat);
    case M_DOP853: return erk_lookup_feat<Prob, M_DOP853>(feat);
    case M_RK4: return erk_lookup_feat<Prob, M_RK4>(feat);
    default: return nullptr;
  }
}

}  // namespace ivpb
