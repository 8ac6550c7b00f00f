// ivpb_kernels.cuh -- __global__ entry points and the per-problem lookup used by the runtime.
// Included by one translation unit per built-in problem (ivpb_inst.cu, compiled with
// -DIVPB_PROBLEM=<struct> -DIVPB_PROBLEM_TAG=<name>), and by the NVRTC program for user problems.
#pragma once
#include "ivpb_erk.cuh"
#ifdef IVPB_WITH_IMPLICIT
#include "ivpb_implicit.cuh"
#include "ivpb_implicit_warp.cuh"
#endif

namespace ivpb {

// Small systems (n <= 2) leave registers to spare: asking for 7 resident blocks per SM caps the registers at 72 (24 bytes
// of spills in the DOP853 kernel) and keeps 28 warps resident.  North star, ms per 2^20 trajectories with 5 / 6 / 7 / 8 /
// 10 blocks: (5 measured +4 % over 1 earlier) 15.54 / 15.49 / 15.21 / 17.09 / 20.07; VdP DOPRI5 19.79 (5) / 18.93 (7) / 24.08 (8).  Larger systems keep
// the whole register file.  Kernels with a user SolOut hook (K_USER) stay at 5: the hook's own code needs the registers
// (ball_bounce: 6.8 ms at 5 blocks, 7.3 ms at 7).
#ifndef IVPB_MB_SMALL
#define IVPB_MB_SMALL 7
#endif
#ifndef IVPB_MB_MID
#define IVPB_MB_MID 5      // n = 3, 4: Lorenz DOPRI5 8.18 (1) / 7.88 (5) / 9.47 (6) ms per 2^20 trajectories
#endif
#ifndef IVPB_MB_BIG
#define IVPB_MB_BIG 1
#endif
// Threads per block of the thread-per-trajectory kernels: 128, or 256 for the block-synchronous kernels of the larger
// systems (one block per SM at 255 registers, so all eight resident warps run in step; ErkTraj::BLOCK_SYNC).
template <class Prob, int METHOD, int FEAT>
__host__ __device__ constexpr int erk_block_threads() {
  return (ErkTraj<Prob, METHOD, FEAT>::BLOCK_SYNC && Prob::N > 4) ? 2 * IVPB_BLOCK : IVPB_BLOCK;
}
// PIPE: with (host-buffer path) or without (device-resident path) the arrival / completion flags of run_schedule.
template <class Prob, int METHOD, int FEAT, bool PIPE = true>
__global__ void __launch_bounds__((erk_block_threads<Prob, METHOD, FEAT>()), (Prob::N <= 2 ? ((FEAT & K_USER) ? 5 : IVPB_MB_SMALL) : (Prob::N <= 4 ? IVPB_MB_MID : IVPB_MB_BIG))) erk_kernel(const __grid_constant__ KArgs a) {
  erk_body<Prob, METHOD, FEAT, PIPE>(a);
}

// Kernel variants per (problem, method): feature 0 (final state only), K_OUT (sampled output),
// K_OUT|K_EVENTS (only for problems that define events).
template <class Prob, int METHOD>
__host__ inline const void* erk_lookup_feat(int feat) {
  switch (feat) {
    case K_NOPIPE: return (const void*)&erk_kernel<Prob, METHOD, 0, false>;
    case K_NOPIPE | K_OUT: return (const void*)&erk_kernel<Prob, METHOD, K_OUT, false>;
    case K_NOPIPE | K_OUT | K_EVENTS:
      if constexpr (Prob::NEV > 0) return (const void*)&erk_kernel<Prob, METHOD, K_OUT | K_EVENTS, false>;
      else return nullptr;
    case 0: return (const void*)&erk_kernel<Prob, METHOD, 0>;
    case K_OUT: return (const void*)&erk_kernel<Prob, METHOD, K_OUT>;
    case K_OUT | K_EVENTS:
      if constexpr (Prob::NEV > 0) return (const void*)&erk_kernel<Prob, METHOD, K_OUT | K_EVENTS>;
      else return nullptr;
    case K_USER:                         // Options.user_solout: the problem's own SolOut
      if constexpr (Prob::HAS_SOLOUT) return (const void*)&erk_kernel<Prob, METHOD, K_USER>;
      else return nullptr;
    default: return nullptr;
  }
}

// Warp-per-trajectory variant for n > 32 (WarpLayout, ivpb_erk.cuh): 4 warps per block, dynamic shared memory
// IVPB_BLOCK / 32 * WarpLayout::SMEM_DOUBLES_PER_WARP doubles.
template <class Prob, int METHOD, int FEAT>
__global__ void __launch_bounds__(IVPB_BLOCK, 1) erk_warp_kernel(const __grid_constant__ KArgs a) {
  erk_warp_body<Prob, METHOD, FEAT>(a);
}

template <class Prob, int METHOD>
__host__ inline const void* erk_lookup_feat_any(int feat) {
  if constexpr (Prob::N <= MAX_N) return erk_lookup_feat<Prob, METHOD>(feat);
  else {
    switch (feat & ~K_NOPIPE) {      // the warp kernels have no twin: one flag test per trajectory is nothing there
      case 0: return (const void*)&erk_warp_kernel<Prob, METHOD, 0>;
      case K_OUT: return (const void*)&erk_warp_kernel<Prob, METHOD, K_OUT>;
      case K_OUT | K_EVENTS:
        if constexpr (Prob::NEV > 0) return (const void*)&erk_warp_kernel<Prob, METHOD, K_OUT | K_EVENTS>;
        else return nullptr;
      case K_USER:                       // Options.user_solout: the problem's own SolOut (WarpHook, ivpb_erk.cuh)
        if constexpr (Prob::HAS_SOLOUT) return (const void*)&erk_warp_kernel<Prob, METHOD, K_USER>;
        else return nullptr;
      default: return nullptr;
    }
  }
}

// block size of the kernel erk_lookup returns for (method, feat)
template <class Prob>
__host__ inline int erk_lookup_block(int method) {
  if constexpr (Prob::N > MAX_N) return IVPB_BLOCK;
  else {
    switch (method) {      // BLOCK_SYNC does not depend on the feature set
      case M_RK23: return erk_block_threads<Prob, M_RK23, 0>();
      case M_DOPRI5: return erk_block_threads<Prob, M_DOPRI5, 0>();
      case M_DOP853: return erk_block_threads<Prob, M_DOP853, 0>();
      default: return erk_block_threads<Prob, M_RK4, 0>();
    }
  }
}

template <class Prob>
__host__ inline const void* erk_lookup(int method, int feat) {
  switch (method) {
    case M_RK23: return erk_lookup_feat_any<Prob, M_RK23>(feat);
    case M_DOPRI5: return erk_lookup_feat_any<Prob, M_DOPRI5>(feat);
    case M_DOP853: return erk_lookup_feat_any<Prob, M_DOP853>(feat);
    case M_RK4: return erk_lookup_feat_any<Prob, M_RK4>(feat);
    default: return nullptr;
  }
}


#ifdef IVPB_WITH_IMPLICIT
// Implicit kernels (ivpb_implicit.cuh).  They are latency-bound (ncu, profiles/r1h_*_ncu_full.txt: issue slots 21-26 %
// busy, top stalls "no instruction" and "wait", 8 resident warps per SM at 178-182 registers), so the register-resident
// variants (n <= 3) trade registers for resident warps: the launch bound below caps the registers at 168 / 128 / 96 and
// accepts a few hundred bytes of spills.  Measured per 2^18 trajectories with 1 / 3 / 4 / 5 / 6 blocks per SM:
//   RADAU  VdP mu=1000 (n=2)   99.8 / 74.1 / 64.8 / 58.9 / 60.8 ms      BDF  VdP mu=1000 (n=2)   92.9 / 92.8 / 80.6 / 78.4 / 84.4 ms
//   RADAU  Robertson   (n=3)   12.5 / 10.8 / 11.7 / 11.8 / 13.1 ms      BDF  Robertson   (n=3)   62.3 / 46.9 / 46.3 / 57.0 / 60.2 ms
// The shared-memory variants (4 <= n <= 8) are limited by their matrices, not by registers.
template <class Prob, int METHOD, int FEAT>
__global__ void __launch_bounds__(ImplicitSel<Prob, METHOD, FEAT>::BLK, (implicit_min_blocks<Prob::N, METHOD>())) implicit_kernel(const __grid_constant__ KArgs a) {
  implicit_body<Prob, METHOD, FEAT>(a);
}

// n > IMPLICIT_MAX_N: one trajectory per warp, matrices in the warp's shared memory (ivpb_implicit_warp.cuh)
template <class Prob, int METHOD, int FEAT>
__global__ void __launch_bounds__(ImplicitWarpSel<Prob, METHOD, FEAT>::BLK, 1) implicit_warp_kernel(const __grid_constant__ KArgs a) {
  implicit_warp_body<Prob, METHOD, FEAT>(a);
}

template <class Prob, int METHOD>
__host__ inline const void* implicit_lookup_feat(int feat, int* block, int* smem, int* units) {
  if constexpr (Prob::N > IMPLICIT_MAX_N) {
    using Sel = ImplicitWarpSel<Prob, METHOD, 0>;
    if constexpr (!Sel::FITS) return nullptr;
    else {
      if (block) *block = Sel::BLK;
      if (smem) *smem = Sel::SMEM_BYTES;
      if (units) *units = -Sel::WARPS;      // negative: warp-per-trajectory, needs KArgs::scratch (one Jacobian per warp)
      switch (feat) {
        case 0: return (const void*)&implicit_warp_kernel<Prob, METHOD, 0>;
        case K_OUT: return (const void*)&implicit_warp_kernel<Prob, METHOD, K_OUT>;
        case K_OUT | K_EVENTS:
          if constexpr (Prob::NEV > 0) return (const void*)&implicit_warp_kernel<Prob, METHOD, K_OUT | K_EVENTS>;
          else return nullptr;
        case K_USER:                     // Options.user_solout: the problem's own SolOut (WarpHook, ivpb_erk.cuh)
          if constexpr (Prob::HAS_SOLOUT) return (const void*)&implicit_warp_kernel<Prob, METHOD, K_USER>;
          else return nullptr;
        default: return nullptr;
      }
    }
  } else {
    if (block) *block = ImplicitSel<Prob, METHOD, 0>::BLK;
    if (smem) *smem = ImplicitSel<Prob, METHOD, 0>::SMEM_BYTES;
    if (units) *units = ImplicitSel<Prob, METHOD, 0>::BLK;
    switch (feat) {
      case 0: return (const void*)&implicit_kernel<Prob, METHOD, 0>;
      case K_OUT: return (const void*)&implicit_kernel<Prob, METHOD, K_OUT>;
      case K_OUT | K_EVENTS:
        if constexpr (Prob::NEV > 0) return (const void*)&implicit_kernel<Prob, METHOD, K_OUT | K_EVENTS>;
        else return nullptr;
      case K_USER:                       // Options.user_solout: the problem's own SolOut
        if constexpr (Prob::HAS_SOLOUT) return (const void*)&implicit_kernel<Prob, METHOD, K_USER>;
        else return nullptr;
      default: return nullptr;
    }
  }
}

template <class Prob>
__host__ inline const void* implicit_lookup(int method, int feat, int* block, int* smem, int* units) {
  switch (method) {
    case M_RADAU: return implicit_lookup_feat<Prob, M_RADAU>(feat, block, smem, units);
    case M_BDF: return implicit_lookup_feat<Prob, M_BDF>(feat, block, smem, units);
    default: return nullptr;
  }
}
#endif

}  // namespace ivpb
