// ivpb_inst.cu -- instantiates every kernel of ONE built-in problem.  Built once per problem and per
// floating-point mode:
//   nvcc -DIVPB_PROBLEM=PVdpMu -DIVPB_TAG=vdp_mu                      (FMA contraction on: the fast path)
//   nvcc -DIVPB_PROBLEM=PVdpMu -DIVPB_TAG=vdp_mu -DIVPB_STRICT -fmad=false
// The strict build renames the namespace so the two sets of template instantiations cannot be merged
// by the linker.
#if defined(IVPB_STRICT) && defined(IVPB_DEFER_GUARDS)
#define ivpb ivpb_strictd      // strict arithmetic with deferred division / square-root guards (ivpb_exact.cuh)
#elif defined(IVPB_STRICT)
#define ivpb ivpb_strict
#endif
#include "ivpb_problems.cuh"
#include "ivpb_kernels.cuh"

#define IVPB_CAT2(a, b) a##b
#define IVPB_CAT(a, b) IVPB_CAT2(a, b)
#if defined(IVPB_STRICT) && defined(IVPB_DEFER_GUARDS)
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_strictd_, tag)
#elif defined(IVPB_STRICT)
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_strict_, tag)
#else
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_, tag)
#endif

#include "ivpb_runtime.h"

// The explicit kernels keep the guarded divisions (grouped per right-hand side / per error-norm pair, ivpb_exact.cuh).
// Their strictd twins are not built: they have few divisions per step to gain from, and the round-2 attempt did not
// return on ensembles of 2^18 trajectories and more (not diagnosed).  -DIVPB_ERK_STRICTD=1 builds them (A/B).
#ifndef IVPB_ERK_STRICTD
#define IVPB_ERK_STRICTD 0
#endif

extern "C" const void* IVPB_SYM(IVPB_TAG)(int method, int feat, ivpb_pinfo* info) {
  using P = ivpb::IVPB_PROBLEM;
  if (info) {
    info->n = P::N; info->p = P::P; info->nev = P::NEV; info->has_jac = (P::HAS_JAC ? 1 : 0) | (P::HAS_MASS ? 2 : 0) | (P::HAS_SOLOUT ? 4 : 0);
    for (int e = 0; e < 8; ++e) {
      info->ev_dir[e] = e < P::NEV ? P::default_dir(e) : 0;
      info->ev_term[e] = e < P::NEV ? (long long)P::default_term(e) : -1;
    }
  }
  if (method < 0) return nullptr;
#ifdef IVPB_DEFER_GUARDS
  // thread-per-trajectory kernels with DefaultSolOut or none: the warp kernels and the hooks keep the guarded build
  if constexpr (P::N > ivpb::MAX_N || !IVPB_ERK_STRICTD) return nullptr;
  else {
    if (feat & ivpb::K_USER) return nullptr;
    if (info) info->block = ivpb::erk_lookup_block<P>(method);
    return ivpb::erk_lookup<P>(method, feat);
  }
#else
  if (info) info->block = ivpb::erk_lookup_block<P>(method);
  return ivpb::erk_lookup<P>(method, feat);
#endif
}
