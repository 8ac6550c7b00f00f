// ivpb_inst_implicit.cu -- instantiates the RADAU / BDF kernels of ONE built-in problem (same scheme as
// ivpb_inst.cu: one object per problem and per floating-point mode, -DIVPB_PROBLEM / -DIVPB_TAG).
#if defined(IVPB_STRICT) && defined(IVPB_DEFER_GUARDS)
#define ivpb ivpb_strictd      // strict arithmetic with deferred division / square-root guards (ivpb_exact.cuh)
#elif defined(IVPB_STRICT)
#define ivpb ivpb_strict
#endif
#define IVPB_WITH_IMPLICIT 1
#include "ivpb_problems.cuh"
#include "ivpb_kernels.cuh"

#define IVPB_CAT2(a, b) a##b
#define IVPB_CAT(a, b) IVPB_CAT2(a, b)
#if defined(IVPB_STRICT) && defined(IVPB_DEFER_GUARDS)
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_impl_strictd_, tag)
#elif defined(IVPB_STRICT)
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_impl_strict_, tag)
#else
#define IVPB_SYM(tag) IVPB_CAT(ivpb_lookup_impl_, tag)
#endif

extern "C" const void* IVPB_SYM(IVPB_TAG)(int method, int feat, int* block, int* smem, int* units) {
#ifdef IVPB_DEFER_GUARDS
  if constexpr (ivpb::IVPB_PROBLEM::N > ivpb::IMPLICIT_MAX_N) return nullptr;      // warp kernels, hooks: guarded build
  else {
    if (feat & ivpb::K_USER) return nullptr;
    return ivpb::implicit_lookup<ivpb::IVPB_PROBLEM>(method, feat, block, smem, units);
  }
#else
  return ivpb::implicit_lookup<ivpb::IVPB_PROBLEM>(method, feat, block, smem, units);
#endif
}
