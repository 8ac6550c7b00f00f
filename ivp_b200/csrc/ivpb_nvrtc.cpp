// ivpb_nvrtc.cpp -- user problems given as CUDA C, compiled with NVRTC together with the solver headers.
// (placeholder: the NVRTC path lands after the built-in path is measured)
#include "ivpb_runtime.h"

int ivpb_nvrtc_launch(ivpb_ctx* ctx, ivpb_user_problem&, int, int, int, int, const void*, size_t, cudaStream_t) {
  ivpb_set_error(ctx, "NVRTC user problems are not available in this build");
  return IVPB_ERR_NVRTC;
}
void ivpb_nvrtc_release(ivpb_user_problem&) {}
