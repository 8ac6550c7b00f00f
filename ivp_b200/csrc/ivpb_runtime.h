// ivpb_runtime.h -- internal declarations shared by the runtime translation units (not installed).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/ivpb.h"

#define IVPB_USER_HANDLE_BASE 1000

// What the runtime needs to know about a problem (filled by the per-problem lookup of ivpb_inst.cu).
struct ivpb_pinfo {
  int block;                // threads per block of the kernel the lookup returned (explicit methods)
  int n, p, nev, has_jac;   // has_jac: bit 0 = analytic Jacobian, bit 1 = mass matrix (IVP::mass), bit 2 = own SolOut hook
  int ev_dir[8];          // IVP::event_config defaults
  long long ev_term[8];
};

// A user problem given as CUDA C (ivpb_nvrtc_problem); compiled lazily per (method, feature set).
struct ivpb_user_problem {
  int n = 0, p = 0, n_events = 0, has_jac = 0;
  std::string src;
  void* impl = nullptr;   // NVRTC module cache, owned by ivpb_nvrtc.cpp
};

int ivpb_nvrtc_launch(ivpb_ctx* ctx, ivpb_user_problem& up, int device, int sms, int method, int feat, int strict,
                      const void* kargs, size_t kargs_bytes, long long N, int static_sched, long long max_warps, cudaStream_t stream);
void ivpb_nvrtc_release(ivpb_user_problem& up);
void ivpb_set_error(ivpb_ctx* ctx, const std::string& msg);
