// ivpb_nvrtc.cpp -- user problems given as CUDA C (`ivp_ode`, `ivp_events`, `ivp_jac`), compiled at run time
// with NVRTC together with the same solver templates the built-in problems use (the headers are embedded in
// the library by tools/embed_headers.py).  This is the device form of "implement the IVP trait"
// (reference src/ivp.rs:27-121): the user's functions are inlined into the persistent kernel.
//
// libnvrtc and libcuda are opened lazily with dlopen so that libivpb.so itself loads on a machine without a
// GPU driver (the CPU test-suite checks its exported symbols there).
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ivpb_runtime.h"
#include "ivpb_common.cuh"
#include "embedded_headers.inc"

namespace {

struct Api {
  bool ok = false;        // NVRTC + driver
  bool nvrtc_ok = false;  // NVRTC alone (enough to compile; used by the CPU-side compile check)
  std::string err;
  // NVRTC
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  const char* (*GetErrorString)(nvrtcResult);
  // driver
  CUresult (*ModuleLoadData)(CUmodule*, const void*);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**);
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t);
  CUresult (*GetErrorStringDrv)(CUresult, const char**);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
};

void* open_first(const std::vector<const char*>& names) {
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_LOCAL)) return h;
  return nullptr;
}

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* hn = open_first({"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"});
    void* hc = open_first({"libcuda.so.1", "libcuda.so"});
    if (!hn) { a.err = "cannot dlopen libnvrtc.so.12"; return; }
#define SYM(handle, field, name)                                                        \
  *(void**)(&a.field) = dlsym(handle, name);                                            \
  if (!a.field) { a.err = std::string("missing symbol ") + name; return; }
    SYM(hn, CreateProgram, "nvrtcCreateProgram") SYM(hn, CompileProgram, "nvrtcCompileProgram")
    SYM(hn, GetProgramLogSize, "nvrtcGetProgramLogSize") SYM(hn, GetProgramLog, "nvrtcGetProgramLog")
    SYM(hn, GetCUBINSize, "nvrtcGetCUBINSize") SYM(hn, GetCUBIN, "nvrtcGetCUBIN")
    SYM(hn, DestroyProgram, "nvrtcDestroyProgram") SYM(hn, GetErrorString, "nvrtcGetErrorString")
    a.nvrtc_ok = true;
    if (!hc) { a.err = "cannot dlopen libcuda.so.1 (no NVIDIA driver)"; return; }
    SYM(hc, ModuleLoadData, "cuModuleLoadData") SYM(hc, ModuleUnload, "cuModuleUnload")
    SYM(hc, ModuleGetFunction, "cuModuleGetFunction") SYM(hc, LaunchKernel, "cuLaunchKernel")
    SYM(hc, OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
    SYM(hc, GetErrorStringDrv, "cuGetErrorString") SYM(hc, FuncSetAttribute, "cuFuncSetAttribute")
#undef SYM
    a.ok = true;
  });
  return a;
}

struct Compiled { CUmodule mod = nullptr; CUfunction fn = nullptr; };
struct Cache {
  // key: device * 4096 + method * 64 + feat * 2 + strict
  std::map<int, Compiled> mods;
};

std::string drv_err(CUresult r) {
  const char* s = nullptr;
  api().GetErrorStringDrv(r, &s);
  return s ? s : "unknown driver error";
}

// The translation unit handed to NVRTC: solver headers, the user's functions, the problem adaptor, one kernel.
std::string program_source(const ivpb_user_problem& up, bool implicit) {
  std::string s;
  s += "#include \"ivpb_problems_min.cuh\"\n";
  s += implicit ? "#include \"ivpb_implicit_warp.cuh\"\n" : "#include \"ivpb_erk.cuh\"\n";
  s += "// ---- user source ----\n";
  s += up.src;
  s += "\n// ---- adaptor ----\n";
  s += "struct PUser : ivpb::ProblemDefaults<" + std::to_string(up.n) + ", " + std::to_string(up.p) + ", " +
       std::to_string(up.n_events) + "> {\n";
  if (up.n > 32)
    s += "  static constexpr bool HAS_ODE_I = true;\n"
         "  static __device__ __forceinline__ double ode_i(double t, const double* y, const double* p, int i) { return ivp_ode_i(t, y, p, i); }\n";
  else
    s += "  static __device__ __forceinline__ void ode(double t, const double* y, const double* p, double* d) { ivp_ode(t, y, p, d); }\n";
  if (up.n_events > 0)
    s += "  static __device__ __forceinline__ void events(double t, const double* y, const double* p, double* g) { ivp_events(t, y, p, g); }\n";
  if (up.has_jac & 4) {
    s += "  static constexpr bool HAS_SOLOUT = true;\n";
    s += "  template <class Interp, class Emit>\n"
         "  static __device__ __forceinline__ int solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit) {\n"
         "    return ivp_solout(xold, x, y, p, state, dense, emit);\n  }\n";
  }
  if (up.has_jac & 2) {
    s += "  static constexpr bool HAS_MASS = true;\n";
    s += "  static __device__ __forceinline__ void mass(const double* p, double* M) { ivp_mass(p, M); }\n";
  }
  if (up.has_jac & 1) {
    s += "  static constexpr bool HAS_JAC = true;\n";
    s += "  static __device__ __forceinline__ void jac(double t, const double* y, const double* p, double* J) { ivp_jac(t, y, p, J); }\n";
  }
  s += "};\n";
  if (implicit && up.n > 8) {
    s += "extern \"C\" __global__ void __launch_bounds__((ivpb::ImplicitWarpSel<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>::BLK), 1) "
         "ivpb_user_kernel(const __grid_constant__ ivpb::KArgs a) {\n";
    s += "  ivpb::implicit_warp_body<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>(a);\n}\n";
  } else if (implicit) {
    s += "extern \"C\" __global__ void __launch_bounds__((ivpb::ImplicitSel<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>::BLK), (ivpb::implicit_min_blocks<PUser::N, IVPB_USER_METHOD>())) "
         "ivpb_user_kernel(const __grid_constant__ ivpb::KArgs a) {\n";
    s += "  ivpb::implicit_body<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>(a);\n}\n";
  } else {
    s += "extern \"C\" __global__ void __launch_bounds__(IVPB_BLOCK) ivpb_user_kernel(const __grid_constant__ ivpb::KArgs a) {\n";
    s += up.n > 32 ? "  ivpb::erk_warp_body<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>(a);\n}\n"
                   : "  ivpb::erk_body<PUser, IVPB_USER_METHOD, IVPB_USER_FEAT>(a);\n}\n";
  }
  return s;
}

// NVRTC only (no driver needed): user problem -> sm_100a cubin.
int compile_cubin(const ivpb_user_problem& up, int method, int feat, int strict, std::vector<char>& cubin,
                  std::string& log) {
  Api& A = api();
  if (!A.nvrtc_ok) { log = "NVRTC unavailable: " + A.err; return IVPB_ERR_NVRTC; }
  const bool implicit = method >= 4;
  const std::string src = program_source(up, implicit);
  nvrtcProgram prog;
  nvrtcResult r = A.CreateProgram(&prog, src.c_str(), "ivpb_user_problem.cu", IVPB_N_HDR, IVPB_HDR_SRCS, IVPB_HDR_NAMES);
  if (r != NVRTC_SUCCESS) { log = std::string("nvrtcCreateProgram: ") + A.GetErrorString(r); return IVPB_ERR_NVRTC; }
  const std::string dm = "-DIVPB_USER_METHOD=" + std::to_string(method), df = "-DIVPB_USER_FEAT=" + std::to_string(feat);
  std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "-diag-suppress=177",
                                   dm.c_str(), df.c_str(), "-DIVPB_NVRTC=1"};
  if (strict) { opts.push_back("-fmad=false"); opts.push_back("-DIVPB_STRICT=1"); }
  r = A.CompileProgram(prog, (int)opts.size(), opts.data());
  size_t n = 0;
  A.GetProgramLogSize(prog, &n);
  log.assign(n, '\0');
  if (n) A.GetProgramLog(prog, &log[0]);
  if (r != NVRTC_SUCCESS) {
    A.DestroyProgram(&prog);
    log = std::string("NVRTC compilation of the user problem failed (") + A.GetErrorString(r) + "):\n" + log;
    return IVPB_ERR_NVRTC;
  }
  size_t sz = 0;
  A.GetCUBINSize(prog, &sz);
  cubin.resize(sz);
  A.GetCUBIN(prog, cubin.data());
  A.DestroyProgram(&prog);
  return 0;
}

}  // namespace

// Debug hook (not in include/ivpb.h): compile a user problem without a GPU; returns the cubin size or -rc.
extern "C" long long ivpb_debug_nvrtc_compile(const char* src, int n, int p, int n_events, int has_jac, int method,
                                              int feat, int strict, char* log, size_t log_cap) {
  ivpb_user_problem up;
  up.n = n; up.p = p; up.n_events = n_events; up.has_jac = has_jac; up.src = src;
  std::vector<char> cubin;
  std::string l;
  const int rc = compile_cubin(up, method, feat, strict, cubin, l);
  if (log && log_cap) { snprintf(log, log_cap, "%s", l.c_str()); }
  return rc ? -(long long)rc : (long long)cubin.size();
}

int ivpb_nvrtc_launch(ivpb_ctx* ctx, ivpb_user_problem& up, int device, int sms, int method, int feat, int strict,
                      const void* kargs, size_t kargs_bytes, long long N, int static_sched, long long max_warps,
                      cudaStream_t stream) {
  Api& A = api();
  if (!A.ok) { ivpb_set_error(ctx, "NVRTC path unavailable: " + A.err); return IVPB_ERR_NVRTC; }
  if (!up.impl) up.impl = new Cache();
  Cache* cache = (Cache*)up.impl;
  const int key = device * 4096 + method * 64 + feat * 2 + (strict ? 1 : 0);
  auto it = cache->mods.find(key);
  if (it == cache->mods.end()) {
    std::vector<char> cubin;
    std::string log;
    if (int rc = compile_cubin(up, method, feat, strict, cubin, log)) { ivpb_set_error(ctx, log); return rc; }
    Compiled c;
    CUresult cr = A.ModuleLoadData(&c.mod, cubin.data());
    if (cr != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuModuleLoadData: " + drv_err(cr)); return IVPB_ERR_CUDA; }
    cr = A.ModuleGetFunction(&c.fn, c.mod, "ivpb_user_kernel");
    if (cr != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuModuleGetFunction: " + drv_err(cr)); return IVPB_ERR_CUDA; }
    it = cache->mods.emplace(key, c).first;
  }
  // launch shape: explicit kernels 128 threads; implicit kernels follow ivpb::MatSel (ivpb_implicit.cuh)
  int block = 128;
  size_t smem = 0;
  long long units_per_block = 0;
  if (method >= 4 && up.n > 8) {
    // one trajectory per warp; must match ivpb::ImplicitWarpSel (ivpb_implicit_warp.cuh)
    const ivpb::WarpImplShape sh = ivpb::warp_impl_shape(up.n, method, (up.has_jac & 2) != 0);
    const size_t bytes = (size_t)sh.smem_doubles * 8;
    if (bytes > 227 * 1024) { ivpb_set_error(ctx, "implicit methods: the per-warp vectors of this state size do not fit shared memory"); return IVPB_ERR_CONFIG; }
    const int warps = sh.warps;
    block = 32 * warps; smem = bytes * warps; units_per_block = warps;
    CUresult ar = A.FuncSetAttribute(it->second.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
    if (ar != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuFuncSetAttribute: " + drv_err(ar)); return IVPB_ERR_CUDA; }
  } else if (method >= 4) {
    // must match ivpb::MatSel / RadauTraj / BdfTraj::SMEM_DOUBLES_PER_THREAD (ivpb_implicit.cuh)
    block = up.n <= 6 ? 128 : 64;
    size_t doubles = 0;
    // the Jacobian always lives in shared memory; the iteration matrices (+ RADAU's mass matrix) for n > 3
    if (up.n > 3) doubles += (size_t)(method == 4 ? ((up.has_jac & 2) ? 5 : 4) : 2) * up.n * up.n;
    else doubles += (size_t)up.n * up.n;
    if (method == 4) doubles += (size_t)4 * up.n;                           // RADAU: dense-output coefficients (cont)
    if (method == 5) doubles += (size_t)15 * up.n;                          // BDF: D (8 rows), scratch (6), Jacobian point
    smem = doubles * block * 8;
    if (smem > 0) {
      CUresult ar = A.FuncSetAttribute(it->second.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
      if (ar != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuFuncSetAttribute: " + drv_err(ar)); return IVPB_ERR_CUDA; }
    }
  }
  if (units_per_block == 0) units_per_block = block;
  if (method < 4 && up.n > 32) {       // one trajectory per warp: WarpLayout needs 2 n doubles of shared memory per warp
    smem = (size_t)(block / 32) * 2 * up.n * 8;
    units_per_block = block / 32;
    if (smem > 48 * 1024) {
      CUresult ar = A.FuncSetAttribute(it->second.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem);
      if (ar != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuFuncSetAttribute: " + drv_err(ar)); return IVPB_ERR_CUDA; }
    }
  }
  int occ = 1;
  A.OccupancyMaxActiveBlocksPerMultiprocessor(&occ, it->second.fn, block, smem);
  if (occ < 1) occ = 1;
  long long grid = (long long)sms * occ;
  const long long need = (N + units_per_block - 1) / units_per_block;
  if (static_sched || need < grid) grid = need;
  if (max_warps > 0 && method >= 4 && up.n > 8) grid = std::max<long long>(1, std::min<long long>(grid, max_warps / units_per_block));
  std::vector<char> copy((const char*)kargs, (const char*)kargs + kargs_bytes);
  void* params[] = {copy.data()};
  CUresult cr = A.LaunchKernel(it->second.fn, (unsigned)grid, 1, 1, block, 1, 1, (unsigned)smem, (CUstream)stream, params, nullptr);
  if (cr != CUDA_SUCCESS) { ivpb_set_error(ctx, "cuLaunchKernel: " + drv_err(cr)); return IVPB_ERR_CUDA; }
  return 0;
}

void ivpb_nvrtc_release(ivpb_user_problem& up) {
  if (!up.impl) return;
  Cache* cache = (Cache*)up.impl;
  if (api().ok)
    for (auto& kv : cache->mods)
      if (kv.second.mod) api().ModuleUnload(kv.second.mod);
  delete cache;
  up.impl = nullptr;
}
