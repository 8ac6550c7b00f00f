"""CPU-side checks: the C-ABI library loads and exports every symbol include/ivpb.h declares (no compute
calls without a GPU), the ctypes structs match the header, and the host mirror behaves like the reference's
types (names, defaults, conversions)."""
import ctypes
import os
import re

import numpy as np
import pytest

import ivp_b200 as ib
from ivp_b200 import _abi, api, synth
from ivp_b200.types import Method, Options, Status

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ivpb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ivpb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    syms = header_symbols()
    assert set(syms) == set(api.ABI_SYMBOLS)
    for s in syms:
        assert hasattr(lib, s), f"libivpb.so does not export {s}"
    assert b"sm_100a" in lib.ivpb_version()


def test_struct_layout_matches_header():
    # field order of the ctypes mirrors == field order in include/ivpb.h
    src = open(os.path.join(ROOT, "include", "ivpb.h")).read()
    body = src[src.index("typedef struct {", src.index("Per-trajectory outputs")):]
    body = body[: body.index("} ivpb_outputs;")]
    names = re.findall(r"\*\s*(\w+);", body)
    assert names == _abi.OUTPUT_FIELDS
    assert ctypes.sizeof(_abi.IvpbOutputs) == 8 * len(names)
    obody = src[src.index("typedef struct {", src.index("Mirrors `Options`")):]
    obody = re.sub(r"/\*.*?\*/", "", obody[: obody.index("} ivpb_options;")], flags=re.S)
    onames = []
    for decl in obody.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        parts = decl.split(",")
        first = parts[0].split()[-1].lstrip("*")
        onames.append(first)
        onames += [p.strip().lstrip("*") for p in parts[1:]]
    assert onames == [f for f, _ in _abi.IvpbOptions._fields_]


def test_builtin_problem_dims_without_gpu(oracle):
    for name, pid in api.PROBLEMS.items():
        p = api.Problem.builtin(name)
        assert (p.n, p.p, p.n_events) == oracle.dims(pid), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.Context()
    with pytest.raises(RuntimeError):
        ib.solve_ivp_batch("sho", 0.0, 1.0, np.array([[1.0, 0.0]]), None, Options())


def test_method_and_status_mirror_reference_enums():
    # reference src/solve/options.rs:14-27,61-73 and src/status.rs:4-19
    assert [m.name for m in Method] == ["RK23", "DOPRI5", "DOP853", "RK4", "RADAU", "BDF"]
    assert Method.from_str("rk45") == Method.DOPRI5 and Method.from_str("Radau5") == Method.RADAU
    assert Method.from_str("BDF15") == Method.BDF and Method.from_str("nonsense") == Method.DOPRI5
    assert [s.name for s in Status] == ["Success", "UserInterrupt", "NeedLargerNMax", "StepSizeTooSmall",
                                       "ProbablyStiff", "SingularMatrix", "PoorConvergence"]
    assert Status.UserInterrupt.is_success() and not Status.ProbablyStiff.is_success()
    assert [Method(m).coeffs_per_state() for m in range(6)] == [4, 5, 8, 4, 4, 7]


def test_options_defaults_and_builder():
    # reference src/solve/options.rs:75-123
    o = Options()
    assert o.method == Method.DOPRI5 and o.rtol == 1e-3 and o.atol == 1e-6 and o.dense_output is False
    assert o.max_steps is None and o.t_eval is None and o.first_step is None and o.max_step is None
    b = Options.builder().method("DOP853").rtol(1e-9).atol([1e-9, 1e-8]).t_eval([0.0, 1.0]).build()
    assert b.method == Method.DOP853 and b.t_eval == [0.0, 1.0]
    mo = _abi.MarshalledOptions(b, 2, 0)
    assert mo.struct.n_rtol == 1 and mo.struct.n_atol == 2 and mo.struct.has_t_eval == 1 and mo.cap == 3
    with pytest.raises(ValueError):
        _abi.MarshalledOptions(Options(rtol=[1, 2, 3]), 2, 0)


def test_synth_is_counter_based_and_shardable():
    a = synth.uniform(1000, 3)
    b = np.concatenate([synth.uniform(400, 3), synth.uniform(600, 3, offset=400)])
    assert np.array_equal(a, b) and a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.02
    # splitmix64 known answer (seed 0 -> first output of the reference splitmix64 stream)
    assert int(synth.splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF
    for wl in ("vdp", "decay", "lorenz", "cr3bp", "ball", "robertson", "vdp_stiff"):
        prob, y0, par, t0, tf = synth.ensemble(wl, 7)
        p = api.Problem.builtin(prob)
        assert y0.shape == (7, p.n) and (par is None or par.shape == (7, p.p))


def test_batch_solution_views():
    from ivp_b200.types import BatchSolution
    N, n = 3, 2
    b = BatchSolution(n=n, n_events=1, status=np.array([0, 1, 0], dtype=np.int32),
                      counters=np.arange(18, dtype=np.uint32).reshape(3, 6), t_final=np.array([1.0, 0.5, 1.0]),
                      y_final=np.ones((N, n)), n_out=np.array([2, 3, 1], dtype=np.int32),
                      t_out=np.arange(6.0).reshape(3, 2), y_out=np.ones((3, 2, 2)),
                      ev_count=np.array([[0], [1], [0]], dtype=np.int32), ev_t=np.full((3, 1, 2), 0.5),
                      ev_y=np.zeros((3, 1, 2, 2)))
    s = b.solution(1)
    assert s.status == Status.UserInterrupt and s.truncated and len(s.t) == 2 and len(s.t_events[0]) == 1
    assert (s.nfev, s.njev, s.nlu, s.nstep, s.naccpt, s.nrejct) == (6, 7, 8, 9, 10, 11)
    assert not b.solution(0).truncated and len(b.solutions()) == 3


USER_VDP = r"""
// user problem in CUDA C: the device form of `impl IVP for VanDerPol` (reference src/ivp.rs:27-53)
__device__ void ivp_ode(double t, const double* y, const double* p, double* dydt) {
  dydt[0] = y[1];
  dydt[1] = p[0] * (1.0 - y[0] * y[0]) * y[1] - y[0];
}
__device__ void ivp_events(double t, const double* y, const double* p, double* g) { g[0] = y[0]; }
"""


def _nvrtc_compile(src, n, p, nev, has_jac, method, feat, strict=0):
    lib = api.load_library()
    lib.ivpb_debug_nvrtc_compile.restype = ctypes.c_longlong
    log = ctypes.create_string_buffer(1 << 16)
    r = lib.ivpb_debug_nvrtc_compile(src.encode(), n, p, nev, has_jac, method, feat, strict, log, 1 << 16)
    return r, log.value.decode()


def test_nvrtc_compiles_user_problem_for_sm100a_without_gpu():
    """The NVRTC half of the user-problem path needs no driver: solver headers (embedded in libivpb.so) +
    user CUDA C -> sm_100a cubin, for every explicit method and feature set, fast and strict."""
    for method in (0, 1, 2, 3):
        for feat, nev in ((0, 0), (1, 0), (3, 1)):
            for strict in (0, 1):
                size, log = _nvrtc_compile(USER_VDP, 2, 1, nev, 0, method, feat, strict)
                assert size > 10000, log


def test_nvrtc_compiles_implicit_kernels_without_gpu():
    """RADAU / BDF for user problems: finite-difference Jacobian, and ivp_jac when has_jac; n = 5 uses the
    shared-memory matrix storage."""
    jac = USER_VDP + """
__device__ void ivp_jac(double t, const double* y, const double* p, double* J) {
  J[0] = 0.0; J[1] = 1.0; J[2] = -2.0 * p[0] * y[0] * y[1] - 1.0; J[3] = p[0] * (1.0 - y[0] * y[0]);
}
"""
    for method in (4, 5):
        size, log = _nvrtc_compile(jac, 2, 1, 1, 1, method, 3, 1)
        assert size > 10000, log
        size, log = _nvrtc_compile(USER_VDP, 2, 1, 0, 0, method, 0, 0)
        assert size > 10000, log
    five = "__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { for (int i = 0; i < 5; ++i) d[i] = -y[(i + 1) % 5] * y[i]; }"
    size, log = _nvrtc_compile(five, 5, 0, 0, 0, 4, 1, 0)
    assert size > 10000, log


def test_nvrtc_reports_compile_errors():
    size, log = _nvrtc_compile("__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { d[0] = undefined_symbol; }",
                               1, 0, 0, 0, 2, 0)
    assert size == -3 and "undefined_symbol" in log          # IVPB_ERR_NVRTC
    size, log = _nvrtc_compile("// no ivp_ode at all", 1, 0, 0, 0, 2, 0)
    assert size == -3 and "ivp_ode" in log


def test_libm_pow_transcription_matches_host_libm(tmp_path):
    """ivpb_libm_pow.cuh (glibc's pow, operation for operation; compiled here for the host) must be bit-identical
    to the pow of the libm the oracle -- and a Rust build of the reference -- calls on this machine."""
    import subprocess
    src = tmp_path / "powcheck.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cmath>
#include <cstring>
#include <random>
#include "ivpb_libm_pow.cuh"
int main() {
  std::mt19937_64 g(2024);
  std::uniform_real_distribution<double> U(0, 1);
  const double ys[] = {0.125, 0.2, 0.17, 0.04, 0.25, 0.8, -1.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0, -0.5, -0.25, -1.0 / 6.0, 4.0, 1.0};
  long bad = 0;
  for (long it = 0; it < 6000000; ++it) {
    double x, y;
    switch (it % 4) {
      case 0: x = std::exp(40 * (U(g) - 0.5)); y = ys[g() % (sizeof(ys) / sizeof(ys[0]))]; break;
      case 1: x = std::exp(20 * (U(g) - 0.5)); y = 8 * (U(g) - 0.5); break;
      case 2: x = 2 * U(g); y = ys[g() % (sizeof(ys) / sizeof(ys[0]))]; break;
      default: x = std::exp(600 * (U(g) - 0.5)); y = 2 * (U(g) - 0.5);
    }
    const double a = std::pow(x, y), b = ivpb_libm_pow(x, y);
    if (std::memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) ++bad;
  }
  const double sx[] = {0.0, -1.0, 1.0, INFINITY, 1e-310, 4.9e-324, 2.2e-308, 2.0, 1e300, 1e-300, NAN};
  const double sy[] = {0.125, 0.5, 2.0, -0.25, 1e-70, INFINITY, 5.0, -5.0, 0.0};
  for (double x : sx) for (double y : sy) {
    const double a = std::pow(x, y), b = ivpb_libm_pow(x, y);
    if (std::memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) ++bad;
  }
  std::printf("%ld\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "powcheck"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma",
                           "-I", os.path.join(ROOT, "ivp_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout


def test_scipy_style_front_end_argument_handling():
    """ivp_b200.scipy_api.solve_ivp (reference src/python/solve.rs:153-222): everything that can be checked without a GPU."""
    from ivp_b200 import scipy_api
    with pytest.raises(TypeError, match="callable cannot run on the device"):
        scipy_api.solve_ivp(lambda t, y: -y, (0, 1), [1.0])
    with pytest.raises(TypeError, match="unknown options"):
        scipy_api.solve_ivp("decay", (0, 1), [1.0], args=(0.5,), bogus=1)
    with pytest.raises(ValueError, match="parameters"):
        scipy_api.solve_ivp("decay", (0, 1), [1.0])
    with pytest.raises(ValueError, match="event specs"):
        scipy_api.solve_ivp("sho", (0, 1), [1.0, 0.0], events=[object(), object()])

    class Ev:
        terminal, direction = True, -1.0
    cfgs, has = scipy_api._event_configs(Ev(), 1)
    assert has and cfgs[0].terminal_count == 1 and int(cfgs[0].direction) < 0
    cfgs, has = scipy_api._event_configs(None, 1)
    assert not has and cfgs[0].terminal_count is None


def test_options_struct_layout_matches_header():
    """Field order of the ctypes `ivpb_options` mirror == include/ivpb.h (fields were appended for mass matrices and hooks)."""
    src = open(os.path.join(ROOT, "include", "ivpb.h")).read()
    body = src[src.index("typedef struct {", src.index("Mirrors `Options`")):]
    body = re.sub(r"/\*.*?\*/", "", body[: body.index("} ivpb_options;")], flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        first, *rest = decl.split(",")
        names.append(re.findall(r"(\w+)\s*$", first.strip())[0])
        names += [r.strip().lstrip("*") for r in rest]
    assert names == [f[0] for f in _abi.IvpbOptions._fields_]


def test_new_option_fields_are_marshalled():
    from ivp_b200.api import IVPB_FLAG_NO_SORT, IVPB_FLAG_SORT
    o = _abi.MarshalledOptions(Options(method=Method.RADAU, mass_storage="Full", nind2=1, user_solout=True,
                                       flags=IVPB_FLAG_NO_SORT), 3, 0).struct
    assert (o.mass_storage, o.nind1, o.nind2, o.nind3, o.user_solout, o.flags) == (1, -1, 1, -1, 1, 16)
    d = _abi.MarshalledOptions(Options(), 3, 0).struct
    assert (d.mass_storage, d.nind1, d.nind2, d.nind3, d.user_solout) == (0, -1, -1, -1, 0) and IVPB_FLAG_SORT == 32
    with pytest.raises(ValueError, match="mass_storage"):
        _abi.MarshalledOptions(Options(mass_storage="Banded"), 3, 0)


def test_nvrtc_compiles_mass_matrix_and_solout_hooks_without_gpu():
    """has_jac bit 1 (`ivp_mass`) and bit 2 (`ivp_solout`): the adaptor and the K_USER / mass kernels compile for sm_100a."""
    src = USER_VDP + """
__device__ void ivp_mass(const double* p, double* M) { M[0] = 2.0; M[1] = 0.5; M[2] = 0.0; M[3] = 1.0; }
template <class Interp, class Emit>
__device__ int ivp_solout(double xold, double& x, double* y, const double* p, double* state, const Interp& dense, Emit& emit) {
  if (!dense.valid()) { emit(x, y); return 0; }
  double yi[2];
  dense.eval(0.5 * (xold + x), yi);
  if (yi[0] < 0.0) { y[0] = -y[0]; state[0] += 1.0; return 2; }
  return y[0] > 10.0 ? 1 : 0;
}
"""
    for method in (0, 1, 2, 3, 4, 5):                      # feature 4 = K_USER: the hook in every solver
        size, log = _nvrtc_compile(src, 2, 1, 0, 4, method, 4, 1 if method >= 4 else 0)
        assert size > 10000, log
    for feat in (0, 1):                                     # RADAU with the mass matrix (strict and FMA builds)
        for strict in (0, 1):
            size, log = _nvrtc_compile(src, 2, 1, 0, 2, 4, feat, strict)
            assert size > 10000, log
    five = ("__device__ void ivp_ode(double t, const double* y, const double* p, double* d) { for (int i = 0; i < 5; ++i) d[i] = -y[i]; }\n"
            "__device__ void ivp_mass(const double* p, double* M) { for (int i = 0; i < 25; ++i) M[i] = (i % 6 == 0) ? 1.0 + i : 0.1; }")
    size, log = _nvrtc_compile(five, 5, 0, 0, 2, 4, 1, 1)   # n = 5: the mass matrix is the fifth shared-memory matrix
    assert size > 10000, log


def _greedy_groups(S):
    """Python restatement of group_columns (reference src/python/sparsity.rs:109-154)."""
    n = S.shape[0]
    groups, used = [-1] * n, []
    for col in range(n):
        rows = np.nonzero(S[:, col])[0]
        for g, u in enumerate(used):
            if not u[rows].any():
                groups[col] = g
                u[rows] = True
                break
        else:
            u = np.zeros(n, dtype=bool)
            u[rows] = True
            used.append(u)
            groups[col] = len(used) - 1
    return groups, len(used)


def test_jac_sparsity_marshalling_and_column_groups(oracle):
    """Options.jac_sparsity -> compressed columns (include/ivpb.h), and the runtime's column grouping == the oracle's ==
    a Python restatement of the reference's greedy first-fit rule, on the MEDAKZO structure and on random structures."""
    lib = api.load_library()
    olib = oracle.lib()
    rng = np.random.default_rng(3)
    cases = [synth.medakzo_sparsity(200), synth.medakzo_sparsity(5), np.eye(7, dtype=np.int8), np.ones((6, 6), dtype=np.int8),
             np.zeros((4, 4), dtype=np.int8)] + [(rng.random((n, n)) < d).astype(np.int8) for n, d in ((40, 0.05), (64, 0.1), (9, 0.4))]
    for S in cases:
        n = S.shape[0]
        mo = _abi.MarshalledOptions(Options(method=Method.RADAU, jac_sparsity=S), n, 0)
        o = mo.struct
        assert o.has_jac_sparsity == 1
        colptr = np.ctypeslib.as_array(o.jac_sparsity_colptr, shape=(n + 1,))
        assert colptr[0] == 0 and colptr[n] == int(S.sum())
        rows = np.ctypeslib.as_array(o.jac_sparsity_rows, shape=(max(int(colptr[n]), 1),))
        for c in range(n):
            assert np.array_equal(rows[colptr[c]:colptr[c + 1]], np.nonzero(S[:, c])[0])
        g_rt = np.zeros(n, dtype=np.int32)
        g_or = np.zeros(n, dtype=np.int32)
        ng_rt = lib.ivpb_debug_group_columns(n, _abi.ptr(mo.sp_colptr), _abi.ptr(mo.sp_rows), _abi.ptr(g_rt))
        ng_or = olib.oracle_group_columns(n, _abi.ptr(mo.sp_colptr), _abi.ptr(mo.sp_rows), _abi.ptr(g_or))
        g_py, ng_py = _greedy_groups(S)
        assert ng_rt == ng_or == ng_py and list(g_rt) == list(g_or) == g_py
    g, ng = _greedy_groups(synth.medakzo_sparsity(200))
    assert ng <= 6          # 400 columns, a handful of RHS evaluations per Jacobian
    assert _abi.MarshalledOptions(Options(), 3, 0).struct.has_jac_sparsity == 0
    with pytest.raises(ValueError):
        _abi.MarshalledOptions(Options(jac_sparsity=np.ones((2, 3))), 3, 0)


def test_nvrtc_compiles_medakzo_400_for_the_warp_cooperative_kernels():
    """n = 400 (the reference's own MEDAKZO size): the warp-cooperative RADAU kernel compiles for sm_100a with its iteration
    matrices in global memory (warp_impl_shape says so; shared memory holds 6 n doubles per warp)."""
    size, log = _nvrtc_compile(synth.medakzo_cuda_source(200), 400, 0, 0, 0, 4, 0, 1)
    assert size > 10000, log
