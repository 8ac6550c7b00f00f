"""The C++ host API (include/ivp_batch.hpp, the stand-in for the Rust `ivp-batch` crate -- no rustc here):
builds tests/cpp/test_ivp_batch.cpp against libivpb.so.  CPU: the API-surface checks and the loud failure
without a device.  GPU: the reference's own integration tests restated against solve_ivp / solve_ivp_batch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from ivp_b200 import api
    api.load_library()      # builds nothing; fails loudly if libivpb.so is missing
    out = tmp_path_factory.mktemp("cpp") / "test_ivp_batch"
    libdir = os.path.join(ROOT, "ivp_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_ivp_batch.cpp"), "-o", str(out),
                           "-L", libdir, "-livpb", f"-Wl,-rpath,{libdir}"])
    return str(out)


def test_cpp_api_surface_and_no_cpu_fallback(exe):
    import torch
    args = [exe, "--api-only"] + ([] if torch.cuda.is_available() else ["--expect-no-device"])
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_reference_integration_tests(exe):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout
