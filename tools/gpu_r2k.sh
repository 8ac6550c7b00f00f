#!/bin/bash
# round-2 (session 3): large-n / sparse-FD tests on the warp-cooperative implicit kernels, then the rest of the warp tests
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --durations=8 -k "radau_n100 or sparse_fd or medakzo_400" > $O/r2k_pytest_a.log 2>&1; tail -25 $O/r2k_pytest_a.log
timeout 900 python -m pytest tests -m gpu -q -k "warp or medakzo or mass or linear100 or nvrtc" > $O/r2k_pytest_b.log 2>&1; tail -6 $O/r2k_pytest_b.log
