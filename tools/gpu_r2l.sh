#!/bin/bash
# round-2 (session 3): SolOut hooks in the warp kernels + the whole GPU suite
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_solout_hook.py -m gpu -q -x > $O/r2l_pytest_a.log 2>&1; tail -25 $O/r2l_pytest_a.log
timeout 2400 python -m pytest tests -m gpu -q > $O/r2l_pytest.log 2>&1; tail -8 $O/r2l_pytest.log
