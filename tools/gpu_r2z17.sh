#!/bin/bash
# final check of the round: the full GPU suite on the last build
cd "$GRAFT_REPO_ROOT"
timeout 118 python -m pytest tests -m gpu -q -x > gpurun_out/r2z17_pytest.log 2>&1; tail -2 gpurun_out/r2z17_pytest.log
