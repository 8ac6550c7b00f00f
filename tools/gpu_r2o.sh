#!/bin/bash
# round-2 (session 3): grouped guards (one branch per RHS / per error-norm pair) in the strict CR3BP kernel: parity tests + A/B
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
V=${1:-grp}
IVPB_LIB=ivp_b200/lib/libivpb_$V.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strict or bit_exact or golden or cr3bp or exact_div" > $O/r2o_pytest_$V.log 2>&1; tail -3 $O/r2o_pytest_$V.log
run() { # tag lib args...
  tag=$1; lib=$2; shift; shift
  IVPB_LIB=$lib python bench.py "$@" --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))" || tail -3 $O/$tag.err
}
for v in "" _$V; do
  run r2o_teval$v ivp_b200/lib/libivpb$v.so --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3
  run r2o_plain$v ivp_b200/lib/libivpb$v.so --workload cr3bp_dop853 --strict --trajectories 262144 --steps 3
done
