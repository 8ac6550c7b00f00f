#!/bin/bash
# round-2 (session 3): strict CR3BP A/B: out-of-line RHS, with / without block-synchronous trips, 2 blocks per SM; ncu of the base
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
run() { # tag lib args...
  tag=$1; lib=$2; shift; shift
  IVPB_LIB=$lib python bench.py "$@" --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))" || tail -3 $O/$tag.err
}
for v in "" _noinl _noinl_nosync; do
  run r2m_teval$v ivp_b200/lib/libivpb$v.so --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3
  run r2m_plain$v ivp_b200/lib/libivpb$v.so --workload cr3bp_dop853 --strict --trajectories 262144 --steps 3
done
cap() { # tag, kernel regex, bench args...
  tag=$1; k=$2; shift; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/$tag -f python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline > $O/$tag.log 2>&1
  python tools/ncu_summary.py $O/$tag.ncu-rep $O/${tag}_ncu_full.txt > /dev/null 2>&1
  grep -E "duration|grid_size|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|no_instruction|stalled_wait|scoreboard|barrier|dram__bytes" $O/${tag}_ncu_full.txt
}
cap r2m_cr3bp_teval_strict erk_kernel --workload cr3bp_dop853_teval --strict --trajectories 262144
