#!/bin/bash
# round-2 (session 3): ncu --set full of the strict CR3BP kernel and the strict VdP mu=1000 RADAU / BDF kernels as they stand
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
cap() { # tag, kernel regex, bench args...
  tag=$1; k=$2; shift; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/$tag -f python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline > $O/$tag.log 2>&1
  python tools/ncu_summary.py $O/$tag.ncu-rep $O/${tag}_ncu_full.txt > /dev/null 2>&1
  grep -E "duration|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|no_instruction|stalled_wait|dram__bytes" $O/${tag}_ncu_full.txt
}
cap r2i_cr3bp_teval_strict erk_kernel --workload cr3bp_dop853_teval --trajectories 262144
cap r2i_vdpstiff_radau implicit_kernel --workload vdpstiff_radau --trajectories 131072
cap r2i_vdpstiff_bdf implicit_kernel --workload vdpstiff_bdf --trajectories 131072
ls -la $O | grep r2i
