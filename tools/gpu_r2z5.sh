#!/bin/bash
# round 2, session 4: ncu source-level captures of the RADAU and the new BDF kernel (VdP mu=1000, 131072 trajectories)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
cap() { tag=$1; k=$2; skip=$3; shift; shift; shift
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$tag -f python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline > $O/$tag.log 2>&1
  python tools/ncu_summary.py $O/$tag.ncu-rep $O/${tag}_ncu_full.txt > /dev/null 2>&1
  grep -E "Kernel Name|duration|registers_per|issue_active|thread_inst_executed_per|pipe_fp64_cycles|no_instruction|stalled_wait|inst_executed.sum" $O/${tag}_ncu_full.txt | cut -c1-150; }
cap r2z5_vdpstiff_bdf implicit_kernel 2 --workload vdpstiff_bdf --trajectories 131072
cap r2z5_vdpstiff_radau implicit_kernel 2 --workload vdpstiff_radau --trajectories 131072
