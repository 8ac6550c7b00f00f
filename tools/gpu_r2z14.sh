#!/bin/bash
# end of round 2: ncu --set full of the north-star kernel (third erk_kernel launch: the first two are the parity pilot's)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:erk_kernel -s 2 -c 1 -o $O/r2z14_vdp_dop853 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2z14_vdp_dop853.log 2>&1
python tools/ncu_summary.py $O/r2z14_vdp_dop853.ncu-rep $O/r2z14_vdp_dop853_ncu_full.txt > /dev/null 2>&1
grep -E "Kernel Name|duration|grid_size|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|dram__bytes" $O/r2z14_vdp_dop853_ncu_full.txt | cut -c1-150
