#!/bin/bash
# round 2, session 4: source-level ncu capture of the Lorenz DOPRI5 kernel (config 2)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:erk_kernel -s 4 -c 1 -o $O/r2z12_lorenz_dopri5 -f python bench.py --workload lorenz_dopri5 --steps 1 --warmup 1 --no-cpu-baseline --trajectories 262144 > $O/r2z12_lorenz.log 2>&1
python tools/ncu_summary.py $O/r2z12_lorenz_dopri5.ncu-rep $O/r2z12_lorenz_dopri5_ncu_full.txt > /dev/null 2>&1
grep -E "Kernel Name|duration|grid_size|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|stalled" $O/r2z12_lorenz_dopri5_ncu_full.txt | cut -c1-150
