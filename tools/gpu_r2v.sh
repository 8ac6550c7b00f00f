#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
IVPB_LIB=ivp_b200/lib/libivpb_gdbg.so timeout 100 python - > $O/r2v_guard_debug.log 2>&1 <<PY
import numpy as np, ivp_b200 as ib
from ivp_b200 import Method, Options, synth
prob, y0, par, t0, tf = synth.ensemble("robertson", 256)
g = ib.solve_ivp_batch(prob, t0, tf, y0, par, Options(method=Method.BDF, rtol=1e-6, atol=1e-6))
print("status", np.unique(g.status), "reruns", ib.api.default_context().last_reruns())
PY
head -12 $O/r2v_guard_debug.log
# north-star kernel, ncu --set full (flag given: no parity pilot in front of it)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:erk_kernel -s 1 -c 1 -o $O/r2_vdp_dop853 -f python bench.py --fast --steps 1 --warmup 1 --no-cpu-baseline > $O/r2v_ncu.log 2>&1
python tools/ncu_summary.py $O/r2_vdp_dop853.ncu-rep $O/r2_vdp_dop853_ncu_full.txt vdp_dop853 > /dev/null 2>&1
grep -E "Kernel Name|duration|grid_size|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|no_instruction|stalled_wait|dram__bytes" $O/r2_vdp_dop853_ncu_full.txt | cut -c1-150
