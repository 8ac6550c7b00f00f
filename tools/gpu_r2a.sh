#!/bin/bash
# round-2 baseline: strict CR3BP bench lines + ncu of the strict kernel + ncu of the fp64 peak microbenchmark
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
python bench.py --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3 --cpu-sample 4096 > $O/r2a_cr3bp_teval_strict.json 2> $O/r2a_cr3bp_teval_strict.err
python bench.py --workload cr3bp_dop853 --strict --trajectories 262144 --steps 3 --cpu-sample 4096 > $O/r2a_cr3bp_strict.json 2> $O/r2a_cr3bp_strict.err
python bench.py --workload cr3bp_dop853_teval --trajectories 262144 --steps 3 --no-cpu-baseline > $O/r2a_cr3bp_teval_fma.json 2> $O/r2a_cr3bp_teval_fma.err
python bench.py --workload vdp_dop853 --strict --steps 3 --no-cpu-baseline > $O/r2a_vdp_strict.json 2> $O/r2a_vdp_strict.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:erk_kernel -s 1 -c 1 -o $O/r2a_cr3bp_teval_strict -f python bench.py --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 1 --no-cpu-baseline > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dfma_peak -c 1 -o $O/r2a_dfma_peak -f python bench.py --trajectories 65536 --steps 1 --no-cpu-baseline > $O/ncu2.log 2>&1
ls -la $O | tail -20
