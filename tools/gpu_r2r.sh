#!/bin/bash
# debug: which explicit strictd configuration does not return (every command under its own timeout)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 90 "$@" > $O/r2r_$tag.log 2>&1; echo "$tag rc=$? $(tail -c 300 $O/r2r_$tag.log | tr '\n' ' ')"; }
t plain8k python bench.py --workload cr3bp_dop853 --strict --trajectories 8192 --steps 1 --warmup 1 --no-cpu-baseline
t teval8k python bench.py --workload cr3bp_dop853_teval --strict --trajectories 8192 --steps 1 --warmup 1 --no-cpu-baseline
t teval64k python bench.py --workload cr3bp_dop853_teval --strict --trajectories 65536 --steps 1 --warmup 1 --no-cpu-baseline
t teval256k python bench.py --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 1 --warmup 1 --no-cpu-baseline
t teval256k_pilot python bench.py --workload cr3bp_dop853_teval --trajectories 262144 --steps 1 --warmup 1 --no-cpu-baseline
t vdp_strict python bench.py --workload vdp_dop853 --strict --steps 2 --warmup 1 --no-cpu-baseline
t lorenz_strict python bench.py --workload lorenz_dopri5 --strict --steps 2 --warmup 1 --no-cpu-baseline
