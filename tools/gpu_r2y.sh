#!/bin/bash
# round 2, session 4: resident blocks per SM for small (strong-scaling) shards, locality order on small shards, the new
# full-size random-subset parity test
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 150 "$@" > $O/r2y_$tag.json 2> $O/r2y_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2y_$tag.json'))
print('$tag', d['config']['trajectories_per_gpu'], 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3))" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2y_$tag.err | tr '\n' ' ')"; }
timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "full_size" > $O/r2y_pytest.log 2>&1; tail -3 $O/r2y_pytest.log
for occ in 7 6 5 4 3; do IVPB_GRID_OCC=$occ t ns131k_occ$occ python bench.py --trajectories 131072 --steps 30 --warmup 5 --no-cpu-baseline; done
IVPB_GRID_OCC=7 t ns131k_occ7_sort python bench.py --trajectories 131072 --steps 30 --warmup 5 --no-cpu-baseline --sort
IVPB_GRID_OCC=5 t ns131k_occ5_sort python bench.py --trajectories 131072 --steps 30 --warmup 5 --no-cpu-baseline --sort
for occ in 7 5 4; do IVPB_GRID_OCC=$occ t ns262k_occ$occ python bench.py --trajectories 262144 --steps 20 --warmup 5 --no-cpu-baseline; done
for occ in 7 5; do IVPB_GRID_OCC=$occ t ns524k_occ$occ python bench.py --trajectories 524288 --steps 20 --warmup 5 --no-cpu-baseline; done
t ns1M python bench.py --steps 10 --warmup 3 --no-cpu-baseline
t cr3bp131k python bench.py --workload cr3bp_dop853_teval --trajectories 131072 --steps 3 --warmup 1 --no-cpu-baseline
t cr3bp131k_sort python bench.py --workload cr3bp_dop853_teval --trajectories 131072 --steps 3 --warmup 1 --no-cpu-baseline --sort
t cr3bp262k_sort python bench.py --workload cr3bp_dop853_teval --trajectories 262144 --steps 3 --warmup 1 --no-cpu-baseline --sort
