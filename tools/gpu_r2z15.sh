#!/bin/bash
# e2e route of a small shard (131072 trajectories): chunked pipeline (default) vs copies after the kernel vs mapped outputs
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 100 "$@" > $O/r2z15_$tag.json 2> $O/r2z15_$tag.err; python -c "
import json;d=json.load(open('$O/r2z15_$tag.json'));print('$tag', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" 2>/dev/null || echo "$tag failed $(tail -c 200 $O/r2z15_$tag.err)"; }
t default python bench.py --trajectories 131072 --steps 40 --warmup 5 --no-cpu-baseline
t nopipeline python bench.py --trajectories 131072 --steps 40 --warmup 5 --no-cpu-baseline --no-pipeline
t zcout python bench.py --trajectories 131072 --steps 40 --warmup 5 --no-cpu-baseline --zerocopy-out
