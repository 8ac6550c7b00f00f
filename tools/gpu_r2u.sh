#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 120 "$@" > $O/r2u_$tag.json 2> $O/r2u_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2u_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'launches', d['gpu_launches'], 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2u_$tag.err | tr '\n' ' ')"; }
for wl in vdpstiff_bdf robertson_bdf; do
  t $wl python bench.py --workload $wl --steps 3 --cpu-sample 2048
done
t vdp python bench.py --workload vdp_dop853 --steps 10 --no-cpu-baseline
t vdp2 python bench.py --workload vdp_dop853 --steps 10 --no-cpu-baseline
