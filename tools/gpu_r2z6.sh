#!/bin/bash
# round 2, session 4: DOPRI5 / RK23 fast kernels after moving the log2 / exp2 polynomial coefficients into the constant bank
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "vdp or decay or lorenz or ball or sho or nvrtc or fastmath or step_mode or t_eval or pilot or edge or warp_per" > $O/r2z6_pytest.log 2>&1; tail -3 $O/r2z6_pytest.log
t() { tag=$1; shift; timeout 200 "$@" > $O/r2z6_$tag.json 2> $O/r2z6_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2z6_$tag.json'));c=d.get('cpu_baseline') or {}
print('$tag', d['config']['trajectories_per_gpu'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'tol', c.get('in_tolerance_on_sample'))" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2z6_$tag.err | tr '\n' ' ')"; }
t vdp_dopri5 python bench.py --workload vdp_dopri5 --steps 10 --cpu-sample 8192
t lorenz_dopri5 python bench.py --workload lorenz_dopri5 --steps 10 --cpu-sample 8192
t decay_dopri5 python bench.py --workload decay_dopri5 --steps 10 --cpu-sample 8192
t ball_dopri5_events python bench.py --workload ball_dopri5_events --steps 10 --cpu-sample 8192
t ball_bounce_dopri5 python bench.py --workload ball_bounce_dopri5 --steps 10 --cpu-sample 4096
