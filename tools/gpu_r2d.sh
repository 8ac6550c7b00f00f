#!/bin/bash
# round-2: full GPU suite + bench lines (completion-flag D2H, implicit reciprocal caching)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r2d_pytest.log 2>&1
tail -15 $O/r2d_pytest.log
run() { # tag, args...
  tag=$1; shift
  python bench.py "$@" > $O/$tag.json 2> $O/$tag.err
  python - <<PY
import json
try:
    d=json.load(open('$O/$tag.json'))
    fp=d['config']['fp']; fp=(fp['mode']+'/'+fp['source']) if isinstance(fp,dict) else fp
    c=d.get('cpu_baseline') or {}
    print('$tag', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), fp, 'par', c.get('step_count_parity_on_sample'), 'tol', c.get('in_tolerance_on_sample'), 'smp', d['clocks']['samples'])
except Exception as e:
    print('$tag FAILED', e); print(open('$O/$tag.err').read()[-1500:])
PY
}
run r2d_vdp_dop853 --steps 10
run r2d_vdp_dop853_nopipe --steps 10 --no-cpu-baseline --no-pipeline
run r2d_vdp_dop853_zcout --steps 10 --no-cpu-baseline --zerocopy-out
run r2d_cr3bp_dop853_teval --workload cr3bp_dop853_teval --steps 3 --cpu-sample 8192
run r2d_ball --workload ball_dopri5_events --steps 10
run r2d_ball_1M --workload ball_dopri5_events --steps 10 --trajectories 1048576 --no-cpu-baseline
run r2d_decay --workload decay_dopri5 --steps 10
run r2d_lorenz --workload lorenz_dopri5 --steps 10
run r2d_robertson_radau --workload robertson_radau --steps 5
run r2d_robertson_bdf --workload robertson_bdf --steps 5
run r2d_vdpstiff_radau --workload vdpstiff_radau --steps 5
run r2d_vdpstiff_bdf --workload vdpstiff_bdf --steps 5
run r2d_robertson_dae_radau --workload robertson_dae_radau --steps 5
run r2d_strong_131072 --steps 10 --trajectories 131072 --no-cpu-baseline
