#!/bin/bash
# round-2 (session 2): full GPU suite + bench lines after the PIPE twin kernels, input arrival flags, warp-kernel mass / jac
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2g_pytest.log 2>&1
tail -15 $O/r2g_pytest.log
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
run r2g_vdp_dop853 --steps 10
run r2g_vdp_dop853_nopipe --steps 10 --no-cpu-baseline --no-pipeline
run r2g_ball --workload ball_dopri5_events --steps 10 --no-cpu-baseline
run r2g_decay --workload decay_dopri5 --steps 10 --no-cpu-baseline
run r2g_decay_nopipe --workload decay_dopri5 --steps 10 --no-cpu-baseline --no-pipeline
run r2g_lorenz --workload lorenz_dopri5 --steps 10 --no-cpu-baseline
run r2g_vdpstiff_bdf --workload vdpstiff_bdf --steps 3 --no-cpu-baseline
run r2g_vdpstiff_radau --workload vdpstiff_radau --steps 3 --no-cpu-baseline
