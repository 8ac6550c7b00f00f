#!/bin/bash
# round-2: full GPU suite + bench lines after exact div / block-sync / pilot / pipelined D2H
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1
tail -5 $O/r2c_pytest.log
run() { # tag, args...
  tag=$1; shift
  python bench.py "$@" > $O/$tag.json 2> $O/$tag.err
  python - <<PY
import json
try:
    d=json.load(open('$O/$tag.json'))
    print('$tag', 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'fp', d['config']['fp'], 'cpu', {k:v for k,v in (d.get('cpu_baseline') or {}).items() if k not in ('sample','unit','kind')}, 'clk', d['clocks'])
except Exception as e:
    print('$tag FAILED', e); print(open('$O/$tag.err').read()[-1500:])
PY
}
run r2c_vdp_dop853 --steps 10
run r2c_cr3bp_dop853_teval --workload cr3bp_dop853_teval --steps 3 --cpu-sample 8192
run r2c_cr3bp_dop853 --workload cr3bp_dop853 --steps 3 --cpu-sample 8192
run r2c_ball --workload ball_dopri5_events --steps 10
run r2c_ball_1M --workload ball_dopri5_events --steps 10 --trajectories 1048576 --no-cpu-baseline
run r2c_decay --workload decay_dopri5 --steps 10
run r2c_lorenz --workload lorenz_dopri5 --steps 10
run r2c_vdp_dopri5 --workload vdp_dopri5 --steps 10
run r2c_vdp_zcout --steps 10 --no-cpu-baseline
run r2c_cr3bp_teval_fast --workload cr3bp_dop853_teval --steps 3 --fast --no-cpu-baseline
