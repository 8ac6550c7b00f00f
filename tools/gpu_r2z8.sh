#!/bin/bash
# round 2, session 4: Robertson BDF with zero-accepting solve / norm divisions (Prob::NEWTON_ZEROS) and the dy_norm == 0 exit
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stiff or implicit or golden or hook or mass or vdp_eps or dense_output or locality or nvrtc" > $O/r2z8_pytest.log 2>&1; tail -3 $O/r2z8_pytest.log
t() { tag=$1; shift; timeout 200 "$@" > $O/r2z8_$tag.json 2> $O/r2z8_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2z8_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2z8_$tag.err | tr '\n' ' ')"; }
for wl in robertson_bdf vdpstiff_bdf robertson_radau; do t $wl python bench.py --workload $wl --steps 5 --cpu-sample 4096; done
IVPB_NO_DEFER=1 t robertson_bdf_guarded python bench.py --workload robertson_bdf --steps 5 --no-cpu-baseline
