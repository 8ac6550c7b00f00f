#!/bin/bash
# round-2 (session 3): strictd implicit kernels (flag in a register, NZ forms); second pass exercised with IVPB_DEBUG_RERUN; every command under its own timeout
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 120 "$@" > $O/r2t_$tag.json 2> $O/r2t_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2t_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'launches', d['gpu_launches'], 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2t_$tag.err | tr '\n' ' ')"; }
t vdp python bench.py --workload vdp_dop853 --steps 10 --cpu-sample 4096
for wl in vdpstiff_radau vdpstiff_bdf robertson_radau robertson_bdf robertson_dae_radau; do
  t $wl python bench.py --workload $wl --steps 3 --cpu-sample 2048
done
IVPB_DEBUG_RERUN=3 t robertson_bdf_rerun3 python bench.py --workload robertson_bdf --steps 2 --cpu-sample 2048
IVPB_DEBUG_RERUN=1 t vdpstiff_radau_rerun1 python bench.py --workload vdpstiff_radau --steps 2 --cpu-sample 2048
IVPB_DEBUG_RERUN=2 timeout 600 python -m pytest tests -m gpu -q -x -k "stiff or implicit or radau or bdf or robertson or mass or hook or dense or golden" > $O/r2t_pytest_rerun.log 2>&1; tail -4 $O/r2t_pytest_rerun.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2t_pytest.log 2>&1; tail -4 $O/r2t_pytest.log
