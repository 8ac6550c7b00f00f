#!/bin/bash
# round-2 (session 3): strictd kernels (deferred division / square-root guards) + guarded re-run pass: whole GPU suite, then A/B against IVPB_NO_DEFER=1
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2q_pytest.log 2>&1; tail -6 $O/r2q_pytest.log
run() { # tag args...
  tag=$1; shift
  python bench.py "$@" > $O/$tag.json 2> $O/$tag.err
  python -c "
import json;d=json.load(open('$O/$tag.json'));c=d.get('cpu_baseline') or {}
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'launches', d['gpu_launches'])" || tail -3 $O/$tag.err
}
for wl in vdpstiff_radau vdpstiff_bdf robertson_radau robertson_bdf robertson_dae_radau; do
  run r2q_${wl} --workload $wl --steps 3 --cpu-sample 2048
  IVPB_NO_DEFER=1 run r2q_${wl}_nodefer --workload $wl --steps 3 --no-cpu-baseline
done
run r2q_cr3bp_teval --workload cr3bp_dop853_teval --trajectories 262144 --steps 3 --cpu-sample 2048
IVPB_NO_DEFER=1 run r2q_cr3bp_teval_nodefer --workload cr3bp_dop853_teval --trajectories 262144 --steps 3 --no-cpu-baseline
run r2q_cr3bp_plain --workload cr3bp_dop853 --trajectories 262144 --steps 3 --cpu-sample 2048
run r2q_vdp_strict --workload vdp_dop853 --strict --steps 5 --cpu-sample 4096
IVPB_NO_DEFER=1 run r2q_vdp_strict_nodefer --workload vdp_dop853 --strict --steps 5 --no-cpu-baseline
run r2q_vdp --workload vdp_dop853 --steps 10 --cpu-sample 4096
