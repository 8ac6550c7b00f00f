#!/bin/bash
# round-2 (session 3): strictd kernels with the flag in a register; every command under its own timeout
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 150 "$@" > $O/r2s_$tag.json 2> $O/r2s_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2s_$tag.json'));c=d.get('cpu_baseline') or {}
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'launches', d['gpu_launches'])" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2s_$tag.err | tr '\n' ' ')"; }
t plain8k python bench.py --workload cr3bp_dop853 --strict --trajectories 8192 --steps 1 --warmup 1 --no-cpu-baseline
t teval8k python bench.py --workload cr3bp_dop853_teval --strict --trajectories 8192 --steps 1 --warmup 1 --no-cpu-baseline
for wl in vdpstiff_radau vdpstiff_bdf robertson_radau robertson_bdf robertson_dae_radau; do
  t $wl python bench.py --workload $wl --steps 3 --cpu-sample 2048
done
t cr3bp_teval python bench.py --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3 --cpu-sample 2048
t cr3bp_teval_pilot python bench.py --workload cr3bp_dop853_teval --trajectories 262144 --steps 3 --no-cpu-baseline
IVPB_NO_DEFER=1 t cr3bp_teval_nodefer python bench.py --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3 --no-cpu-baseline
t cr3bp_plain python bench.py --workload cr3bp_dop853 --strict --trajectories 262144 --steps 3 --cpu-sample 2048
t vdp_strict python bench.py --workload vdp_dop853 --strict --steps 5 --cpu-sample 4096
IVPB_NO_DEFER=1 t vdp_strict_nodefer python bench.py --workload vdp_dop853 --strict --steps 5 --no-cpu-baseline
t lorenz_strict python bench.py --workload lorenz_dopri5 --strict --steps 5 --cpu-sample 4096
t vdp python bench.py --workload vdp_dop853 --steps 10 --cpu-sample 4096
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2s_pytest.log 2>&1; tail -6 $O/r2s_pytest.log
