#!/bin/bash
# round-2 (session 3): resident blocks per SM of the register-resident implicit kernels (n <= 3) after the round-2 changes; CR3BP strict at 2 blocks
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
run() { # tag lib args...
  tag=$1; lib=$2; shift; shift
  IVPB_LIB=$lib python bench.py "$@" --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))" || tail -3 $O/$tag.err
}
for v in "" _mb3 _mb4 _mb6; do
  for wl in vdpstiff_radau vdpstiff_bdf robertson_radau robertson_bdf; do
    run r2n_${wl}$v ivp_b200/lib/libivpb$v.so --workload $wl --steps 3
  done
done
run r2n_cr3bp_teval_cmb2 ivp_b200/lib/libivpb_cmb2.so --workload cr3bp_dop853_teval --strict --trajectories 262144 --steps 3
run r2n_cr3bp_plain_cmb2 ivp_b200/lib/libivpb_cmb2.so --workload cr3bp_dop853 --strict --trajectories 262144 --steps 3
