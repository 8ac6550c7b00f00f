#!/bin/bash
# round-2 (session 3): implicit kernels: main (4/3 blocks, RADAU constants + pow head from the constant bank) vs pow head as immediates vs divisions without guards (upper bound of what grouping the guards can give)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
run() { # tag lib args...
  tag=$1; lib=$2; shift; shift
  IVPB_LIB=$lib python bench.py "$@" > $O/$tag.json 2> $O/$tag.err
  python -c "
import json;d=json.load(open('$O/$tag.json'));c=d.get('cpu_baseline') or {}
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'))" || tail -3 $O/$tag.err
}
for v in "" _plainhead _ung; do
  for wl in vdpstiff_radau vdpstiff_bdf robertson_radau robertson_bdf; do
    run r2p_${wl}$v ivp_b200/lib/libivpb$v.so --workload $wl --steps 3 --cpu-sample 2048
  done
done
