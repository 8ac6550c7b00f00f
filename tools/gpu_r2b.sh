#!/bin/bash
# round-2 experiment: shared-reciprocal exact div/sqrt + block-synchronous trips, A/B on CR3BP
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strict or bit_exact or golden or cr3bp or t_eval" > $O/r2b_pytest.log 2>&1
tail -3 $O/r2b_pytest.log
for fp in "--strict" ""; do
for cfg in "0 128" "1 128" "1 256" "0 256"; do
  set -- $cfg
  tag="r2b_cr3bp_teval${fp:+_strict}_s$1_t$2"
  IVPB_BLOCK_SYNC=$1 IVPB_BLOCK_THREADS=$2 python bench.py --workload cr3bp_dop853_teval $fp --trajectories 262144 --steps 3 --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
done
for cfg in "0 128" "1 128"; do
  set -- $cfg
  for wl in vdp_dop853 lorenz_dopri5; do
  tag="r2b_${wl}_s$1"
  IVPB_BLOCK_SYNC=$1 python bench.py --workload $wl --steps 5 --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
  done
  for wl in robertson_bdf vdpstiff_radau; do
  tag="r2b_${wl}_s$1"
  IVPB_BLOCK_SYNC=$1 python bench.py --workload $wl --trajectories 262144 --steps 3 --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err
  python -c "import json;d=json.load(open('$O/$tag.json'));print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
  done
done
