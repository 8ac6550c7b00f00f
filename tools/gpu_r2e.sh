#!/bin/bash
# round-2 A/B: implicit reciprocal-cache / occupancy variants and chunk-signal overhead, same box
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
one() { # lib tag workload extra
  lib=$1; tag=$2; wl=$3; shift 3
  IVPB_LIB=$lib python bench.py --workload $wl --no-cpu-baseline "$@" > $O/r2e_${tag}_$wl.json 2> $O/r2e_${tag}_$wl.err
  python -c "
import json
try:
    d=json.load(open('$O/r2e_${tag}_$wl.json')); print('$tag $wl', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
except Exception as e: print('$tag $wl FAILED', open('$O/r2e_${tag}_$wl.err').read()[-600:])"
}
L=ivp_b200/lib
for rep in 1 2; do
  one $L/libivpb.so main$rep vdp_dop853 --steps 10
  one $L/libivpb_nosig.so nosig$rep vdp_dop853 --steps 10
done
for v in main p0s0 mb4 p0s1 p1s0 mb3 p0s0mb4; do
  [ $v != main ] && [ ! -f $L/libivpb_$v.so ] && continue
  lib=$L/libivpb_$v.so; [ $v = main ] && lib=$L/libivpb.so
  for wl in robertson_radau robertson_bdf vdpstiff_radau vdpstiff_bdf robertson_dae_radau; do
    one $lib $v $wl --steps 3
  done
done
