#!/bin/bash
# round 2, session 4: resident blocks of the strictd RADAU / BDF kernels (n = 2) re-measured after the instruction cuts
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
t() { tag=$1; shift; timeout 200 "$@" > $O/r2z10_$tag.json 2> $O/r2z10_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2z10_$tag.json'))
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2z10_$tag.err | tr '\n' ' ')"; }
for v in mb3 mb5; do for wl in vdpstiff_bdf vdpstiff_radau; do IVPB_LIB=ivp_b200/lib/libivpb_$v.so t ${wl}_$v python bench.py --workload $wl --steps 5 --no-cpu-baseline; done; done
