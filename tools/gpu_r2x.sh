#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
IVPB_LIB=ivp_b200/lib/libivpb_gdbg.so timeout 100 python - > $O/r2x_guard_debug.log 2>&1 <<PY
import numpy as np, ivp_b200 as ib
from ivp_b200 import Method, Options, synth
for wl, m in (("robertson", Method.BDF), ("robertson", Method.RADAU), ("vdp_stiff", Method.BDF)):
    prob, y0, par, t0, tf = synth.ensemble(wl, 256)
    g = ib.solve_ivp_batch(prob, t0, tf, y0, par, Options(method=m, rtol=1e-6, atol=1e-6))
    print(wl, m, "status", np.unique(g.status), "reruns", ib.api.default_context().last_reruns(), flush=True)
PY
grep -v "div<0>.*a=0 \|div<0> a=-0 " $O/r2x_guard_debug.log | head -20
t() { tag=$1; shift; timeout 120 "$@" > $O/r2x_$tag.json 2> $O/r2x_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2x_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2x_$tag.err | tr '\n' ' ')"; }
for wl in robertson_bdf vdpstiff_bdf robertson_radau vdpstiff_radau; do t $wl python bench.py --workload $wl --steps 3 --cpu-sample 2048; done
