#!/bin/bash
# RADAU n <= 2 at 3 resident blocks per SM: stiff / implicit tests and the RADAU bench lines
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "stiff or implicit or golden or vdp_eps or mass" > $O/r2z16_pytest.log 2>&1; tail -2 $O/r2z16_pytest.log
t() { tag=$1; shift; timeout 100 "$@" > $O/r2z16_$tag.json 2> $O/r2z16_$tag.err; python -c "
import json;d=json.load(open('$O/r2z16_$tag.json'));c=d.get('cpu_baseline') or {}
print('$tag', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'parity', c.get('step_count_parity_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'))" 2>/dev/null || echo "$tag failed $(tail -c 200 $O/r2z16_$tag.err)"; }
t vdpstiff_radau python bench.py --workload vdpstiff_radau --steps 5 --cpu-sample 2048
