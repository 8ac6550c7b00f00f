#!/bin/bash
# does the NVML clock poller (every 2 ms, python thread) perturb the device-timed leg?
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
for ms in 2 20 200 2; do IVPB_BENCH_POLL_MS=$ms timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2z13_poll$ms.json 2> $O/r2z13_poll$ms.err; python -c "
import json;d=json.load(open('$O/r2z13_poll$ms.json'));print('poll', $ms, 'ms:', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['clocks']['samples'])"; done
