#!/bin/bash
# round-2 (session 3): N-GPU box: the two fixed tests + multi-device test, then the bench exactly as the driver launches it
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests -m gpu -q -k "analytic_jacobian_warp or multi_device" > $O/r2j_pytest.log 2>&1; tail -4 $O/r2j_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2j_bench_n$N.json 2> $O/r2j_bench_n$N.err
tail -5 $O/r2j_bench_n$N.err
python - <<PY
import json
for n in ($N,):
    try:
        d=json.load(open('$O/r2j_bench_n%d.json'%n))
        print(n, 'value %.4g'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'scaling', d['scaling'], 'by rank', d['e2e'].get('ms_per_step_by_rank'), 'clk', d['clocks'])
        print('   weak:', d.get('weak'))
        print('   single:', d.get('single_context'))
    except Exception as e: print(n, 'FAILED', e)
PY
