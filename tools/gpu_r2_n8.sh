#!/bin/bash
# round-2: 8-GPU box: multi-device test, the bench exactly as the driver launches it (strong scaling + weak + single context), N = 1 beside it
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 200 python -m pytest tests -m gpu -q -k "multi_device" > $O/r2z9n8_pytest.log 2>&1; tail -2 $O/r2z9n8_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2z9_bench_n$N.json 2> $O/r2z9_bench_n$N.err
tail -3 $O/r2z9_bench_n$N.err
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2z9n8_bench_n1.json 2> $O/r2z9n8_bench_n1.err
python - <<PY
import json
for f in ('$O/r2z9n8_bench_n1.json', '$O/r2z9_bench_n$N.json'):
    try:
        s=open(f).read(); d=json.loads(s[s.index('{"metric'):])
        print(d['n_gpus'], 'value %.4g'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e val %.4g'%d['e2e']['value'], d['scaling'], 'by rank', d['e2e'].get('ms_per_step_by_rank'), 'clk', d['clocks'])
        print('   weak:', d.get('weak'))
        print('   single:', d.get('single_context'))
    except Exception as e: print(f, 'FAILED', e)
PY
