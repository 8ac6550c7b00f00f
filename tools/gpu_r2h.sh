#!/bin/bash
# round-2: N-GPU box: full GPU suite (multi-device test included) + the bench exactly as the driver launches it
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 2400 python -m pytest tests -m gpu -q > $O/r2h_pytest.log 2>&1; tail -8 $O/r2h_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/r2h_bench_n$N.json 2> $O/r2h_bench_n$N.err
tail -c 3000 $O/r2h_bench_n$N.json; tail -5 $O/r2h_bench_n$N.err
python bench.py --steps 10 --no-cpu-baseline > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.load(open('$O/r2h_bench_n%d.json'%n))
        print(n, 'value %.4g'%d['value'], 'ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'scaling', d['scaling'], 'by rank', d['e2e'].get('ms_per_step_by_rank'), 'clk', d['clocks'])
        print('   weak:', d.get('weak'))
        print('   single:', d.get('single_context'))
    except Exception as e: print(n, 'FAILED', e)
PY
