#!/bin/bash
# round 2, session 4: 2 GPUs -- the multi-device context test and the driver-style torchrun bench (strong scaling default)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "multi_device" > $O/r2z7_pytest.log 2>&1; tail -3 $O/r2z7_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2z7_bench_n2.json 2> $O/r2z7_bench_n2.err
python -c "
import json;d=json.load(open('$O/r2z7_bench_n2.json'));print('n2', d['scaling'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('strong'), d.get('single_context'))" || tail -c 600 $O/r2z7_bench_n2.err
