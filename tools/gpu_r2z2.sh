#!/bin/bash
# round 2, session 4: full GPU suite on the rebuilt tree, smoke, default bench + reference arm, the strong-scaling shard with
# the small-shard grid rule, and two ncu source-level captures (BDF VdP mu=1000, strict CR3BP) for the next optimisation step
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2z2_pytest.log 2>&1; tail -3 $O/r2z2_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z2_smoke.log 2>&1; tail -1 $O/r2z2_smoke.log
timeout 300 python bench.py > $O/r2z2_bench_n1.json 2> $O/r2z2_bench_n1.err; python -c "
import json;d=json.load(open('$O/r2z2_bench_n1.json'));print('n1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference > $O/r2z2_bench_reference_n1.json 2> $O/r2z2_bench_reference_n1.err; tail -c 300 $O/r2z2_bench_reference_n1.json; echo
timeout 200 python bench.py --trajectories 131072 --steps 30 --warmup 5 --no-cpu-baseline > $O/r2z2_bench_strong_131072.json 2> $O/r2z2_strong.err; python -c "
import json;d=json.load(open('$O/r2z2_bench_strong_131072.json'));print('131072', d['ms_per_step'], d['e2e']['ms_per_step'])"
cap() { tag=$1; k=$2; skip=$3; shift; shift; shift
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$tag -f python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline > $O/$tag.log 2>&1
  python tools/ncu_summary.py $O/$tag.ncu-rep $O/${tag}_ncu_full.txt > /dev/null 2>&1
  grep -E "Kernel Name|duration|registers_per|issue_active|thread_inst_executed_per|pipe_fp64_cycles|no_instruction|stalled_wait" $O/${tag}_ncu_full.txt | cut -c1-150; }
cap r2z2_vdpstiff_bdf implicit_kernel 2 --workload vdpstiff_bdf --trajectories 131072
cap r2z2_cr3bp_strict erk_kernel 1 --workload cr3bp_dop853_teval --strict --trajectories 131072
ls -la $O/*.ncu-rep
