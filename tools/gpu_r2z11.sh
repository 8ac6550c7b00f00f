#!/bin/bash
# round 2, END OF ROUND (session 4): full GPU suite, smoke, default bench + reference arm, side workloads, launch list + north-star ncu
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2z11_pytest.log 2>&1; tail -3 $O/r2z11_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z11_smoke.log 2>&1; tail -1 $O/r2z11_smoke.log
b() { tag=$1; shift; timeout 300 python bench.py "$@" > $O/r2z11_bench_$tag.json 2> $O/r2z11_bench_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2z11_bench_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', d['config']['trajectories_per_gpu'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'fp', fp.get('mode') if isinstance(fp,dict) else fp, 'parity', c.get('step_count_parity_on_sample'), 'tol', c.get('in_tolerance_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2z11_bench_$tag.err | tr '\n' ' ')"; }
b n1
timeout 300 python bench.py --impl reference > $O/r2z11_bench_reference_n1.json 2> $O/r2z11_bench_reference_n1.err; tail -c 300 $O/r2z11_bench_reference_n1.json; echo
b cr3bp_dop853_teval --workload cr3bp_dop853_teval --steps 3 --cpu-sample 2048
b cr3bp_dop853 --workload cr3bp_dop853 --steps 3 --cpu-sample 2048
for wl in robertson_radau robertson_bdf robertson_dae_radau vdpstiff_radau vdpstiff_bdf; do b $wl --workload $wl --steps 5 --cpu-sample 2048; done
b medakzo_bdf --workload medakzo_bdf --steps 3 --cpu-sample 64
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z11_vdp_dop853_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2z11_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:erk_kernel -s 1 -c 1 -o $O/r2z11_vdp_dop853 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r2z11_vdp_dop853.log 2>&1
python tools/ncu_summary.py $O/r2z11_vdp_dop853.ncu-rep $O/r2z11_vdp_dop853_ncu_full.txt vdp_dop853 > /dev/null 2>&1
grep -E "Kernel Name|duration|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|dram__bytes" $O/r2z11_vdp_dop853_ncu_full.txt | cut -c1-150
