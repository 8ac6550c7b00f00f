#!/bin/bash
# round-2 END OF ROUND evidence (1 GPU): GPU suite, bench lines of every workload (with the CPU oracle sample), parity report,
# ncu launch list + --set full captures.  Every command runs under its own timeout.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2z_pytest.log 2>&1; tail -3 $O/r2z_pytest.log
b() { tag=$1; shift; timeout 300 python bench.py "$@" > $O/r2z_bench_$tag.json 2> $O/r2z_bench_$tag.err; rc=$?
  python -c "
import json;d=json.load(open('$O/r2z_bench_$tag.json'));c=d.get('cpu_baseline') or {}
fp=d['config']['fp']
print('$tag', d['config']['trajectories_per_gpu'], 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'fp', fp.get('mode') if isinstance(fp,dict) else fp, 'parity', c.get('step_count_parity_on_sample'), 'tol', c.get('in_tolerance_on_sample'), 'bits', c.get('bit_identical_y_final_on_sample'), 'reruns', fp.get('second_pass_trajectories') if isinstance(fp,dict) else None)" 2>/dev/null || echo "$tag rc=$rc $(tail -c 300 $O/r2z_bench_$tag.err | tr '\n' ' ')"; }
b n1
timeout 300 python bench.py --impl reference > $O/r2z_bench_reference_n1.json 2> $O/r2z_bench_reference_n1.err; tail -c 400 $O/r2z_bench_reference_n1.json; echo
b vdp_dopri5 --workload vdp_dopri5 --steps 10 --cpu-sample 8192
b decay_dopri5 --workload decay_dopri5 --steps 10 --cpu-sample 8192
b lorenz_dopri5 --workload lorenz_dopri5 --steps 10 --cpu-sample 8192
b cr3bp_dop853_teval --workload cr3bp_dop853_teval --steps 3 --cpu-sample 2048
b cr3bp_dop853 --workload cr3bp_dop853 --steps 3 --cpu-sample 2048
b ball_dopri5_events --workload ball_dopri5_events --steps 10 --cpu-sample 8192
b ball_dopri5_events_1M --workload ball_dopri5_events --trajectories 1048576 --steps 10 --no-cpu-baseline
b ball_bounce_dopri5 --workload ball_bounce_dopri5 --steps 10 --cpu-sample 4096
for wl in robertson_radau robertson_bdf robertson_dae_radau vdpstiff_radau vdpstiff_bdf; do b $wl --workload $wl --steps 5 --cpu-sample 2048; done
b linear100_dopri5 --workload linear100_dopri5 --steps 5 --cpu-sample 512
b medakzo_radau --workload medakzo_radau --steps 3 --cpu-sample 64
b medakzo_bdf --workload medakzo_bdf --steps 3 --cpu-sample 64
b strong_131072 --trajectories 131072 --steps 20 --no-cpu-baseline
timeout 600 python tools/parity_report.py > $O/r2z_parity.md 2> $O/r2z_parity.err; tail -3 $O/r2z_parity.md
# ncu: launch list of the default bench command, then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_vdp_dop853_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2z_ncu_launches.log 2>&1
cap() { tag=$1; k=$2; skip=$3; shift; shift; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$tag -f python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline > $O/$tag.log 2>&1
  python tools/ncu_summary.py $O/$tag.ncu-rep $O/${tag}_ncu_full.txt "$WL" > /dev/null 2>&1
  grep -E "Kernel Name|duration|registers_per|warps_active|issue_active|thread_inst_executed_per|pipe_fp64_cycles|local_ld|local_st|no_instruction|stalled_wait|dram__bytes" $O/${tag}_ncu_full.txt | cut -c1-160; }
WL=vdp_dop853 cap r2z_vdp_dop853 erk_kernel 1
WL= cap r2z_vdpstiff_radau_strictd implicit_kernel 2 --workload vdpstiff_radau
WL= cap r2z_vdpstiff_bdf_strictd implicit_kernel 2 --workload vdpstiff_bdf
WL= cap r2z_cr3bp_dop853_teval_strict erk_kernel 1 --workload cr3bp_dop853_teval --strict --trajectories 262144
rm -f $O/r2z_*.ncu-rep.tmp; ls -la $O | grep r2z | wc -l
