#!/usr/bin/env python
"""bench.py -- accepted steps/sec of the batched IVP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vdp_dop853]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: every rank integrates its shard of the synthetic
ensemble (default: 2**20 Van der Pol mu=1 trajectories PER GPU, DOP853, rtol=atol=1e-8, t in [0,100],
final state only -- the north-star workload of BASELINE.json / SURVEY 8d).  Trajectories are independent,
so ranks exchange nothing on the data path (weak scaling; torch.distributed is used for the barrier
and the max-over-ranks reduction of the device time only).

`value`   accepted steps/s with y0/params already resident in HBM (CUDA events around the K solves).
`e2e`     the same metric through the public host-buffer call (ivpb_solve_batch via
          ivp_b200.Context.solve_host): pinned host y0/params in, results back out, copies inside the
          timed region.
`roofline` algorithmic fp64 flops of one launch / its measured duration, against the DFMA-pipe peak
          measured on this GPU by libivpb's FMA-chain microbenchmark (MEASURED_PEAKS.json has no fp64
          entry); the path is FP64-FMA bound, not HBM or tensor bound.
`cpu_baseline` the CPU oracle (port of the reference's algorithm; the Rust crate cannot be built here)
          on all host threads over a bounded sample of the same ensemble.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ensemble, method, rtol, atol, F = flops of one RHS call, n)      [BASELINE.json configs]
    "vdp_dop853": ("vdp", "DOP853", 1e-8, 1e-8, 5, 2),                       # north star (the default bench line)
    "vdp_dopri5": ("vdp", "DOPRI5", 1e-6, 1e-9, 5, 2),
    "decay_dopri5": ("decay", "DOPRI5", 1e-6, 1e-9, 2, 1),                   # configs[1]
    "lorenz_dopri5": ("lorenz", "DOPRI5", 1e-6, 1e-9, 8, 3),
    "lorenz_rk4": ("lorenz", "RK4", 1e-6, 1e-9, 8, 3),
    "cr3bp_dop853_teval": ("cr3bp", "DOP853", 1e-10, 1e-12, 48, 6),          # configs[2]: 101 t_eval samples
    "cr3bp_dop853": ("cr3bp", "DOP853", 1e-10, 1e-12, 48, 6),                # same, final state only (A/B)
    "ball_dopri5_events": ("ball", "DOPRI5", 1e-8, 1e-10, 4, 2),             # configs[3]
    "ball_bounce_dopri5": ("ball_bounce", "DOPRI5", 1e-8, 1e-10, 4, 2),       # SolOut hook, bounces inside the solve
    "robertson_radau": ("robertson", "RADAU", 1e-6, 1e-6, 13, 3),            # configs[4]
    "robertson_bdf": ("robertson", "BDF", 1e-6, 1e-6, 13, 3),
    "robertson_dae_radau": ("robertson_dae", "RADAU", 1e-6, 1e-10, 14, 3),   # M y' = f, M = diag(1, 1, 0)
    "vdpstiff_radau": ("vdp_stiff", "RADAU", 1e-4, 1e-6, 5, 2),
    "vdpstiff_bdf": ("vdp_stiff", "BDF", 1e-4, 1e-6, 5, 2),
    "linear100_dopri5": ("linear100", "DOPRI5", 1e-6, 1e-8, 100, 100),       # warp-per-trajectory kernels (n > 32)
    "medakzo_radau": ("medakzo", "RADAU", 1e-5, 1e-7, 900, 64),              # warp-cooperative LU, n = 64
    "medakzo_bdf": ("medakzo", "BDF", 1e-5, 1e-7, 900, 64),
}
N_T_EVAL = {"cr3bp_dop853_teval": 101}
EXTRA_OPTIONS = {"robertson_dae_radau": {"mass_storage": "Full"}, "ball_bounce_dopri5": {"user_solout": True}}
NOMINAL_FP64_TFLOPS = 37.0   # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz


def algorithmic_flops(method: str, F: int, n: int, nstep, naccpt, dense: bool, counters=None, n_samples: int = 0) -> float:
    """SURVEY 8d table (mul/add/sub/div/sqrt = 1, FMA = 2); only work the kernel executes is counted."""
    nstep = np.asarray(nstep, dtype=np.float64)
    naccpt = np.asarray(naccpt, dtype=np.float64)
    if method in ("RADAU", "BDF"):
        # SURVEY appendix B (approximate: Newton iterations recovered from nfev; FD Jacobian = n + 1 RHS calls)
        c = np.asarray(counters, dtype=np.float64)
        nfev, njev, nlu = c[:, 0], c[:, 1], c[:, 2]
        if method == "RADAU":
            iters = np.maximum(nfev - 1 - naccpt, 0) / 3.0
            decomps = np.maximum(nlu - nstep, 0) / 2.0                        # nlu also counts one error solve per attempt
            fl = iters * (3 * F + 10 * n * n + 49 * n) + decomps * (10.0 / 3.0 * n ** 3) + nstep * (3 * n * n + 8 * n) \
                + njev * (n + 1) * F + naccpt * (F + 12 * n)
        else:
            iters = np.maximum(nfev - 1, 0)
            fl = iters * (F + 2 * n * n + 8 * n) + nlu * (2.0 / 3.0 * n ** 3 + n * n) + njev * (n + 1) * F + nstep * 20 * n
        return float(fl.sum())
    if method == "DOP853":
        fl = nstep * (11 * F + 157 * n + 45) + naccpt * F
        if dense:
            fl = fl + naccpt * (3 * F + 153 * n + 6)
    elif method == "DOPRI5":
        fl = nstep * (6 * F + 64 * n + 30 + (12 * n if dense else 0)) + (naccpt * 6 * n if dense else 0)
    elif method == "RK23":
        fl = nstep * (3 * F + 27 * n + 20)
    else:
        fl = nstep * (4 * F + 18 * n + 8)
    interp = {"DOP853": 14 * n + 3, "DOPRI5": 8 * n + 3, "RK23": 7 * n + 4}.get(method, 9 * n + 16)
    return float(fl.sum()) + float(n_samples) * interp


def workload_options(name: str, t0: float, tf: float, flags: int = 0, jac_mode: int = 0):
    """Options of a named workload (shared by the GPU arm, the cpu_baseline leg and --impl reference)."""
    from ivp_b200 import Method, Options
    ens, method, rtol, atol, F, n = WORKLOADS[name]
    n_te = N_T_EVAL.get(name, 0)
    extra = {"t_eval": np.linspace(t0, tf, n_te)} if n_te else {}
    if method == "RK4":
        extra["first_step"] = (tf - t0) / 1000.0
    if jac_mode:
        extra["jac_mode"] = 1
    extra.update(EXTRA_OPTIONS.get(name, {}))
    return Options(method=Method[method], rtol=rtol, atol=atol, flags=flags, max_events=1, **extra)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin: float, t_end: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t_begin <= t <= t_end + 0.2] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_reference(args, rank, world):
    """Reference arm: the reference's own algorithm on the host cores (oracle port; the Rust crate cannot be
    compiled in this image), all hardware threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    from ivp_b200 import Method, Options, synth
    from ivp_b200.api import PROBLEMS
    from oracle import pyoracle
    ens, method, rtol, atol, F, n = WORKLOADS[args.workload]
    cores = pyoracle.hardware_threads()
    sample = args.cpu_sample
    prob, y0, par, t0, tf = synth.ensemble(ens, sample)
    opts = workload_options(args.workload, t0, tf, jac_mode=args.jac_mode)
    for _ in range(args.warmup):
        pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0[:256], par[:256] if par is not None else None, opts, nthreads=cores)
    t_begin = time.perf_counter()
    acc = 0
    for _ in range(args.steps):
        o = pyoracle.solve_batch(PROBLEMS[prob], t0, tf, y0, par, opts, nthreads=cores, want=["status", "counters"])
        acc += int(o.naccpt.sum())
    dt = time.perf_counter() - t_begin
    val = acc / dt
    desc = f"{sample} trajectories/step of the {args.workload} ensemble (first rows of the seeded batch), std::thread x{cores}"
    print(json.dumps({
        "impl": "reference", "metric": "accepted_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "trajectories_per_step": sample, "method": method, "rtol": rtol, "atol": atol},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vdp_dop853", choices=sorted(WORKLOADS))
    ap.add_argument("--trajectories", type=int, default=1 << 20, help="trajectories per GPU per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --trajectories per GPU (default, the bench contract); strong: --trajectories in total, split over the ranks")
    ap.add_argument("--cpu-sample", type=int, default=32768, help="trajectories per step of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--static", action="store_true", help="disable work-queue refill (A/B)")
    ap.add_argument("--strict", action="store_true", help="-fmad=false kernel variant (A/B)")
    ap.add_argument("--fast-implicit", action="store_true", help="RADAU / BDF workloads: FMA-contracted kernels (default: strict)")
    ap.add_argument("--no-zerocopy", action="store_true", help="e2e arm: staged H2D/D2H copies instead of mapped pinned buffers (A/B)")
    ap.add_argument("--no-sort", action="store_true", help="RADAU / BDF: index order instead of the locality order of the ensemble (A/B)")
    ap.add_argument("--sort", action="store_true", help="explicit methods: locality order too (A/B)")
    ap.add_argument("--jac-mode", type=int, default=0, help="implicit workloads: 0 finite differences, 1 analytic")
    args = ap.parse_args()
    from ivp_b200.dist import dist_env, reduce_time_and_count, weak_offset
    rank, local_rank, world = dist_env()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from ivp_b200 import Method, Options, _abi, api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ivp_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    ens, method, rtol, atol, F, n = WORKLOADS[args.workload]
    Nper = args.trajectories if args.scaling == "weak" else (args.trajectories * (rank + 1)) // world - (args.trajectories * rank) // world
    offset = weak_offset(Nper, rank) if args.scaling == "weak" else (args.trajectories * rank) // world
    prob_name, y0_h, par_h, t0, tf = synth.ensemble(ens, Nper, offset=offset)
    problem = api.Problem.builtin(prob_name)
    flags = (api.IVPB_FLAG_NO_REFILL if args.static else 0) | (api.IVPB_FLAG_STRICT_FP if args.strict else 0) | \
        (api.IVPB_FLAG_NO_ZEROCOPY if args.no_zerocopy else 0) | (api.IVPB_FLAG_NO_SORT if args.no_sort else 0) | (api.IVPB_FLAG_SORT if args.sort else 0) | (api.IVPB_FLAG_FAST_FP if args.fast_implicit else 0)
    n_te = N_T_EVAL.get(args.workload, 0)
    opts = workload_options(args.workload, t0, tf, flags, args.jac_mode)
    mo = _abi.MarshalledOptions(opts, problem.n, problem.n_events)
    ctx = api.Context([local_rank])
    ne = problem.n_events

    # ---- device-resident arm -------------------------------------------------------------------
    y0_d = torch.from_numpy(y0_h).to(dev)
    par_d = torch.from_numpy(par_h).to(dev) if par_h is not None else None
    status_d = torch.empty(Nper, dtype=torch.int32, device=dev)
    counters_d = torch.empty((Nper, 6), dtype=torch.int32, device=dev)
    tfin_d = torch.empty(Nper, dtype=torch.float64, device=dev)
    yfin_d = torch.empty((Nper, problem.n), dtype=torch.float64, device=dev)
    d_out = {"status": status_d.data_ptr(), "counters": counters_d.data_ptr(), "t_final": tfin_d.data_ptr(),
             "y_final": yfin_d.data_ptr()}
    out_bytes_per_traj = 4 + 24 + 8 + 8 * problem.n
    if n_te:        # t_eval samples written straight to the preallocated [N][cap][n] block
        nout_d = torch.zeros(Nper, dtype=torch.int32, device=dev)
        yout_d = torch.zeros((Nper, mo.cap, problem.n), dtype=torch.float64, device=dev)
        d_out.update({"n_out": nout_d.data_ptr(), "y_out": yout_d.data_ptr()})
        out_bytes_per_traj += 4 + 8 * n_te * problem.n
    if ne:
        evc_d = torch.zeros((Nper, ne), dtype=torch.int32, device=dev)
        evt_d = torch.zeros((Nper, ne, 1), dtype=torch.float64, device=dev)
        d_out.update({"ev_count": evc_d.data_ptr(), "ev_t": evt_d.data_ptr()})
        out_bytes_per_traj += ne * 12
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def solve_device():
        ctx.solve_device(problem, t0, tf, Nper, y0_d.data_ptr(), par_d.data_ptr() if par_d is not None else None,
                         mo, d_out, stream=torch.cuda.current_stream().cuda_stream)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solve_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)                       # flush L2 between timed iterations (untimed)
        ev[s][0].record()
        solve_device()
        ev[s][1].record()
    barrier()
    wall1 = time.time()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(wall0, wall1)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    counters = counters_d.cpu().numpy().view(np.uint32)
    status = status_d.cpu().numpy()
    nstep, naccpt, nrejct = counters[:, 3], counters[:, 4], counters[:, 5]
    acc_local = int(naccpt.sum())
    n_samples = int(nout_d.sum().item()) if n_te else 0
    flops_launch = algorithmic_flops(method, F, n, nstep, naccpt, dense=bool(n_te or ne), counters=counters,
                                     n_samples=n_samples)

    total_ms_max, acc_all = reduce_time_and_count(total_ms, acc_local, dev, use_dist)
    value = acc_all * args.steps / (total_ms_max * 1e-3)

    # ---- end-to-end arm: public host-buffer API, pinned host memory, copies inside the timed region ----
    y0_p = torch.from_numpy(y0_h).pin_memory()
    par_p = torch.from_numpy(par_h).pin_memory() if par_h is not None else None
    h_status = torch.empty(Nper, dtype=torch.int32).pin_memory()
    h_counters = torch.empty((Nper, 6), dtype=torch.int32).pin_memory()
    h_tfin = torch.empty(Nper, dtype=torch.float64).pin_memory()
    h_yfin = torch.empty((Nper, problem.n), dtype=torch.float64).pin_memory()
    st = _abi.IvpbOutputs()
    import ctypes as C
    st.status = C.cast(h_status.data_ptr(), _abi.c_int32_p)
    st.counters = C.cast(h_counters.data_ptr(), _abi.c_uint32_p)
    st.t_final = C.cast(h_tfin.data_ptr(), _abi.c_double_p)
    st.y_final = C.cast(h_yfin.data_ptr(), _abi.c_double_p)
    y0_np, par_np = y0_p.numpy(), (par_p.numpy() if par_p is not None else None)
    h2d = y0_np.nbytes + (par_np.nbytes if par_np is not None else 0)
    d2h = h_status.numel() * 4 + h_counters.numel() * 4 + h_tfin.numel() * 8 + h_yfin.numel() * 8
    if n_te:
        h_nout = torch.empty(Nper, dtype=torch.int32).pin_memory()
        h_yout = torch.empty((Nper, mo.cap, problem.n), dtype=torch.float64).pin_memory()
        st.n_out = C.cast(h_nout.data_ptr(), _abi.c_int32_p)
        st.y_out = C.cast(h_yout.data_ptr(), _abi.c_double_p)
        d2h += h_nout.numel() * 4 + h_yout.numel() * 8
    if ne:
        h_evc = torch.empty((Nper, ne), dtype=torch.int32).pin_memory()
        h_evt = torch.empty((Nper, ne, 1), dtype=torch.float64).pin_memory()
        st.ev_count = C.cast(h_evc.data_ptr(), _abi.c_int32_p)
        st.ev_t = C.cast(h_evt.data_ptr(), _abi.c_double_p)
        d2h += h_evc.numel() * 4 + h_evt.numel() * 8
    for _ in range(2):
        ctx.solve_host(problem, t0, tf, y0_np, par_np, mo, st)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.solve_host(problem, t0, tf, y0_np, par_np, mo, st)     # returns after the D2H copies completed
    e2e_s = time.perf_counter() - w0
    e2e_acc = int(h_counters.numpy().view(np.uint32)[:, 4].sum())
    e2e_s_max, e2e_acc_all = reduce_time_and_count(e2e_s, float(e2e_acc), dev, use_dist)
    e2e_value = e2e_acc_all * e2e_steps / e2e_s_max
    assert e2e_acc == acc_local, "host-buffer and device-buffer paths disagree"

    # ---- roofline of the dominant (only) kernel, rank 0 ----
    peak_meas = ctx.measure_fp64_peak() if rank == 0 else 0.0
    kernel_ms = statistics.mean(step_ms)        # one launch per step (+ an 8-byte memset)
    achieved = flops_launch / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    alg_bytes = Nper * (8 * problem.n + 8 * problem.p + out_bytes_per_traj)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s",
                "frac": achieved / peak_meas if peak_meas else None, "traffic": traffic,
                "peak_source": "measured on this GPU: libivpb DFMA-chain microbenchmark (MEASURED_PEAKS.json has no fp64 entry)",
                "nominal_peak": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                "flops_per_launch": flops_launch, "kernel_ms": kernel_ms,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9}}

    # ---- CPU baseline (rank 0, N=1 only): oracle port on all host threads, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from ivp_b200.api import PROBLEMS
        from oracle import pyoracle
        cores = pyoracle.hardware_threads()
        sample = min(args.cpu_sample, Nper)
        best = 0.0
        for _ in range(2):
            c0 = time.perf_counter()
            o = pyoracle.solve_batch(PROBLEMS[prob_name], t0, tf, y0_h[:sample], par_h[:sample] if par_h is not None else None,
                                     opts, nthreads=cores, want=["status", "counters"])
            dtc = time.perf_counter() - c0
            best = max(best, float(o.naccpt.sum()) / dtc)
        parity = float(np.mean((o.naccpt == naccpt[:sample]) & (o.nrejct == nrejct[:sample])))
        cpu = {"value": best, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} trajectories of the same seeded ensemble, std::thread x{cores}, best of 2",
               "step_count_parity_on_sample": parity}

    if rank == 0:
        print(json.dumps({
            "metric": "accepted_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "problem": prob_name, "method": method, "rtol": rtol, "atol": atol,
                       "t_span": [t0, tf], "trajectories_per_gpu": Nper, "trajectories_total": Nper * world if args.scaling == "weak" else args.trajectories,
                       "outputs": "final state + status + counters" + (f" + {n_te} t_eval samples" if n_te else "") +
                                  (" + event times" if ne else ""), "parallelism": f"trajectory-sharded x{world}",
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "schedule": ("static" if args.static else "work-queue refill") + (", locality order" if (args.sort or (method in ("RADAU", "BDF") and not args.no_sort)) else ""), "fp": "strict" if (args.strict or (method in ("RADAU", "BDF") and not args.fast_implicit)) else "fma"},
            "accepted_steps_per_step": acc_all, "rejected_steps_rank0": int(nrejct.sum()),
            "status_success_frac_rank0": float(np.mean(status == 0)),
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s_max / e2e_steps * 1e3, "api": "ivpb_solve_batch (pinned host buffers" + (", staged copies)" if (args.no_zerocopy or args.sort or (method in ("RADAU", "BDF") and not args.no_sort)) else
                                                                  ", kernel reads/writes them over PCIe while integrating, t_eval samples included; event blocks staged)")},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
