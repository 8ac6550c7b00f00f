#!/usr/bin/env python
"""bench.py -- accepted steps/sec of the batched IVP hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vdp_dop853]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: the synthetic ensemble of the workload is integrated once.
The default workload is the north star of BASELINE.json / SURVEY 8d: 2**20 Van der Pol mu=1 trajectories, DOP853,
rtol=atol=1e-8, t in [0,100], final state only.  Every workload carries its BASELINE.json size (WORKLOADS).

Scaling.  BASELINE.json's metric is "the 1M-trajectory ensemble at 1/2/4/8 B200", so with N > 1 ranks the headline is
STRONG scaling: the same ensemble split into N contiguous shards, one per rank (`--scaling weak` keeps the workload
size PER GPU instead; its numbers are added to the same JSON line under "weak").  Trajectories are independent, so
ranks exchange nothing on the data path; torch.distributed only provides the barrier and the MAX / SUM reductions of
the device time and the step counts.

`value`     accepted steps/s with y0/params already resident in HBM (CUDA events around the K solves, MAX over ranks).
`e2e`       the same metric through the public host-buffer call (ivpb_solve_batch via ivp_b200.Context.solve_host):
            pinned host y0/params in, results back out, copies inside the timed region.
`roofline`  algorithmic fp64 flops of one launch / its measured duration.  `peak` / `frac` use the NOMINAL 37 TFLOP/s
            (148 SM x 64 DFMA/clk x 2 x 1.965 GHz) because MEASURED_PEAKS.json carries no fp64 entry; the DFMA-chain
            microbenchmark of this GPU (ivpb_measure_fp64_peak, ncu capture under profiles/) is reported next to it
            as `peak_measured` / `frac_measured`.  The path is FP64-pipe bound, not HBM or tensor bound.
`cpu_baseline` the CPU oracle (port of the reference's algorithm; the Rust crate cannot be built here) on all host
            threads over a seeded random sample of the same ensemble; the same sample is the in-run parity check.
`single_context` (N > 1, rank 0) ONE ivpb_ctx over all N devices: ivpb_solve_batch from host buffers and
            ivpb_solve_batch_device with device-0-resident buffers (peer scatter / gather over NVLink).
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

M1, K256 = 1 << 20, 1 << 18
WORKLOADS = {
    # name: (ensemble, method, rtol, atol, F = flops of one RHS call, n, BASELINE.json size)      [BASELINE.json configs]
    "vdp_dop853": ("vdp", "DOP853", 1e-8, 1e-8, 5, 2, M1),                       # north star (the default bench line)
    "vdp_dopri5": ("vdp", "DOPRI5", 1e-6, 1e-9, 5, 2, M1),
    "decay_dopri5": ("decay", "DOPRI5", 1e-6, 1e-9, 2, 1, M1),                   # configs[1]
    "lorenz_dopri5": ("lorenz", "DOPRI5", 1e-6, 1e-9, 8, 3, M1),
    "lorenz_rk4": ("lorenz", "RK4", 1e-6, 1e-9, 8, 3, M1),
    "cr3bp_dop853_teval": ("cr3bp", "DOP853", 1e-10, 1e-12, 48, 6, M1),          # configs[2]: 101 t_eval samples
    "cr3bp_dop853": ("cr3bp", "DOP853", 1e-10, 1e-12, 48, 6, M1),                # same, final state only (A/B)
    "ball_dopri5_events": ("ball", "DOPRI5", 1e-8, 1e-10, 4, 2, K256),           # configs[3]
    "ball_bounce_dopri5": ("ball_bounce", "DOPRI5", 1e-8, 1e-10, 4, 2, K256),    # SolOut hook, bounces inside the solve
    "robertson_radau": ("robertson", "RADAU", 1e-6, 1e-6, 13, 3, K256),          # configs[4]
    "robertson_bdf": ("robertson", "BDF", 1e-6, 1e-6, 13, 3, K256),
    "robertson_dae_radau": ("robertson_dae", "RADAU", 1e-6, 1e-10, 14, 3, K256), # M y' = f, M = diag(1, 1, 0)
    "vdpstiff_radau": ("vdp_stiff", "RADAU", 1e-4, 1e-6, 5, 2, K256),
    "vdpstiff_bdf": ("vdp_stiff", "BDF", 1e-4, 1e-6, 5, 2, K256),
    "linear100_dopri5": ("linear100", "DOPRI5", 1e-6, 1e-8, 100, 100, 1 << 16),  # warp-per-trajectory kernels (n > 32)
    "medakzo_radau": ("medakzo", "RADAU", 1e-5, 1e-7, 900, 64, 1 << 12),         # warp-cooperative LU, n = 64
    "medakzo_bdf": ("medakzo", "BDF", 1e-5, 1e-7, 900, 64, 1 << 12),
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
    ens, method, rtol, atol, F, n, _ = WORKLOADS[name]
    n_te = N_T_EVAL.get(name, 0)
    extra = {"t_eval": np.linspace(t0, tf, n_te)} if n_te else {}
    if method == "RK4":
        extra["first_step"] = (tf - t0) / 1000.0
    if jac_mode:
        extra["jac_mode"] = 1
    extra.update(EXTRA_OPTIONS.get(name, {}))
    return Options(method=Method[method], rtol=rtol, atol=atol, flags=flags, max_events=1, **extra)


def sample_rows(n_total: int, n_sample: int, seed: int = 20261018) -> np.ndarray:
    """Seeded random sample of trajectory indices (sorted), the rows the CPU legs integrate."""
    if n_sample >= n_total:
        return np.arange(n_total)
    return np.sort(np.random.default_rng(seed).choice(n_total, size=n_sample, replace=False))


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled every ~2 ms
    from a thread so that even a 20 ms timed region carries samples; nvidia-smi -lms is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.nv, self.h = gpu_index, [], None, None, None
        self.stop_flag = False
        self.mx = None
        self.period = float(os.environ.get("IVPB_BENCH_POLL_MS", "2")) * 1e-3

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.time(), sm, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                sm = float(f[0]); self.mx = float(f[1])
            except Exception:
                continue
            self.rows.append((time.time(), sm, [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]))

    def window(self, t_begin: float, t_end: float) -> dict:
        rows = [r for r in self.rows if t_begin <= r[0] <= t_end]
        where = "timed region"
        if not rows:      # region shorter than one poll: the nearest samples around it
            rows = [r for r in self.rows if t_begin - 0.05 <= r[0] <= t_end + 0.05]
            where = "timed region +-50 ms"
        sm = [r[1] for r in rows]
        reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "window": where, "source": "nvml" if self.nv else "nvidia-smi"}

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()


def run_reference(args, rank, world):
    """Reference arm: the reference's own algorithm on the host cores (oracle port; the Rust crate cannot be
    compiled in this image), all hardware threads, each step a bounded sample of the workload."""
    if rank != 0:
        return
    from ivp_b200 import synth
    from ivp_b200.api import PROBLEMS
    from oracle import pyoracle
    ens, method, rtol, atol, F, n, size = WORKLOADS[args.workload]
    total = args.trajectories or size
    cores = pyoracle.hardware_threads()
    rows = sample_rows(total, args.cpu_sample)
    prob, y0, par, t0, tf = synth.ensemble(ens, total)
    y0 = np.ascontiguousarray(y0[rows])
    par = np.ascontiguousarray(par[rows]) if par is not None else None
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
    desc = (f"{len(rows)} trajectories/step: a seeded random sample of the {total}-trajectory {args.workload} ensemble, "
            f"std::thread x{cores}")
    print(json.dumps({
        "impl": "reference", "metric": "accepted_steps_per_sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "trajectories_per_step": int(len(rows)), "trajectories_total": total,
                   "method": method, "rtol": rtol, "atol": atol},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


class Arm:
    """One shard of the workload on this rank's GPU: device-resident and host-buffer (e2e) solves."""

    def __init__(self, args, ctx, dev, local_rank, Nper, offset, flags):
        import ctypes as C

        import torch

        from ivp_b200 import _abi, api, synth
        self.torch, self.ctx, self.dev, self.Nper = torch, ctx, dev, Nper
        ens, method, rtol, atol, F, n, _ = WORKLOADS[args.workload]
        prob_name, y0_h, par_h, t0, tf = synth.ensemble(ens, Nper, offset=offset)
        self.prob_name, self.y0_h, self.par_h, self.t0, self.tf = prob_name, y0_h, par_h, t0, tf
        self.problem = problem = api.Problem.builtin(prob_name)
        self.n_te = n_te = N_T_EVAL.get(args.workload, 0)
        self.opts = workload_options(args.workload, t0, tf, flags, args.jac_mode)
        self.mo = mo = _abi.MarshalledOptions(self.opts, problem.n, problem.n_events)
        self.ne = ne = problem.n_events
        # ---- device-resident buffers
        self.y0_d = torch.from_numpy(y0_h).to(dev)
        self.par_d = torch.from_numpy(par_h).to(dev) if par_h is not None else None
        self.status_d = torch.empty(Nper, dtype=torch.int32, device=dev)
        self.counters_d = torch.empty((Nper, 6), dtype=torch.int32, device=dev)
        self.tfin_d = torch.empty(Nper, dtype=torch.float64, device=dev)
        self.yfin_d = torch.empty((Nper, problem.n), dtype=torch.float64, device=dev)
        self.d_out = {"status": self.status_d.data_ptr(), "counters": self.counters_d.data_ptr(),
                      "t_final": self.tfin_d.data_ptr(), "y_final": self.yfin_d.data_ptr()}
        self.out_bytes_per_traj = 4 + 24 + 8 + 8 * problem.n
        self.nout_d = None
        if n_te:        # t_eval samples written straight to the preallocated [N][cap][n] block
            self.nout_d = torch.zeros(Nper, dtype=torch.int32, device=dev)
            self.yout_d = torch.zeros((Nper, mo.cap, problem.n), dtype=torch.float64, device=dev)
            self.d_out.update({"n_out": self.nout_d.data_ptr(), "y_out": self.yout_d.data_ptr()})
            self.out_bytes_per_traj += 4 + 8 * n_te * problem.n
        if ne:
            self.evc_d = torch.zeros((Nper, ne), dtype=torch.int32, device=dev)
            self.evt_d = torch.zeros((Nper, ne, 1), dtype=torch.float64, device=dev)
            self.d_out.update({"ev_count": self.evc_d.data_ptr(), "ev_t": self.evt_d.data_ptr()})
            self.out_bytes_per_traj += ne * 12
        # ---- pinned host buffers of the e2e arm
        self.y0_p = torch.from_numpy(y0_h).pin_memory()
        self.par_p = torch.from_numpy(par_h).pin_memory() if par_h is not None else None
        self.h_status = torch.empty(Nper, dtype=torch.int32).pin_memory()
        self.h_counters = torch.empty((Nper, 6), dtype=torch.int32).pin_memory()
        self.h_tfin = torch.empty(Nper, dtype=torch.float64).pin_memory()
        self.h_yfin = torch.empty((Nper, problem.n), dtype=torch.float64).pin_memory()
        st = _abi.IvpbOutputs()
        st.status = C.cast(self.h_status.data_ptr(), _abi.c_int32_p)
        st.counters = C.cast(self.h_counters.data_ptr(), _abi.c_uint32_p)
        st.t_final = C.cast(self.h_tfin.data_ptr(), _abi.c_double_p)
        st.y_final = C.cast(self.h_yfin.data_ptr(), _abi.c_double_p)
        self.y0_np, self.par_np = self.y0_p.numpy(), (self.par_p.numpy() if self.par_p is not None else None)
        self.h2d = self.y0_np.nbytes + (self.par_np.nbytes if self.par_np is not None else 0)
        self.d2h = self.h_status.numel() * 4 + self.h_counters.numel() * 4 + self.h_tfin.numel() * 8 + self.h_yfin.numel() * 8
        if n_te:
            self.h_nout = torch.empty(Nper, dtype=torch.int32).pin_memory()
            self.h_yout = torch.empty((Nper, mo.cap, problem.n), dtype=torch.float64).pin_memory()
            st.n_out = C.cast(self.h_nout.data_ptr(), _abi.c_int32_p)
            st.y_out = C.cast(self.h_yout.data_ptr(), _abi.c_double_p)
            self.d2h += self.h_nout.numel() * 4 + self.h_yout.numel() * 8
        if ne:
            self.h_evc = torch.empty((Nper, ne), dtype=torch.int32).pin_memory()
            self.h_evt = torch.empty((Nper, ne, 1), dtype=torch.float64).pin_memory()
            st.ev_count = C.cast(self.h_evc.data_ptr(), _abi.c_int32_p)
            st.ev_t = C.cast(self.h_evt.data_ptr(), _abi.c_double_p)
            self.d2h += self.h_evc.numel() * 4 + self.h_evt.numel() * 8
        self.st = st

    def solve_device(self):
        self.ctx.solve_device(self.problem, self.t0, self.tf, self.Nper, self.y0_d.data_ptr(),
                              self.par_d.data_ptr() if self.par_d is not None else None, self.mo, self.d_out,
                              stream=self.torch.cuda.current_stream().cuda_stream)

    def solve_host(self):
        self.ctx.solve_host(self.problem, self.t0, self.tf, self.y0_np, self.par_np, self.mo, self.st)


def timed_device(arm, steps, warmup, flush, barrier, sampler):
    """W untimed + K timed device-resident solves; per-step CUDA events, L2 flushed (untimed) before every step."""
    torch = arm.torch
    for _ in range(warmup):
        arm.solve_device()
    barrier()
    launches0 = arm.ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    wall0 = time.time()
    for s in range(steps):
        flush.fill_(s & 0xFF)                       # flush L2 between timed iterations (untimed)
        ev[s][0].record()
        arm.solve_device()
        ev[s][1].record()
    barrier()
    wall1 = time.time()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    return step_ms, arm.ctx.launch_count - launches0, sampler.window(wall0, wall1)


def timed_host(arm, steps, barrier):
    for _ in range(2):
        arm.solve_host()
    barrier()
    w0 = time.perf_counter()
    for _ in range(steps):
        arm.solve_host()                             # returns after the D2H copies completed
    return time.perf_counter() - w0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vdp_dop853", choices=sorted(WORKLOADS))
    ap.add_argument("--trajectories", type=int, default=None,
                    help="ensemble size (default: the workload's BASELINE.json size); strong scaling: in total, weak: per GPU")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: strong (default, the metric: the SAME ensemble on 1/2/4/8 GPUs) or weak (the size per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=32768, help="trajectories per step of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-context", action="store_true", help="N > 1: skip rank 0's one-context-over-all-devices leg")
    ap.add_argument("--static", action="store_true", help="disable work-queue refill (A/B)")
    ap.add_argument("--strict", action="store_true", help="force the -fmad=false kernel variant (A/B)")
    ap.add_argument("--fast", action="store_true", help="force the FMA-contracted kernel variant (A/B)")
    ap.add_argument("--fast-implicit", action="store_true", help="alias of --fast (RADAU / BDF workloads)")
    ap.add_argument("--no-zerocopy", action="store_true", help="e2e arm: staged H2D/D2H copies instead of mapped pinned buffers (A/B)")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e arm: copy the results out after the kernel instead of chunk by chunk while it runs (A/B)")
    ap.add_argument("--zerocopy-out", action="store_true", help="e2e arm: kernel stores results straight into the mapped pinned buffers (round-1 route, A/B)")
    ap.add_argument("--no-sort", action="store_true", help="RADAU / BDF: index order instead of the locality order of the ensemble (A/B)")
    ap.add_argument("--sort", action="store_true", help="explicit methods: locality order too (A/B)")
    ap.add_argument("--jac-mode", type=int, default=0, help="implicit workloads: 0 finite differences, 1 analytic")
    args = ap.parse_args()
    from ivp_b200.dist import dist_env, reduce_time_and_count, shard_range
    rank, local_rank, world = dist_env()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from ivp_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (ivp_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")      # host-side wait that leaves the GPUs idle (single-context leg)
    dev = torch.device("cuda", local_rank)

    ens, method, rtol, atol, F, n, size = WORKLOADS[args.workload]
    size = args.trajectories or size
    scaling = args.scaling or ("strong" if use_dist else "weak")
    if scaling == "strong":
        lo, hi = shard_range(size, rank, world)
        Nper, offset, total = hi - lo, lo, size
    else:
        Nper, offset, total = size, size * rank, size * world
    fast = args.fast or args.fast_implicit
    flags = (api.IVPB_FLAG_NO_REFILL if args.static else 0) | (api.IVPB_FLAG_STRICT_FP if args.strict else 0) | \
        (api.IVPB_FLAG_NO_ZEROCOPY if args.no_zerocopy else 0) | (api.IVPB_FLAG_NO_SORT if args.no_sort else 0) | \
        (api.IVPB_FLAG_SORT if args.sort else 0) | (api.IVPB_FLAG_FAST_FP if fast else 0) | \
        (api.IVPB_FLAG_NO_PIPELINE if args.no_pipeline else 0) | (api.IVPB_FLAG_ZEROCOPY_OUT if args.zerocopy_out else 0)
    ctx = api.Context([local_rank])
    arm = Arm(args, ctx, dev, local_rank, Nper, offset, flags)
    problem, ne, n_te = arm.problem, arm.ne, arm.n_te
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.05)

    # ---- device-resident arm -------------------------------------------------------------------
    step_ms, launches, clocks = timed_device(arm, args.steps, args.warmup, flush, barrier, sampler)
    total_ms = float(sum(step_ms))
    counters = arm.counters_d.cpu().numpy().view(np.uint32)
    status = arm.status_d.cpu().numpy()
    nstep, naccpt, nrejct = counters[:, 3], counters[:, 4], counters[:, 5]
    acc_local = int(naccpt.sum())
    n_samples = int(arm.nout_d.sum().item()) if n_te else 0
    flops_launch = algorithmic_flops(method, F, n, nstep, naccpt, dense=bool(n_te or ne), counters=counters,
                                     n_samples=n_samples)
    total_ms_max, acc_all = reduce_time_and_count(total_ms, acc_local, dev, use_dist)
    value = acc_all * args.steps / (total_ms_max * 1e-3)

    # ---- end-to-end arm: public host-buffer API, pinned host memory, copies inside the timed region ----
    e2e_steps = max(3, min(args.steps, 10))
    e2e_s = timed_host(arm, e2e_steps, barrier)
    e2e_acc = int(arm.h_counters.numpy().view(np.uint32)[:, 4].sum())
    e2e_s_max, e2e_acc_all = reduce_time_and_count(e2e_s, float(e2e_acc), dev, use_dist)
    e2e_value = e2e_acc_all * e2e_steps / e2e_s_max
    assert e2e_acc == acc_local, "host-buffer and device-buffer paths disagree"
    e2e_rank_ms = None
    if use_dist:      # per-rank e2e times (diagnosis of host-side contention at N = 8)
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = e2e_s / e2e_steps * 1e3
        dist.all_reduce(t)
        e2e_rank_ms = [round(float(x), 3) for x in t.tolist()]
    fp_mode = ctx.last_fp_mode() if hasattr(ctx, "last_fp_mode") else None
    if isinstance(fp_mode, dict) and fp_mode.get("mode") == "strict":
        fp_mode["second_pass_trajectories"] = ctx.last_reruns()      # strictd kernels: deferred division guards (ivpb_exact.cuh)

    # ---- N > 1: the other scaling flavour, same JSON line ----
    other = None
    if use_dist:
        if scaling == "strong":
            N2, off2 = size, size * rank
        else:
            lo2, hi2 = shard_range(size, rank, world)
            N2, off2 = hi2 - lo2, lo2
        del arm.y0_d, arm.yfin_d
        arm2 = Arm(args, ctx, dev, local_rank, N2, off2, flags)
        ms2, _, clocks2 = timed_device(arm2, args.steps, args.warmup, flush, barrier, sampler)
        acc2 = int(arm2.counters_d.cpu().numpy().view(np.uint32)[:, 4].sum())
        ms2_max, acc2_all = reduce_time_and_count(float(sum(ms2)), acc2, dev, use_dist)
        e2 = timed_host(arm2, e2e_steps, barrier)
        e2_max, _ = reduce_time_and_count(e2, 0.0, dev, use_dist)
        other = {"scaling": "weak" if scaling == "strong" else "strong", "trajectories_per_gpu": N2,
                 "trajectories_total": N2 * world if scaling == "strong" else size,
                 "value": acc2_all * args.steps / (ms2_max * 1e-3), "ms_per_step": ms2_max / args.steps,
                 "e2e_value": acc2_all * e2e_steps / e2_max, "e2e_ms_per_step": e2_max / e2e_steps * 1e3, "clocks": clocks2}
        del arm2

    # ---- N > 1, rank 0: ONE context over all N devices (the library's own multi-GPU path) ----
    single = None
    if use_dist and not args.no_single_context:
        # The other ranks must leave their GPUs IDLE while rank 0 drives all N devices from one context: a rank parked in
        # an NCCL barrier keeps a spinning all-reduce kernel resident, and the two processes' contexts are then time-sliced
        # on that GPU (measured at N = 2: 17.3 ms instead of 7.6 ms per step).  They wait on the host (gloo) instead.
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            try:
                single = single_context_leg(args, world, size, flags, e2e_steps, acc_all if scaling == "strong" else None)
            except Exception as e:      # reported, never fatal for the headline
                single = {"error": repr(e)}
        dist.barrier(group=cpu_group)
        barrier()

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
    alg_bytes = Nper * (8 * problem.n + 8 * problem.p + arm.out_bytes_per_traj)
    roofline = {"bound": "fp64", "achieved": achieved, "peak": NOMINAL_FP64_TFLOPS, "unit": "TFLOP/s",
                "frac": achieved / NOMINAL_FP64_TFLOPS, "traffic": traffic,
                "peak_source": "nominal: 148 SM x 64 DFMA/clk x 2 x 1.965 GHz (MEASURED_PEAKS.json has no fp64 entry)",
                "peak_measured": peak_meas, "frac_measured": achieved / peak_meas if peak_meas else None,
                "peak_measured_source": "libivpb DFMA-chain microbenchmark on this GPU (ncu capture: profiles/r2_dfma_peak_ncu.txt)",
                "flops_per_launch": flops_launch, "kernel_ms": kernel_ms,
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9}}

    # ---- CPU baseline + in-run parity (rank 0, N=1 only): oracle port on all host threads, seeded random sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from ivp_b200.api import PROBLEMS
        from oracle import pyoracle
        cores = pyoracle.hardware_threads()
        rows = sample_rows(Nper, args.cpu_sample)
        y0_s = np.ascontiguousarray(arm.y0_h[rows])
        par_s = np.ascontiguousarray(arm.par_h[rows]) if arm.par_h is not None else None
        best = 0.0
        for _ in range(2):
            c0 = time.perf_counter()
            o = pyoracle.solve_batch(PROBLEMS[arm.prob_name], arm.t0, arm.tf, y0_s, par_s, arm.opts, nthreads=cores,
                                     want=["status", "counters", "y_final"])
            dtc = time.perf_counter() - c0
            best = max(best, float(o.naccpt.sum()) / dtc)
        parity = float(np.mean((o.naccpt == naccpt[rows]) & (o.nrejct == nrejct[rows])))
        yf = arm.yfin_d.cpu().numpy()[rows]
        tol = np.maximum(10 * rtol * np.abs(o.y_final), 10 * atol)
        in_tol = float(np.mean(np.all(np.abs(yf - o.y_final) <= tol, axis=1)))
        cpu = {"value": best, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"seeded random sample of {len(rows)} of the {Nper} trajectories, std::thread x{cores}, best of 2",
               "step_count_parity_on_sample": parity, "in_tolerance_on_sample": in_tol,
               "status_equal_on_sample": bool(np.array_equal(o.status, status[rows])),
               "bit_identical_y_final_on_sample": float(np.mean(np.all(yf == o.y_final, axis=1)))}

    if rank == 0:
        implicit = method in ("RADAU", "BDF")
        print(json.dumps({
            "metric": "accepted_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "problem": arm.prob_name, "method": method, "rtol": rtol, "atol": atol,
                       "t_span": [arm.t0, arm.tf], "trajectories_per_gpu": Nper, "trajectories_total": total,
                       "outputs": "final state + status + counters" + (f" + {n_te} t_eval samples" if n_te else "") +
                                  (" + event times" if ne else ""), "parallelism": f"trajectory-sharded x{world}",
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "schedule": ("static" if args.static else "work-queue refill") + (", locality order" if (args.sort or (implicit and not args.no_sort)) else ""),
                       "fp": fp_mode or ("strict" if args.strict else "fma" if fast else "default")},
            "accepted_steps_per_step": acc_all, "rejected_steps_rank0": int(nrejct.sum()),
            "status_success_frac_rank0": float(np.mean(status == 0)),
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": int(arm.h2d), "d2h_bytes_per_step": int(arm.d2h),
                    "ms_per_step": e2e_s_max / e2e_steps * 1e3, "ms_per_step_by_rank": e2e_rank_ms,
                    "api": "ivpb_solve_batch (pinned host buffers, H2D + D2H inside the timed region)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            ("weak" if scaling == "strong" else "strong"): other, "single_context": single,
        }))
    sampler.stop()
    if use_dist:
        dist.destroy_process_group()


def single_context_leg(args, world, size, flags, steps, expect_acc):
    """Rank 0 drives ONE ivpb_ctx over all `world` devices: the library's static split + per-device work queues, host
    buffers (ivpb_solve_batch) and device-0-resident buffers with peer scatter / gather (ivpb_solve_batch_device)."""
    import ctypes as C

    import torch

    from ivp_b200 import _abi, api, synth
    ens, method, rtol, atol, F, n, _ = WORKLOADS[args.workload]
    prob_name, y0_h, par_h, t0, tf = synth.ensemble(ens, size)
    problem = api.Problem.builtin(prob_name)
    opts = workload_options(args.workload, t0, tf, flags, args.jac_mode)
    mo = _abi.MarshalledOptions(opts, problem.n, problem.n_events)
    ctx = api.Context(list(range(world)))
    y0_p = torch.from_numpy(y0_h).pin_memory()
    par_p = torch.from_numpy(par_h).pin_memory() if par_h is not None else None
    h_status = torch.empty(size, dtype=torch.int32).pin_memory()
    h_counters = torch.empty((size, 6), dtype=torch.int32).pin_memory()
    h_tfin = torch.empty(size, dtype=torch.float64).pin_memory()
    h_yfin = torch.empty((size, problem.n), dtype=torch.float64).pin_memory()
    st = _abi.IvpbOutputs()
    st.status = C.cast(h_status.data_ptr(), _abi.c_int32_p)
    st.counters = C.cast(h_counters.data_ptr(), _abi.c_uint32_p)
    st.t_final = C.cast(h_tfin.data_ptr(), _abi.c_double_p)
    st.y_final = C.cast(h_yfin.data_ptr(), _abi.c_double_p)
    y0_np, par_np = y0_p.numpy(), (par_p.numpy() if par_p is not None else None)
    for _ in range(3):
        ctx.solve_host(problem, t0, tf, y0_np, par_np, mo, st)
    w0 = time.perf_counter()
    for _ in range(steps):
        ctx.solve_host(problem, t0, tf, y0_np, par_np, mo, st)
    host_ms = (time.perf_counter() - w0) / steps * 1e3
    acc = int(h_counters.numpy().view(np.uint32)[:, 4].sum())
    # device-0-resident buffers, results gathered by peer copies
    dev0 = torch.device("cuda", 0)
    with torch.cuda.device(dev0):
        y0_d = y0_p.to(dev0)
        par_d = par_p.to(dev0) if par_p is not None else None
        status_d = torch.empty(size, dtype=torch.int32, device=dev0)
        counters_d = torch.empty((size, 6), dtype=torch.int32, device=dev0)
        tfin_d = torch.empty(size, dtype=torch.float64, device=dev0)
        yfin_d = torch.empty((size, problem.n), dtype=torch.float64, device=dev0)
        d_out = {"status": status_d.data_ptr(), "counters": counters_d.data_ptr(), "t_final": tfin_d.data_ptr(),
                 "y_final": yfin_d.data_ptr()}
        s = torch.cuda.current_stream(dev0)

        def go():
            ctx.solve_device(problem, t0, tf, size, y0_d.data_ptr(), par_d.data_ptr() if par_d is not None else None,
                             mo, d_out, stream=s.cuda_stream)
        for _ in range(3):
            go()
        torch.cuda.synchronize(dev0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(steps):
            go()
        e1.record(s)
        torch.cuda.synchronize(dev0)
        dev_ms = e0.elapsed_time(e1) / steps
        acc_d = int(counters_d.cpu().numpy().view(np.uint32)[:, 4].sum())
    ctx.close()
    return {"devices": world, "trajectories_total": size,
            "host_buffers": {"api": "ivpb_solve_batch", "ms_per_step": host_ms, "value": acc / (host_ms * 1e-3)},
            "device_buffers": {"api": "ivpb_solve_batch_device (inputs / outputs on device 0, peer scatter + gather)",
                               "ms_per_step": dev_ms, "value": acc_d / (dev_ms * 1e-3)},
            "results_equal_per_process_arm": (None if expect_acc is None else bool(acc == int(expect_acc) and acc_d == int(expect_acc)))}


if __name__ == "__main__":
    main()
