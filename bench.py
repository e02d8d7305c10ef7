#!/usr/bin/env python
"""bench.py — LM iterations/s and residual-block Jacobian evaluations/s of the calico_b200 hot path.

A "step" is one Levenberg-Marquardt iteration of calico::BatchOptimizer::Optimize's solve (reference
calico/batch_optimizer.cpp:73) on a synthetic problem of the BASELINE.json shape named in --config (default C4:
8 cameras + IMU, 5000 frames, ~1.1 M residual blocks): residual+Jacobian sweep, normal equations, Schur solve, step update
and trial-cost evaluation. The timed region runs EXACTLY --steps iterations from the perturbed initial guess with the
convergence tolerances disabled (so that K iterations are always run), after --warmup untimed iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl ours|reference]

Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lm_iterations_per_sec"
UNIT = "LM iterations/s"
FULL_FRAMES = {"C1": 50, "C2": 500, "C3": 2000, "C4": 5000, "C5": 10000}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def sample_config(name, frames):
    from calico_b200 import synthetic
    import dataclasses
    base = synthetic.CONFIGS[name]
    return dataclasses.replace(base, name=f"{name}_sample{frames}", n_frames=frames)


def bench_options(_capi_or_oracle_opts, iters, **kw):
    """Tolerances disabled: exactly `iters` LM iterations are run."""
    return _capi_or_oracle_opts(max_num_iterations=iters, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0,
                                min_trust_region_radius=0.0, minimizer_progress_to_stdout=0, **kw)


def run_cpu(config, frames, steps, warmup, threads):
    """The reference's CPU path restated (oracle/, Ceres-style automatic DENSE_SCHUR ordering) on a bounded sample."""
    from oracle import oracle_py
    from calico_b200 import synthetic
    cfg = sample_config(config, frames)
    truth, prob = synthetic.generate(cfg, oracle_py.oracle_api, noise=True)
    nblocks = prob.counts()[0]

    def one(iters):
        api = oracle_py.oracle_api()
        prob.clone().push(api)
        t0 = time.perf_counter()
        summ, log = api.optimize(bench_options(oracle_py.OracleOptions, iters, linear_solver=2, num_threads=threads))
        dt = time.perf_counter() - t0
        n = max(len(log) - 1, 1)
        api.close()
        return dt, n, summ
    if warmup > 0:
        one(min(warmup, 1))
    dt, n, summ = one(steps)
    scale = frames / FULL_FRAMES.get(config, frames)
    it_per_s_sample = n / dt
    return {"value": it_per_s_sample * scale, "sample_it_per_s": it_per_s_sample, "sample_blocks": nblocks, "n": n, "seconds": dt,
            "jacobian_time": summ.jacobian_time, "linear_solver_time": summ.linear_solver_time, "scale": scale}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=os.environ.get("CB2_BENCH_CONFIG", "C4"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=int(os.environ.get("CB2_CPU_SAMPLE_FRAMES", "250")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    workload = {"C1": "1 OpenCv5 camera, 50 frames, intrinsics only", "C2": "1 OpenCv5 camera + IMU, 500 frames",
                "C3": "4 KannalaBrandt cameras, 2000 frames", "C4": "8 OpenCv5 cameras + IMU, 5000 frames, 25 corners/image",
                "C5": "16 OpenCv5 cameras + 2 IMUs, 10000 frames, Huber"}.get(args.config, args.config)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_cpu(args.config, args.cpu_sample_frames, args.steps, args.warmup, threads)
        sample = (f"{args.cpu_sample_frames} of {FULL_FRAMES.get(args.config, args.cpu_sample_frames)} frames of {args.config} "
                  f"({r['sample_blocks']} residual blocks), {r['n']} LM iterations in {r['seconds']:.2f} s; iterations/s scaled linearly by the frame "
                  f"ratio to the full workload (favours the CPU: its dense reduced solve grows cubically)")
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 / r["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.config}: {workload}", "cpu_sample": sample},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: calico_b200 has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from calico_b200 import _capi, build, synthetic
    build.build()
    lib = _capi.LIB_PATH

    def gpu_api():
        a = _capi.CApi(lib)
        a.set_device(local_rank)
        return a

    t_gen = time.perf_counter()
    truth, prob = synthetic.generate(args.config, gpu_api, noise=True)
    t_gen = time.perf_counter() - t_gen
    nblocks, nres = prob.counts()

    api = gpu_api()
    if world > 1:
        api.comm_init_torch(world, rank)
    prob.clone().push(api)
    api.upload()
    # EXACTLY --steps LM iterations are timed. From the perturbed initial guess this problem converges to rounding level in ~6 iterations,
    # after which LM only produces cheap invalid/rejected steps; so the K timed iterations are run as solves of at most SOLVE_ITERS
    # iterations, each restarted from the initial guess (cb2_reset_parameters, outside the timed LM loops). Every timed iteration is
    # therefore a full one (linear solve + step + trial cost + Jacobian sweep), and every solve pays its initial Jacobian evaluation too.
    SOLVE_ITERS = 5
    plan = [SOLVE_ITERS] * (args.steps // SOLVE_ITERS) + ([args.steps % SOLVE_ITERS] if args.steps % SOLVE_ITERS else [])

    def run_plan(counts):
        logs = []
        for n_it in counts:
            api.reset_parameters()
            summ_, log_ = api.optimize(bench_options(_capi.Options, n_it))
            logs.append((summ_, log_))
        return logs
    if args.warmup > 0:
        run_plan([min(args.warmup, SOLVE_ITERS)] * ((args.warmup + SOLVE_ITERS - 1) // SOLVE_ITERS))
    api.stats_reset()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    logs = run_plan(plan)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    st = api.stats()
    summ = logs[-1][0]
    iters = sum(max(len(lg) - 1, 0) for _, lg in logs)
    iters = max(iters, 1)     # normally == --steps; a solve that stops early (consecutive invalid steps) is reported with what was actually timed
    loop_ms = st.lm_loop_ms
    if world > 1:
        t = torch.tensor([loop_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loop_ms = float(t.item())
    value = iters / (loop_ms * 1e-3)
    accepted = sum(1 for _, lg in logs for it in lg[1:] if it.step_is_successful)
    # Clocks under load: the timed region lasts ~15 ms, below nvidia-smi's 100 ms sampling period, so the same plan is repeated (untimed)
    # for ~0.6 s with the sampler running (B200_PROFILING.md clocks line).
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.6:
        run_plan(plan)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "untimed repeats of the timed plan for 0.6 s right after the timed region"
    if world > 1:
        dist.barrier()

    # ---- end to end through the C ABI with host buffers: assembly + H2D + solve + D2H write-back ----
    # A calibration session calls Optimize repeatedly (outlier marking -> re-optimise, camera.cpp:258-299); each call builds a new
    # problem from the caller's host arrays. The timed call below is such a repeat call: the handle of the kernel-timed run above has been
    # closed, so its device blocks sit in the library's memory pool. Every host->device byte of the problem is copied inside the timed region.
    p2 = prob.clone()          # the caller's own host arrays (Python-side copy, not part of the API call)
    api2 = gpu_api()
    if world > 1:
        api2.comm_clone(api)   # communicator creation is one-time process setup, not part of a solve
    api.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    ids2 = p2.push(api2)
    summ2, log2 = api2.optimize(bench_options(_capi.Options, args.steps))
    p2.pull(api2, ids2)
    for sid in ids2:
        api2.get_residuals(sid)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    st2 = api2.stats()
    iters2 = max(len(log2) - 1, 1)
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    api2.close()

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peak, peak_src = peaks()
    # Roofline of the dominant kernel: the camera residual + analytic-Jacobian kernel eval_kernel<camera, Jacobian>, timed alone with
    # CUDA events on the library's stream (it runs serially on that stream); the whole K0-K3 sweep is reported beside it.
    sweep_ms = st.jacobian_kernel_ms / max(st.jacobian_sweeps, 1)
    sweep_bytes = st.jacobian_bytes / max(st.jacobian_sweeps, 1)
    sweep_achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    if st.camera_kernel_ms > 0:
        jac_ms = st.camera_kernel_ms / max(st.jacobian_sweeps, 1)
        jac_bytes = st.camera_kernel_bytes / max(st.jacobian_sweeps, 1)
        kernel_name = "eval_kernel<camera, Jacobian> (K1: camera residual + analytic Jacobian, one launch per sweep)"
    else:   # no cameras in this workload: the sweep as a whole
        jac_ms, jac_bytes, kernel_name = sweep_ms, sweep_bytes, "eval_kernel<*, Jacobian> (K1-K3 sweep)"
    achieved = jac_bytes / (jac_ms * 1e-3) / 1e9 if jac_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.config)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": iters, "warmup": args.warmup,
        "ms_per_step": loop_ms / iters, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {workload}", "residual_blocks": nblocks, "residuals": nres, "control_points": int(prob.spline.ctrl.shape[0]),
                   "lm_iterations_accepted": accepted, "lm_iterations_rejected": iters - accepted, "requested_steps": args.steps,
                   "l2": "inputs larger than L2: the Jacobian written and re-read every iteration is %.0f MB" % (sweep_bytes / 1e6),
                   "timing": "CUDA events on the library's stream around the LM loops, max over ranks; %d solves of <= %d iterations from the initial guess" % (len(plan), SOLVE_ITERS), "wall_s": wall, "generate_s": t_gen,
                   "final_cost": summ.final_cost, "initial_cost": summ.initial_cost},
        "jacobian_evals_per_sec": st.jacobian_blocks / (st.jacobian_kernel_ms * 1e-3) if st.jacobian_kernel_ms > 0 else None,
        "phases_ms_per_iteration": {"jacobian_sweep": st.jacobian_kernel_ms / iters, "normal_equations": st.normal_eq_ms / iters,
                                    "schur_solve_and_update": st.schur_ms / iters, "trial_cost": st.cost_eval_ms / iters,
                                    "jacobian_sweeps": st.jacobian_sweeps},
        "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": jac_bytes, "ms_per_launch": jac_ms,
                     "whole_sweep": {"kernels": "K0 frames + K1 camera + K2 gyroscope + K3 accelerometer + cost reduction", "achieved": sweep_achieved,
                                     "frac": sweep_achieved / peak, "algorithmic_bytes": sweep_bytes, "ms": sweep_ms}},
        "clocks": clocks,
        "e2e": {"value": iters2 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": st2.h2d_bytes / iters2, "d2h_bytes_per_step": st2.d2h_bytes / iters2,
                "seconds": e2e_s, "what": "assembly from host arrays (cb2_set_trajectory / cb2_add_*) + upload + cb2_optimize + parameter/residual write-back on a fresh problem handle"},
        "gpu_launches": int(st.kernel_launches),
    }
    if not args.no_cpu_baseline:
        r = run_cpu(args.config, args.cpu_sample_frames, 3, 1, threads)
        line["cpu_baseline"] = {
            "value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"{args.cpu_sample_frames} of {FULL_FRAMES.get(args.config, args.cpu_sample_frames)} frames of {args.config} ({r['sample_blocks']} residual blocks), "
                       f"{r['n']} LM iterations in {r['seconds']:.2f} s with the restated Ceres DENSE_SCHUR path; iterations/s scaled linearly by the frame ratio "
                       f"(favours the CPU)"),
            "sample_it_per_s": r["sample_it_per_s"]}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
