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
E2E_CALLS = int(os.environ.get("CB2_BENCH_E2E_CALLS", "5"))   # end-to-end Optimize() calls per run; the median is reported
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
    """Tolerances disabled (negative: Ceres's tests are `<=`, so that an exactly-zero cost change at convergence does not stop the run either):
    exactly `iters` LM iterations are run."""
    return _capi_or_oracle_opts(max_num_iterations=iters, function_tolerance=-1.0, gradient_tolerance=-1.0, parameter_tolerance=-1.0,
                                min_trust_region_radius=0.0, minimizer_progress_to_stdout=0, **kw)


def run_cpu(config, frames, iters, threads, warm=True):
    """The reference's CPU path restated (oracle/: dual-number autodiff in 4-wide passes, Ceres-style automatic DENSE_SCHUR ordering -> dense
    reduced system, blocked multi-threaded Cholesky) on `frames` frames of `config` (frames == the config's own count: the full workload).
    Runs `iters` LM iterations from the perturbed initial guess with the tolerances disabled. Returns per-iteration host-clock times."""
    from oracle import oracle_py
    from calico_b200 import synthetic
    full = FULL_FRAMES.get(config, frames)
    cfg = synthetic.CONFIGS[config] if frames == full else sample_config(config, frames)
    t0 = time.perf_counter()
    truth, prob = synthetic.generate(cfg, oracle_py.oracle_api, noise=True)
    t_gen = time.perf_counter() - t0
    nblocks = prob.counts()[0]
    if warm:   # thread pool / page-in warm-up on a tiny problem (untimed)
        _, wp = synthetic.generate("tiny", oracle_py.oracle_api, noise=True)
        wa = oracle_py.oracle_api()
        wp.push(wa)
        wa.optimize(bench_options(oracle_py.OracleOptions, 1, linear_solver=2, num_threads=threads))
        wa.close()
    api = oracle_py.oracle_api()
    prob.clone().push(api)
    t0 = time.perf_counter()
    summ, log = api.optimize(bench_options(oracle_py.OracleOptions, iters, linear_solver=2, num_threads=threads))
    dt = time.perf_counter() - t0
    api.close()
    n = max(len(log) - 1, 1)
    it_times = [it.iteration_time for it in log[1:]]
    steady = sum(it_times) / max(len(it_times), 1) if it_times else dt
    return {"frames": frames, "full_frames": full, "blocks": nblocks, "n": n, "seconds": dt, "generate_s": t_gen,
            "it_per_s_loop": n / dt,                      # iterations / whole LM loop (initial evaluation included) — how the GPU arm is timed
            "it_per_s_steady": 1.0 / steady,              # 1 / mean host-clock time of iterations 1.. (initial evaluation excluded)
            "initial_eval_s": log[0].iteration_time if log else None, "iteration_s": it_times,
            "jacobian_time": summ.jacobian_time, "linear_solver_time": summ.linear_solver_time, "final_cost": summ.final_cost}


def describe_cpu(r, threads):
    what = "the FULL workload" if r["frames"] == r["full_frames"] else f"{r['frames']} of {r['full_frames']} frames"
    return (f"{what} ({r['blocks']} residual blocks), {r['n']} LM iteration(s) of the restated Ceres DENSE_SCHUR path (oracle/, linear_solver=2) on {threads} "
            f"host thread(s): {r['seconds']:.2f} s for the LM loop incl. the initial evaluation ({r['initial_eval_s']:.2f} s); measured, not extrapolated")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=os.environ.get("CB2_BENCH_CONFIG", "C4"))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=int(os.environ.get("CB2_CPU_SAMPLE_FRAMES", "0")),
                    help="0 (default): the CPU legs run the full workload; > 0: that many frames of it")
    ap.add_argument("--ref-max-steps", type=int, default=int(os.environ.get("CB2_REF_MAX_STEPS", "3")),
                    help="--impl reference times min(--steps, this) full-workload LM iterations (one costs ~10-20 s of all host cores on C4)")
    ap.add_argument("--ref-extras", type=int, default=int(os.environ.get("CB2_REF_EXTRAS", "1")),
                    help="--impl reference: also time a 1-thread run and two smaller samples (scaling exponent)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1
    workload = {"C1": "1 OpenCv5 camera, 50 frames, intrinsics only", "C2": "1 OpenCv5 camera + IMU, 500 frames",
                "C3": "4 KannalaBrandt cameras, 2000 frames", "C4": "8 OpenCv5 cameras + IMU, 5000 frames, 25 corners/image",
                "C5": "16 OpenCv5 cameras + 2 IMUs, 10000 frames, Huber"}.get(args.config, args.config)

    full_frames = FULL_FRAMES.get(args.config, 0)
    cpu_frames = args.cpu_sample_frames if args.cpu_sample_frames > 0 else full_frames
    if args.impl == "reference":
        if rank != 0:
            return 0
        # The reference's own CPU path for this metric, on THIS workload (not a sample): min(--steps, --ref-max-steps) full LM iterations on all
        # host cores. `steps` / `ms_per_step` report what was actually timed.
        k = max(1, min(args.steps, args.ref_max_steps))
        r = run_cpu(args.config, cpu_frames, k, threads)
        extras = {}
        if args.ref_extras:
            # Calico's actual default is num_threads = 1 (batch_optimizer.cpp:10-17 never sets it); the samples show how the dense reduced
            # solve makes the CPU path scale super-linearly with the trajectory length.
            small = [f for f in (max(full_frames // 20, 50), max(full_frames // 5, 100)) if f < cpu_frames]
            samples = [run_cpu(args.config, f, 2, threads, warm=False) for f in small]
            one = run_cpu(args.config, small[-1] if small else cpu_frames, 1, 1, warm=False)
            extras = {"samples_all_cores": [{"frames": q["frames"], "residual_blocks": q["blocks"], "it_per_s": q["it_per_s_loop"], "it_per_s_steady": q["it_per_s_steady"]} for q in samples],
                      "one_thread": {"frames": one["frames"], "residual_blocks": one["blocks"], "it_per_s": one["it_per_s_loop"], "it_per_s_steady": one["it_per_s_steady"],
                                     "note": "num_threads = 1 is what calico::DefaultSolverOptions leaves Ceres at"}}
        value = r["it_per_s_loop"]
        sample = describe_cpu(r, threads)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": r["n"], "warmup": min(args.warmup, 1),
                "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{args.config}: {workload}", "residual_blocks": r["blocks"], "requested_steps": args.steps},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "it_per_s_steady": r["it_per_s_steady"],
                                 "jacobian_s": r["jacobian_time"], "linear_solver_s": r["linear_solver_time"], "generate_s": r["generate_s"], **extras},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    # The repetition count is decided ONCE (rank 0's wall time of the timed region) and broadcast: run_plan contains collectives, so every
    # rank must run it the same number of times (a per-rank wall-clock loop can leave one rank inside an allreduce its peers never join).
    reps = max(1, min(400, int(0.6 / max(wall, 1e-3)) + 1))
    if world > 1:
        t_reps = torch.tensor([reps], device="cuda", dtype=torch.int64)
        dist.broadcast(t_reps, src=0)
        reps = int(t_reps.item())
    for _ in range(reps):
        run_plan(plan)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "%d untimed repeats of the timed plan (~0.6 s) right after the timed region" % reps
    if world > 1:
        dist.barrier()

    # ---- end to end through the C ABI with host buffers: assembly + H2D + solve + D2H write-back ----
    # A calibration session calls Optimize repeatedly (outlier marking -> re-optimise, camera.cpp:258-299); each call builds a new
    # problem from the caller's host arrays. The timed call below is such a repeat call: the handle of the kernel-timed run above has been
    # closed, so its device blocks sit in the library's memory pool. Every host->device byte of the problem is copied inside the timed region.
    # The call is made E2E_CALLS times, each on a fresh handle and a fresh copy of the caller's arrays, and the MEDIAN call is reported (all
    # samples are in the line): a single call is at the mercy of the host (page faults of freshly cloned arrays, scheduling of the
    # per-sensor packing threads) — single samples between 35 and 70 ms were seen on identical builds.
    samples = []
    api_prev = api
    for _ in range(E2E_CALLS):
        p2 = prob.clone()          # the caller's own host arrays (Python-side copy, not part of the API call)
        api2 = gpu_api()
        if world > 1:
            api2.comm_clone(api_prev)   # communicator creation is one-time process setup, not part of a solve
        api_prev.close()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ids2 = p2.push(api2)
        # Exactly --steps LM iterations in ONE Optimize() call: past convergence (~6 iterations on this problem) LM keeps producing steps whose
        # model cost change is at rounding level; they are still solved and evaluated, so the invalid-step limit is lifted to let the call run on.
        summ2, log2 = api2.optimize(bench_options(_capi.Options, args.steps, max_num_consecutive_invalid_steps=args.steps + 1))
        p2.pull(api2, ids2)
        p2.residuals(api2, ids2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        samples.append((dt, api2.stats(), max(len(log2) - 1, 1), summ2.message.decode(errors="replace")))
        api_prev = api2
    api_prev.close()
    e2e_all_ms = [round(1e3 * x[0], 3) for x in samples]
    e2e_s, st2, iters2, e2e_msg = sorted(samples, key=lambda x: x[0])[len(samples) // 2]

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
    gram_bytes = st.camera_kernel_gram_bytes / max(st.jacobian_sweeps, 1)
    achieved_all = (jac_bytes + gram_bytes) / (jac_ms * 1e-3) / 1e9 if jac_ms > 0 else 0.0
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
                     "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture (a constant of the build, not measured by this run)",
                     "algorithmic_bytes_per_launch": jac_bytes, "ms_per_launch": jac_ms,
                     "all_outputs": {"what": "the same launch also forms the cameras' normal-equation blocks on the FP64 tensor pipe (what accumulate_kernel re-read the "
                                             "Jacobian for in round 1) and writes them as compact per-image Gram slots + per-warp calibration partials",
                                     "extra_bytes_per_launch": gram_bytes, "achieved": achieved_all, "frac": achieved_all / peak},
                     "whole_sweep": {"kernels": "K0 frames + K1 camera + K2 gyroscope + K3 accelerometer + cost reduction", "achieved": sweep_achieved,
                                     "frac": sweep_achieved / peak, "algorithmic_bytes": sweep_bytes, "ms": sweep_ms}},
        "clocks": clocks,
        "e2e": {"value": iters2 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": st2.h2d_bytes / iters2, "d2h_bytes_per_step": st2.d2h_bytes / iters2,
                "seconds": e2e_s, "ms_per_optimize_call": 1e3 * e2e_s, "iterations": iters2, "termination": e2e_msg, "calls_ms": e2e_all_ms,
                "what": "assembly from host arrays (cb2_set_trajectory / cb2_add_*) + upload + cb2_optimize + parameter/residual write-back on a fresh problem handle; median of %d such calls (calls_ms lists them all)" % E2E_CALLS},
        "gpu_launches": int(st.kernel_launches),
    }
    if not args.no_cpu_baseline:
        # One LM iteration of the FULL workload (about 10-30 s of all host cores on C4): measured, not extrapolated. `value` is the steady
        # iteration rate (1 / time of iteration 1; the initial evaluation is reported beside it).
        r = run_cpu(args.config, cpu_frames, 1, threads)
        line["cpu_baseline"] = {"value": r["it_per_s_steady"], "unit": UNIT, "cores": threads, "kind": "port", "sample": describe_cpu(r, threads),
                                "it_per_s_incl_initial_evaluation": r["it_per_s_loop"], "jacobian_s": r["jacobian_time"], "linear_solver_s": r["linear_solver_time"]}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
