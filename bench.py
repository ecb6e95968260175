#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native ar_slam solver.

A "step" is one Levenberg-Marquardt iteration (Jacobian evaluation +
accumulation + Schur/linear solve + candidate cost + accept/reject) of the
bundle adjustment on a synthetic map.  Default
workload: BASELINE config 3, 100k captures x 5k tags, ~8 tags per capture
(3.2 M observation corners), the shape the metric is quoted on.

  python bench.py --gpus N --steps K --warmup W      (torchrun for N > 1)
  python bench.py --impl reference ...               CPU restatement of the reference's Ceres path

value  = observation corners processed per second by the whole job
         (corners x LM iterations / device time, CUDA events on the solver's
         stream, max over ranks), problem resident in HBM.
e2e    = the same through the C-ABI with host buffers: set_problem +
         set_params + solve + get_params inside the timed region.

With N > 1 the NAMED shape is split across the ranks (strong scaling, the
default); `extra.weak` carries a short weak-scaling measurement (every rank
brings the workload's captures against the same map).  `n_gpu_check` compares
the sharded solve's cost with a single-GPU solve of the same problem.
At N = 1 short runs of the other BASELINE configurations ride along in
`extra_workloads`.
"""
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

WORKLOADS = {
    # name: (captures, tags, tags per capture, BASELINE config)
    "ba_100k_5k": (100000, 5000, 8, 3),
    "ba_1k_200": (1000, 200, 8, 2),
    "ba_20k_2k": (20000, 2000, 8, 5),
    "loc_1m_5k": (1000000, 5000, 8, 4),   # batched localisation (kernel 5), a step = one full batch
    # marker detection, the step before the path: (frames per batch, -, markers per frame, -);
    # a step = one batch of 1020 x 768 BGR frames (the demo images' size) through arslam_detect_markers
    "detect_1020x768": (64, 0, 8, 0),
}
ITERS_PER_SOLVE = 5  # LM iterations per solve call; the solve restarts from the same initial state
INIT_STATE = ("ground truth perturbed by N(0, 2 cm) / N(0, 2 deg) per pose component, f0 = 800 px (true 760); "
              "NOT the reference's f0 = 3000 + heuristic seeds (those are exercised by the schedule tests)")
# what bounds each kernel; bytes are the compulsory HBM bytes the library reports per launch
KERNEL_BOUND = {"dense_cholesky": "tensor"}
KERNEL_NOTE = {
    "accum_E": "fused evaluation + accumulation, J never in HBM: 18 B/corner in + 72 B/corner W out + 264 B/pose",
    "accum_F": "tag-sorted pass, Jacobians recomputed: 18 B/corner in + 264 B/pose out",
    "schur_eliminate": "W read once (72 B/corner) + E records + the lower reduced blocks written once; limited by FP64 reductions into L2",
    "pcg_solve": "persistent cooperative PCG, matrix resident in shared memory: compulsory HBM bytes = the block values "
                 "read once per solve; bound by grid-barrier + halo latency per iteration, not by bandwidth",
    "backsub": "W read once (72 B/corner) + E records",
    "candidate": "residual-only cost at x + delta: 18 B/corner",
    "eval_jacobian": "kernel (1) with J materialised (arslam_evaluate, the Problem::Evaluate parity API; not on the LM path, which "
                     "fuses it into accum_E/F): 18 B/corner in + 16 B r + 240 B J out = 274 B/corner, stores staged per warp",
    "localize": "whole LM solve per capture in registers: FP64-latency bound, not HBM bound",
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def dgemm_peak():
    fp = os.path.join(ROOT, "profiles", "r1_fp64_peaks.json")
    if os.path.exists(fp):
        with open(fp) as f:
            return max(json.load(f).values()), "measured cuBLAS DGEMM on this pool's B200 (profiles/r1_fp64_peaks.json)"
    return 35.5, "cuBLAS DGEMM measured in round 1"


def static_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (static:
    the profiler cannot run inside a timed bench)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            with open(tp) as f:
                v = json.load(f).get(workload, {}).get(kernel)
            if v is not None:
                return v, "profiles/" + name
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a
    thread (the timed region is a few tens of ms, too short for nvidia-smi's 100 ms loop), with
    nvidia-smi as the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40))

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index
        self.nvml, self.handle, self.stop_flag, self.thread = None, None, False, None
        self.sm, self.mx, self.mask = [], None, 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            reasons = [name for name, bit in self.BITS if self.mask & bit]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


class Ctx:
    """What every measurement needs: the torch / distributed handles and this rank's place in the job."""

    def __init__(self, args):
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = self.dist = self.ar = self.synth = None

    def init_gpu(self):
        import torch
        import ar_slam_b200 as ar
        from ar_slam_b200 import synth
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.torch, self.ar, self.synth = torch, ar, synth
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def pinned(self, a):
        torch = self.torch
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.ascontiguousarray(a[:0])).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t

    def comm_init(self, s):
        torch, dist = self.torch, self.dist
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            uid = torch.frombuffer(bytearray(self.ar.Solver.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        s.comm_init(self.rank, self.world, bytes(uid.cpu().numpy().tobytes()))


def bench_options(ar, iters, args, linear_solver=None, num_intrinsics=None, pcg_tolerance=None, exact_iters=True):
    o = ar.default_options()
    o.max_num_iterations = iters
    if exact_iters:   # run exactly `iters` LM iterations: switch the convergence tests off
        o.function_tolerance = 0.0
        o.parameter_tolerance = 0.0
        o.gradient_tolerance = 0.0
    linear_solver = linear_solver or args.linear_solver
    num_intrinsics = num_intrinsics or args.num_intrinsics
    if linear_solver == "dense":
        o.linear_solver = ar.LINSOLVE_DENSE
        o.dense_max_dim = 1 << 20
    elif linear_solver == "pcg":
        o.linear_solver = ar.LINSOLVE_PCG
    o.pcg_tolerance = args.pcg_tolerance if pcg_tolerance is None else pcg_tolerance
    o.num_intrinsics = num_intrinsics
    return o


def shard(m, rank, world):
    """Contiguous capture ranges balanced by block count (all blocks of a capture on one rank)."""
    if world == 1:
        return m.cap_idx, m.tag_idx, m.obs
    nb = len(m.cap_idx)
    counts = np.bincount(m.cap_idx, minlength=m.n_cap)
    cum = np.concatenate([[0], np.cumsum(counts)])
    cuts = [int(np.searchsorted(cum, nb * r / world)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, m.n_cap
    lo, hi = cum[cuts[rank]], cum[cuts[rank + 1]]
    return m.cap_idx[lo:hi], m.tag_idx[lo:hi], m.obs[lo:hi]


def run_solves(s, total_iters, start):
    """Runs solves of ITERS_PER_SOLVE iterations from the same start until total_iters are done.
    start = (cam, cap, tag) pinned copies of the initial state."""
    cam0, cap0, tag0 = start
    done, summaries = 0, []
    while done < total_iters:
        n = min(ITERS_PER_SOLVE, total_iters - done)
        if s.options.max_num_iterations != n:
            s.options.max_num_iterations = n
            s.set_options(s.options)
        s.set_params(cam0, cap0, tag0)
        summ, _ = s.solve(log=False)
        if summ["iterations"] != n:
            raise RuntimeError("solve stopped after %d of %d iterations (%s)" % (summ["iterations"], n, summ["reason_name"]))
        summaries.append(summ)
        done += n
    return summaries


def cpu_baseline(args, workload, steps, warmup, threads=None):
    """Times the oracle (restated reference, not Ceres) on a bounded sample of the workload."""
    from ar_slam_b200 import synth
    from oracle import pyoracle as po
    n_cap, n_tag, tpc, _ = WORKLOADS[workload]
    scale = max(1, n_cap // args.cpu_sample_captures)
    sc, st = max(50, n_cap // scale), max(10, n_tag // scale)
    m = synth.make_map(sc, st, tpc, seed=0xA55A0000 + 100)
    # torchrun exports OMP_NUM_THREADS=1: ask the OS for the cores this process may use instead
    threads = threads or len(os.sched_getaffinity(0))
    o = po.default_options(num_threads=threads, max_num_iterations=1, function_tolerance=0.0,
                           parameter_tolerance=0.0, gradient_tolerance=0.0)
    nc = 4 * len(m.cap_idx)

    def run(iters):
        o.max_num_iterations = iters
        t = time.perf_counter()
        _, _, _, summ, _ = po.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0, options=o)
        return time.perf_counter() - t, summ
    if warmup:
        run(min(warmup, 2))
    dt, summ = run(steps)
    its = max(1, summ["iterations"])
    same = (sc == n_cap and st == n_tag)
    return {"value": nc * its / dt, "unit": "corners/s", "cores": threads, "kind": "port",
            "lm_iters_per_sec": its / dt, "reduced_dim": summ["reduced_dim"], "same_config": same,
            "sample": "%d captures x %d tags (%d corners), same generator and density as %s, %d LM iterations, "
                      "restated reference (not Ceres): Jet autodiff + dense Schur, OpenMP x%d%s"
                      % (sc, st, nc, workload, its, threads,
                         "" if same else "; NOT the named shape: the reference's DENSE_SCHUR is cubic in the tag count "
                                         "(reduced dimension %d here, %d at the named shape), so corners/s at the named "
                                         "shape would be lower still" % (summ["reduced_dim"], 6 * n_tag + 1))}, dt, its


# --------------------------------------------------------------------------------------------- roofline
def kernel_rooflines(kt, reduced_dim, workload):
    """One entry per profiled kernel: achieved = compulsory bytes (or flops) per launch / CUDA-event time per launch."""
    hbm_peak, hbm_how = read_peaks()
    tot = sum(k["total_ms"] for k in kt) or 1.0
    out = []
    for k in kt:
        if k["launches"] <= 0 or k["total_ms"] <= 0:
            continue
        sec = k["total_ms"] / k["launches"] * 1e-3
        share = k["total_ms"] / tot
        name = k["name"]
        if KERNEL_BOUND.get(name) == "tensor":
            n = reduced_dim
            flops = n ** 3 / 3.0 + 2.0 * n * n
            peak, how = dgemm_peak()
            e = {"kernel": name, "bound": "tensor", "achieved": flops / sec / 1e12, "peak": peak, "unit": "TFLOP/s",
                 "frac": flops / sec / 1e12 / peak, "peak_source": how, "algorithmic_flops": flops,
                 "note": "FP64 DMMA (mma.sync.m8n8k4.f64) blocked Cholesky + solve, n^3/3 + 2 n^2 flop, n = %d; the peak is "
                         "FP64 DGEMM, not the bf16 figure of MEASURED_PEAKS.json" % n}
        else:
            b = k["algorithmic_bytes"]
            if b <= 0:
                continue
            e = {"kernel": name, "bound": "hbm", "achieved": b / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": b / sec / 1e9 / hbm_peak, "peak_source": hbm_how, "algorithmic_bytes": b}
            if name in KERNEL_NOTE:
                e["note"] = KERNEL_NOTE[name]
        e["us_per_launch"] = sec * 1e6
        e["launches"] = k["launches"]
        e["share_of_step"] = share
        tr, src = static_traffic(workload, name)
        e["traffic"] = tr
        if tr is not None:
            e["traffic_source"] = src + " (ncu --set full, static)"
        out.append(e)
    out.sort(key=lambda e: -e["share_of_step"])
    return out


# --------------------------------------------------------------------------------------------------- BA
def bench_ba(ctx, workload, steps, warmup, scaling="strong", linear_solver=None, num_intrinsics=None,
             e2e=True, profile=True, converge=False, check=False, clocks=True):
    """One bundle-adjustment workload on the job's GPUs.  Returns the fields of a bench line (rank 0) or None."""
    args, torch, ar, synth = ctx.args, ctx.torch, ctx.ar, ctx.synth
    rank, world = ctx.rank, ctx.world
    n_cap, n_tag, tpc, cfg_id = WORKLOADS[workload]
    num_intrinsics = num_intrinsics or args.num_intrinsics
    if scaling == "weak":
        n_cap *= world
    m = synth.make_map(n_cap, n_tag, tpc, seed=0xA55A0000 + cfg_id,
                       distortion=(-0.05, 0.01) if num_intrinsics == 3 else (0.0, 0.0))
    cap_idx, tag_idx, obs = shard(m, rank, world)
    n_corner_total = 4 * len(m.cap_idx)

    opts = bench_options(ar, ITERS_PER_SOLVE, args, linear_solver, num_intrinsics)
    s = ar.Solver(device=ctx.local_rank, options=opts)
    for kv in args.tune:
        k, v = kv.split("=")
        s.set_tuning(k, int(v))
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    if world > 1:
        ctx.comm_init(s)
    s.set_problem(m.n_cap, m.n_tag, cap_idx, tag_idx, obs)

    # every solve restarts from the same initial state, re-sent from pinned host memory (the only
    # host -> device traffic inside the timed region: 4.9 MB per 5 LM iterations)
    keep_start = [ctx.pinned(a) for a in (m.cam0, m.cap0, m.tag0)]
    start = tuple(t.numpy() for t in keep_start)
    if warmup > 0:
        run_solves(s, warmup, start)
    # ---- timed region: exactly `steps` LM iterations, CUDA events on the solver's stream
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and clocks:
        sampler.start()
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        summaries = run_solves(s, steps, start)
        ev1.record(stream)
    ctx.barrier()
    ms, = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    clk = sampler.stop() if (rank == 0 and clocks) else None
    launches = int(sum(x["gpu_launches"] for x in summaries))
    pcg_its = int(sum(x["linear_solver_iterations"] for x in summaries))
    final_cost = summaries[-1]["final_cost"]

    # ---- end to end through the C-ABI with host buffers
    e2e_out = None
    if e2e:
        # what ArSlamSolver::optimize does per call (ar_slam_util.cpp:1001-1018 with the blocks of
        # :720-727): the whole problem and the parameters go host -> device from pinned host memory,
        # ITERS_PER_SOLVE LM iterations run, the parameters come back.  One untimed call first.
        keep = [ctx.pinned(a) for a in (cap_idx, tag_idx, obs, m.cam0, m.cap0, m.tag0)]
        p_cap_idx, p_tag_idx, p_obs, p_cam0, p_cap0, p_tag0 = [t.numpy() for t in keep]
        if s.options.max_num_iterations != ITERS_PER_SOLVE:
            s.options.max_num_iterations = ITERS_PER_SOLVE
            s.set_options(s.options)
        keep_out = [torch.empty((m.n_cap, 6), dtype=torch.float64, pin_memory=True),
                    torch.empty((m.n_tag, 6), dtype=torch.float64, pin_memory=True)]
        out = tuple(t.numpy() for t in keep_out)

        def one_call():
            s.set_problem(m.n_cap, m.n_tag, p_cap_idx, p_tag_idx, p_obs)
            s.set_params(p_cam0, p_cap0, p_tag0)
            summ, _ = s.solve(log=False)
            s.get_params(out=out)
            return summ
        one_call()
        n_solves = max(1, -(-steps // ITERS_PER_SOLVE))
        ctx.barrier()
        t0 = time.perf_counter()
        e2e_iters = 0
        for _ in range(n_solves):
            e2e_iters += one_call()["iterations"]
        ctx.barrier()
        dt, = ctx.max_over_ranks(time.perf_counter() - t0)
        param_bytes = 8 * (3 + 6 * m.n_cap + 6 * m.n_tag)
        h2d = n_solves * (obs.nbytes + cap_idx.nbytes + tag_idx.nbytes + param_bytes)
        d2h = n_solves * param_bytes + 8 * 40 * e2e_iters
        e2e_out = {"value": n_corner_total * e2e_iters / dt, "unit": "corners/s",
                   "h2d_bytes_per_step": int(h2d / e2e_iters), "d2h_bytes_per_step": int(d2h / e2e_iters),
                   "lm_iters_per_sec": e2e_iters / dt, "lm_iterations": e2e_iters,
                   "what": "%d x (arslam_set_problem, set_params, solve of %d LM iterations, get_params) from pinned host "
                           "arrays, wall clock; a step is one LM iteration" % (n_solves, ITERS_PER_SOLVE)}

    # ---- per-kernel device times (separate profiled solve; events around every launch)
    rooflines = None
    if profile:
        if rank == 0:
            s.set_profiling(True)
        run_solves(s, ITERS_PER_SOLVE, start)
        if rank == 0:
            kt = s.kernel_times()
            s.set_profiling(False)
            rooflines = kernel_rooflines(kt, summaries[0]["reduced_dim"], workload)

    # ---- kernel (1) on its own: residuals + Jacobians materialised, as ceres::Problem::Evaluate would return them
    eval_roofline = None
    if profile and world == 1 and rank == 0 and n_corner_total <= 4000000:
        s.set_params(*start)
        s.set_profiling(True)
        times = []
        for _ in range(3):
            s.evaluate(jacobians=True)
            times += [k for k in s.kernel_times() if k["name"] == "eval_jacobian"]
        s.set_profiling(False)
        best = min(times, key=lambda k: k["total_ms"])
        hbm_peak, hbm_how = read_peaks()
        sec = best["total_ms"] * 1e-3
        eval_roofline = {"kernel": "eval_jacobian", "bound": "hbm", "achieved": best["algorithmic_bytes"] / sec / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": best["algorithmic_bytes"] / sec / 1e9 / hbm_peak, "peak_source": hbm_how,
                         "algorithmic_bytes": best["algorithmic_bytes"], "us_per_launch": sec * 1e6, "launches": 3,
                         "share_of_step": 0.0, "traffic": static_traffic(workload, "eval_jacobian")[0], "note": KERNEL_NOTE["eval_jacobian"]}

    # ---- time to converge with the reference's own stopping rule (function_tolerance 1e-6, <= 50 iterations)
    conv = None
    if converge:
        conv = {}
        runs = [("default", None, None)]
        if summaries[0]["linear_solver"] == 2 and world == 1:
            runs.append(("pcg_tight_1e-8", "pcg", 1e-8))
        for label, ls, tol in runs:
            o = bench_options(ar, 50, args, ls or linear_solver, num_intrinsics, tol, exact_iters=False)
            s.set_options(o)
            s.set_params(*start)
            s.solve(log=False)   # warm (symbolic phase of a solver switch)
            ctx.barrier()
            with torch.cuda.stream(stream):
                s.set_params(*start)
                ev0.record(stream)
                summ, _ = s.solve(log=False)
                ev1.record(stream)
            ctx.barrier()
            cms, = ctx.max_over_ranks(ev0.elapsed_time(ev1))
            conv[label] = {"ms": cms, "lm_iterations": summ["iterations"], "final_cost": summ["final_cost"],
                           "termination": summ["reason_name"], "pcg_iterations": int(summ["linear_solver_iterations"]),
                           "pcg_tolerance": o.pcg_tolerance if summ["linear_solver"] == 2 else None}
        s.set_options(opts)

    # ---- N-GPU gate: the sharded solve must reproduce the single-GPU cost (tight linear solves, 2 LM iterations)
    gate = None
    if check and world > 1:
        tight = bench_options(ar, 2, args, "pcg" if summaries[0]["linear_solver"] == 2 else "dense", num_intrinsics, 1e-11)
        s.set_options(tight)
        s.set_params(*start)
        summ_n, _ = s.solve(log=False)
        ctx.barrier()
        if rank == 0:
            s1 = ar.Solver(device=ctx.local_rank, options=tight)
            s1.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
            s1.set_params(m.cam0, m.cap0, m.tag0)
            summ_1, _ = s1.solve(log=False)
            s1.close()
            gate = {"cost_%dgpu" % world: summ_n["final_cost"], "cost_1gpu": summ_1["final_cost"],
                    "cost_vs_1gpu_rel": abs(summ_n["final_cost"] - summ_1["final_cost"]) / summ_1["final_cost"],
                    "what": "2 LM iterations from the bench's initial state, PCG tolerance 1e-11, same problem solved "
                            "sharded over %d ranks and on rank 0's GPU alone" % world}
        ctx.barrier()
        s.set_options(opts)
    s.close()
    if rank != 0:
        return None
    value = n_corner_total * steps / (ms * 1e-3)
    lin = {1: "dense_cholesky_dmma", 2: "pcg"}[summaries[0]["linear_solver"]]
    res = {"metric": "observation_corners_per_sec", "value": value, "unit": "corners/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "lm_iters_per_sec": steps / (ms * 1e-3),
           "config": {"workload": workload, "captures": m.n_cap, "tags": m.n_tag, "tags_per_capture": tpc,
                      "blocks": int(len(m.cap_idx)), "corners": int(n_corner_total), "linear_solver": lin,
                      "pcg_tolerance": opts.pcg_tolerance if lin == "pcg" else None,
                      "intrinsics": "f, l1, l2 (radial model)" if num_intrinsics == 3 else "f (focal only, the reference's live model)",
                      "eliminated": {1: "tags", 2: "captures"}[summaries[0]["eliminated_side"]],
                      "reduced_dim": summaries[0]["reduced_dim"], "iters_per_solve": ITERS_PER_SOLVE,
                      "initial_state": INIT_STATE,
                      "l2_policy": "inputs larger than L2 (W + observations > 126 MB) for ba_100k_5k; no explicit flush",
                      "final_cost": final_cost, "pcg_iterations": pcg_its,
                      "pcg_iterations_per_lm_iteration": pcg_its / max(1, steps) if lin == "pcg" else None},
           "clocks": clk, "e2e": e2e_out, "gpu_launches": launches}
    if rooflines:
        res["roofline"] = dict(rooflines[0])
        res["roofline"]["why_this_kernel"] = "largest share of the LM iteration"
        res["rooflines"] = [r for r in rooflines if r["share_of_step"] >= 0.05] + ([eval_roofline] if eval_roofline else [])
        res["kernels"] = {r["kernel"]: {"us_per_launch": r["us_per_launch"], "launches": r["launches"],
                                        "share_of_step": r["share_of_step"]} for r in rooflines}
    if conv:
        res["time_to_converge"] = conv
    if gate:
        res["n_gpu_check"] = gate
    return res


# ------------------------------------------------------------------------------------------ localisation
def bench_localization(ctx, workload, steps, warmup, cpu=True, clocks=True):
    """BASELINE config 4: captures are sharded evenly, no communication at all."""
    args, torch, ar, synth = ctx.args, ctx.torch, ctx.ar, ctx.synth
    rank, world = ctx.rank, ctx.world
    n_loc, n_tag, tpc, cfg_id = WORKLOADS[workload]
    m = synth.make_localization_batch(n_loc, n_tag, tpc, seed=0xA55A0000 + cfg_id)
    lo, hi = n_loc * rank // world, n_loc * (rank + 1) // world
    b0, b1 = m.blk_offsets[lo], m.blk_offsets[hi]
    offs = (m.blk_offsets[lo:hi + 1] - b0).astype(np.int32)
    tag_idx, obs, seed = m.tag_idx[b0:b1], m.obs[b0:b1], m.seed_block[lo:hi]
    s = ar.Solver(device=ctx.local_rank)
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    keep = [ctx.pinned(a) for a in (offs, tag_idx, obs, seed)]   # the batch arrives in pinned host memory
    offs, tag_idx, obs, seed = [t.numpy() for t in keep]
    keep_out = [torch.empty((hi - lo, 6), dtype=torch.float64, pin_memory=True), torch.empty(hi - lo, dtype=torch.int32, pin_memory=True),
                torch.empty(hi - lo, dtype=torch.float64, pin_memory=True), torch.empty(hi - lo, dtype=torch.int32, pin_memory=True)]
    out = tuple(t.numpy() for t in keep_out)
    for _ in range(max(1, warmup)):
        pose, its, cost, term = s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and clocks:
        sampler.start()
    # ---- value: the kernel alone, inputs resident (one chunk: no copy or other chunk runs beside it)
    s.set_tuning("loc_chunk", hi - lo)
    s.set_profiling(True)
    kms, launches = [], 0
    ctx.barrier()
    for _ in range(steps):
        pose, its, cost, term = s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
        kt = s.kernel_times()
        kms.append(sum(k["total_ms"] for k in kt if k["name"] == "localize"))
        launches += int(sum(k["launches"] for k in kt))
    s.set_profiling(False)
    # ---- e2e: the call as a user makes it (chunked upload / kernel / download pipeline on three streams)
    s.set_tuning("loc_chunk", 0)
    s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pose, its, cost, term = s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
    ctx.barrier()
    dt = time.perf_counter() - t0
    dt, ms_kernel = ctx.max_over_ranks(dt, float(np.sum(kms)))
    clk = sampler.stop() if (rank == 0 and clocks) else None
    res = None
    if rank == 0:
        nc = 4 * len(m.tag_idx)
        peak, how = read_peaks()
        err = np.abs(pose - m.cap_true[lo:hi])
        cb = None
        if cpu and world == 1:
            from oracle import pyoracle as po
            ns = min(n_loc, 50000)
            threads = len(os.sched_getaffinity(0))
            t1 = time.perf_counter()
            po.localize_batch(m.blk_offsets[:ns + 1], m.tag_idx[:m.blk_offsets[ns]], m.obs[:m.blk_offsets[ns]],
                              m.seed_block[:ns], m.cam_true, m.tag_true, num_threads=threads)
            dtc = time.perf_counter() - t1
            cb = {"value": 4 * int(m.blk_offsets[ns]) / dtc, "unit": "corners/s", "cores": threads, "kind": "port",
                  "captures_per_sec": ns / dtc, "same_config": ns == n_loc,
                  "sample": "first %d captures of the same batch, restated localizeOne (not Ceres), OpenMP x%d" % (ns, threads)}
        alg = 17.0 * 4 * len(tag_idx) + 116.0 * (hi - lo)
        ach = alg * steps / (ms_kernel * 1e-3) / 1e9
        res = {"metric": "observation_corners_per_sec", "value": nc * steps / (ms_kernel * 1e-3), "unit": "corners/s",
               "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_kernel / steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "captures_per_sec": n_loc * steps / (ms_kernel * 1e-3),
               "config": {"workload": workload, "captures": n_loc, "tags": n_tag, "tags_per_capture": tpc,
                          "corners": nc, "mean_lm_iterations": float(its.mean()),
                          "median_abs_pose_error": float(np.median(err)), "l2_policy": "inputs larger than L2",
                          "sharding": "captures split evenly over the ranks, no communication"},
               "clocks": clk, "gpu_launches": launches,
               "e2e": {"value": nc * steps / dt, "unit": "corners/s", "captures_per_sec": n_loc * steps / dt,
                       "h2d_bytes_per_step": int(obs.nbytes + tag_idx.nbytes + offs.nbytes + seed.nbytes),
                       "d2h_bytes_per_step": int(pose.nbytes + its.nbytes + cost.nbytes + term.nbytes),
                       "what": "arslam_localize_batch from pinned host arrays (H2D, validation + kernel, D2H, in 128 k-capture "
                               "chunks on three streams), wall clock"},
               "roofline": {"bound": "hbm", "kernel": "localize", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": static_traffic(workload, "localize")[0], "peak_source": how,
                            "algorithmic_bytes": alg, "note": KERNEL_NOTE["localize"]},
               "cpu_baseline": cb}
    s.close()
    return res


def bench_detection(ctx, workload, steps, warmup, cpu=True, clocks=True):
    """cv::aruco::detectMarkers replacement (aruco_detector.cpp:106, ar_slam_util.cpp:268) on batches of frames; the
    frames are split evenly over the ranks, no communication."""
    torch, ar, synth = ctx.torch, ctx.ar, ctx.synth
    from ar_slam_b200 import capi
    rank, world = ctx.rank, ctx.world
    n_frames, _, n_markers, _ = WORKLOADS[workload]
    w, h = 1020, 768
    bits = synth.dict_4x4_50_bits()
    distinct = [synth.render_marker_scene(h, w, bits, n_markers, 0xA55A0600 + k, noise=4.0)[0] for k in range(8)]
    mine = list(range(n_frames * rank // world, n_frames * (rank + 1) // world))
    host = torch.empty((len(mine), h, w, 3), dtype=torch.uint8, pin_memory=True)     # frames arrive in pinned host memory
    for i, f in enumerate(mine):
        host[i] = torch.from_numpy(distinct[f % 8])
    frames = host.numpy()
    params = capi.default_detect_params(min_corner_distance_rate=0.1)                # ar_slam_util.cpp:250
    det = capi.Detector(len(mine), w, h, device=ctx.local_rank)
    dev = host.cuda()
    torch.cuda.synchronize()
    for _ in range(max(1, warmup)):
        res = det.detect(None, params, device_ptr=dev.data_ptr(), shape=dev.shape)
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and clocks:
        sampler.start()
    # ---- value: frames already in HBM, device time of the whole call (all stages + the candidate table's D2H)
    stage = {"threshold_ms": 0.0, "starts_ms": 0.0, "follow_ms": 0.0, "approx_ms": 0.0, "identify_ms": 0.0, "total_ms": 0.0}
    launches = 0
    ctx.barrier()
    for _ in range(steps):
        res = det.detect(None, params, device_ptr=dev.data_ptr(), shape=dev.shape)
        t = det.times()
        for k in stage:
            stage[k] += t[k]
        launches += t["launches"]
    # ---- e2e: the call as a user makes it, frames in pinned host memory, ids and corners back on the host
    det.detect(frames, params)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = det.detect(frames, params)
    ctx.barrier()
    dt = time.perf_counter() - t0
    dt, ms_dev = ctx.max_over_ranks(dt, stage["total_ms"])
    clk = sampler.stop() if (rank == 0 and clocks) else None
    out = None
    if rank == 0:
        peak, how = read_peaks()
        found = int(sum(len(ids) for ids, _ in res))
        px = n_frames * w * h
        cb = None
        if cpu and world == 1:
            ns = 16
            t1 = time.perf_counter()
            try:
                import cv2
                p = cv2.aruco.DetectorParameters()
                p.minCornerDistanceRate = 0.1
                cvd = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), p)
                same = True
                for i in range(ns):
                    r, ids, _ = cvd.detectMarkers(frames[i])
                    got_ids, got_c = res[i]
                    same = same and ids is not None and list(ids.ravel()) == list(got_ids) and \
                        all((a.reshape(4, 2) == b).all() for a, b in zip(r, got_c))
                dtc = time.perf_counter() - t1
                cb = {"value": ns / dtc, "unit": "frames/s", "cores": int(cv2.getNumThreads()), "kind": "reference",
                      "identical_to_gpu_result": bool(same), "same_config": True,
                      "sample": "first %d frames of the batch through cv2.aruco.ArucoDetector.detectMarkers (OpenCV %s, "
                                "the library the reference calls), same parameters" % (ns, cv2.__version__)}
            except ImportError:
                from oracle import aruco_detect
                for i in range(2):
                    aruco_detect.detect_markers(frames[i], bits, 1)
                dtc = time.perf_counter() - t1
                cb = {"value": 2 / dtc, "unit": "frames/s", "cores": 1, "kind": "port", "same_config": True,
                      "sample": "2 frames through the numpy restatement (cv2 not importable)"}
        alg = 5.0 * len(mine) * w * h                       # 3 B BGR in + 1 B grey + 1 B threshold bits out per pixel
        ach = alg * steps / (stage["threshold_ms"] * 1e-3) / 1e9
        out = {"metric": "frames_per_sec", "value": n_frames * steps / (ms_dev * 1e-3), "unit": "frames/s",
               "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_dev / steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
               "pixels_per_sec": px * steps / (ms_dev * 1e-3),
               "config": {"workload": workload, "frames_per_batch": n_frames, "width": w, "height": h, "channels": 3,
                          "markers_placed_per_frame": n_markers, "markers_found": found,
                          "dictionary": "DICT_4X4_50", "l2_policy": "inputs larger than L2 (150 MB of frames per batch)",
                          "frames": "8 distinct rendered scenes (synth.render_marker_scene), repeated",
                          "sharding": "frames split evenly over the ranks, no communication"},
               "clocks": clk, "gpu_launches": int(launches),
               "stages_ms_per_step": {k: v / steps for k, v in stage.items()},
               "e2e": {"value": n_frames * steps / dt, "unit": "frames/s", "pixels_per_sec": px * steps / dt,
                       "h2d_bytes_per_step": int(frames.nbytes), "d2h_bytes_per_step": int(found * 36 + 4 * len(mine)),
                       "what": "arslam_detect_markers on pinned host frames (H2D, five kernels, candidate table D2H, "
                               "host-side grouping), wall clock"},
               "roofline": {"bound": "hbm", "kernel": "gray_threshold", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": (static_traffic(workload, "gray_threshold")[0] or 0) * len(mine) / 64.0 or None,
                            "traffic_source": "profiles/r2_traffic.json (ncu --set full of a 64-frame launch, static)",
                            "peak_source": how, "algorithmic_bytes": alg,
                            "us_per_launch": 1e3 * stage["threshold_ms"] / steps,
                            "share_of_step": stage["threshold_ms"] / stage["total_ms"],
                            "note": "grey conversion + three adaptive thresholds from one shared-memory integral tile: "
                                    "3 B/pixel in, 2 B/pixel out; the border following that dominates the step is a "
                                    "chain of dependent byte loads (latency bound), see stages_ms_per_step"},
               "cpu_baseline": cb}
    det.close()
    return out


def short(res, keys=("ms_per_step", "value", "unit", "lm_iters_per_sec", "captures_per_sec", "config", "e2e", "roofline",
                     "gpu_launches", "steps", "warmup", "time_to_converge", "pixels_per_sec", "stages_ms_per_step",
                     "cpu_baseline")):
    return {k: res[k] for k in keys if res and k in res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ba_100k_5k", choices=sorted(WORKLOADS))
    ap.add_argument("--linear-solver", default="auto", choices=["auto", "dense", "pcg"])
    ap.add_argument("--pcg-tolerance", type=float, default=0.1)
    ap.add_argument("--num-intrinsics", type=int, default=1, choices=[1, 3],
                    help="3: BASELINE config 5's radial model (l1 = -0.05, l2 = 0.01 in the data, f, l1, l2 all free)")
    ap.add_argument("--cpu-sample-captures", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads / extra.weak / time_to_converge / n_gpu_check")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="kernel-variant switch passed to arslam_set_tuning (A/B measurements), e.g. accum_pipe=0")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the named workload is split over the GPUs; weak: every GPU gets the workload's "
                         "captures (the map, i.e. the tags, is shared)")
    args = ap.parse_args()
    ctx = Ctx(args)
    n_cap, n_tag, tpc, cfg_id = WORKLOADS[args.workload]

    if args.impl == "reference":
        if ctx.rank != 0:
            return 0
        if args.workload.startswith("detect_"):
            # the reference's own implementation of this step IS OpenCV (aruco_detector.cpp:106): time it directly
            import cv2
            from ar_slam_b200 import synth
            bits = synth.dict_4x4_50_bits()
            frames = [synth.render_marker_scene(768, 1020, bits, 8, 0xA55A0600 + k, noise=4.0)[0] for k in range(8)]
            p = cv2.aruco.DetectorParameters()
            p.minCornerDistanceRate = 0.1
            cvd = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), p)
            per_step = 16
            for _ in range(args.warmup):
                cvd.detectMarkers(frames[0])
            steps = min(args.steps, 10)
            t0 = time.perf_counter()
            for _ in range(steps):
                for i in range(per_step):
                    cvd.detectMarkers(frames[i % 8])
            dt = time.perf_counter() - t0
            v = per_step * steps / dt
            cb = {"value": v, "unit": "frames/s", "cores": int(cv2.getNumThreads()), "kind": "reference",
                  "sample": "%d frames per step (the batch is 64) through cv2.aruco.ArucoDetector.detectMarkers, OpenCV %s"
                            % (per_step, cv2.__version__)}
            print(json.dumps({"impl": "reference", "metric": "frames_per_sec", "value": v, "unit": "frames/s",
                              "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
                              "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
                              "data": "synthetic", "config": {"workload": args.workload, "width": 1020, "height": 768,
                                                              "frames_per_step": per_step},
                              "cpu_baseline": cb,
                              "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        cb, dt, its = cpu_baseline(args, args.workload, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "observation_corners_per_sec", "value": cb["value"], "unit": "corners/s",
                "n_gpus": args.gpus, "steps": its, "warmup": args.warmup, "ms_per_step": 1e3 * dt / its,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "lm_iters_per_sec": cb["lm_iters_per_sec"], "same_config": cb["same_config"],
                "config": {"workload": args.workload, "captures": n_cap, "tags": n_tag, "tags_per_capture": tpc,
                           "note": "CPU arm runs the bounded sample described in cpu_baseline.sample"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "corners/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    ctx.init_gpu()
    extra_on = not args.no_extra
    if args.workload.startswith("detect_"):
        line = bench_detection(ctx, args.workload, min(args.steps, 20), min(args.warmup, 5), cpu=not args.no_cpu_baseline)
    elif args.workload.startswith("loc_"):
        line = bench_localization(ctx, args.workload, args.steps, args.warmup, cpu=not args.no_cpu_baseline)
    else:
        line = bench_ba(ctx, args.workload, args.steps, args.warmup, scaling=args.scaling, e2e=not args.no_e2e,
                        converge=extra_on, check=extra_on)
        if ctx.rank == 0 and not args.no_cpu_baseline and ctx.world == 1:
            line["cpu_baseline"], _, _ = cpu_baseline(args, args.workload, 2, 1)
        elif ctx.rank == 0:
            line["cpu_baseline"] = None
        if extra_on and ctx.world > 1 and args.scaling == "strong":
            # short weak-scaling measurement: every rank brings the workload's captures against the same map
            w = bench_ba(ctx, args.workload, min(args.steps, 20), min(args.warmup, 5), scaling="weak", e2e=False,
                         profile=False, clocks=False)
            if ctx.rank == 0:
                line["extra"] = {"weak": short(w)}
        if extra_on and ctx.world == 1 and args.workload == "ba_100k_5k":
            # the other BASELINE configurations, short runs, so that one invocation witnesses them all
            ex = {}

            def extra(name, fn):
                # an extra workload must never take the headline line down with it
                try:
                    ex[name] = short(fn())
                except Exception as e:  # noqa: BLE001
                    ex[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            extra("ba_1k_200", lambda: bench_ba(ctx, "ba_1k_200", 40, 10, e2e=True, clocks=False, converge=True))
            extra("ba_20k_2k_radial_dense", lambda: bench_ba(ctx, "ba_20k_2k", 5, 3, linear_solver="dense", num_intrinsics=3,
                                                             e2e=False, converge=True, clocks=False))
            extra("ba_20k_2k_radial_pcg", lambda: bench_ba(ctx, "ba_20k_2k", 20, 5, linear_solver="pcg", num_intrinsics=3,
                                                           e2e=False, clocks=False, converge=True))
            extra("loc_1m_5k", lambda: bench_localization(ctx, "loc_1m_5k", 3, 2, cpu=False, clocks=False))
            extra("detect_1020x768", lambda: bench_detection(ctx, "detect_1020x768", 5, 2, cpu=True, clocks=False))
            line["extra_workloads"] = ex
    if ctx.rank == 0:
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
