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
}
ITERS_PER_SOLVE = 5  # LM iterations per solve call; the solve restarts from the same initial state


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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


def bench_options(ar, iters, args):
    o = ar.default_options()
    o.max_num_iterations = iters
    # run exactly `iters` LM iterations: switch the convergence tests off
    o.function_tolerance = 0.0
    o.parameter_tolerance = 0.0
    o.gradient_tolerance = 0.0
    if args.linear_solver == "dense":
        o.linear_solver = ar.LINSOLVE_DENSE
        o.dense_max_dim = 1 << 20
    elif args.linear_solver == "pcg":
        o.linear_solver = ar.LINSOLVE_PCG
    o.pcg_tolerance = args.pcg_tolerance
    o.num_intrinsics = args.num_intrinsics
    if args.num_intrinsics == 3 and args.linear_solver == "auto":
        o.dense_max_dim = 1 << 20   # the radial model is solved with the dense Cholesky
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


def run_solves(s, m, total_iters, start=None):
    """Runs solves of ITERS_PER_SOLVE iterations from the same start until total_iters are done.
    start = (cam, cap, tag) arrays to restart from (pinned copies of the initial state)."""
    cam0, cap0, tag0 = start if start is not None else (m.cam0, m.cap0, m.tag0)
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


def cpu_baseline(args, steps, warmup, threads=None):
    """Times the oracle (restated reference, not Ceres) on a bounded sample of the workload."""
    from ar_slam_b200 import synth
    from oracle import pyoracle as po
    n_cap, n_tag, tpc, _ = WORKLOADS[args.workload]
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
    return {"value": nc * its / dt, "unit": "corners/s", "cores": threads, "kind": "port",
            "lm_iters_per_sec": its / dt, "reduced_dim": summ["reduced_dim"],
            "sample": "%d captures x %d tags (%d corners), same generator and density as %s, %d LM iterations, "
                      "restated reference (not Ceres): Jet autodiff + dense Schur, OpenMP x%d"
                      % (sc, st, nc, args.workload, its, threads)}, dt, its


def bench_localization(args, ar, synth, torch, dist, rank, world, local_rank):
    """BASELINE config 4: captures are sharded evenly, no communication at all."""
    n_loc, n_tag, tpc, cfg_id = WORKLOADS[args.workload]
    m = synth.make_localization_batch(n_loc, n_tag, tpc, seed=0xA55A0000 + cfg_id)
    lo, hi = n_loc * rank // world, n_loc * (rank + 1) // world
    b0, b1 = m.blk_offsets[lo], m.blk_offsets[hi]
    offs = (m.blk_offsets[lo:hi + 1] - b0).astype(np.int32)
    tag_idx, obs, seed = m.tag_idx[b0:b1], m.obs[b0:b1], m.seed_block[lo:hi]
    s = ar.Solver(device=local_rank)
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)

    def pinned(a):  # the batch arrives in pinned host memory; results go back into pinned arrays
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.ascontiguousarray(a[:0])).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep = [pinned(a) for a in (offs, tag_idx, obs, seed)]
    offs, tag_idx, obs, seed = [t.numpy() for t in keep]
    keep_out = [torch.empty((hi - lo, 6), dtype=torch.float64, pin_memory=True), torch.empty(hi - lo, dtype=torch.int32, pin_memory=True),
                torch.empty(hi - lo, dtype=torch.float64, pin_memory=True), torch.empty(hi - lo, dtype=torch.int32, pin_memory=True)]
    out = tuple(t.numpy() for t in keep_out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(1, args.warmup)):
        pose, its, cost, term = s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    s.set_profiling(True)
    kms = []
    for _ in range(args.steps):
        pose, its, cost, term = s.localize_batch(offs, tag_idx, obs, seed, m.cam_true, m.tag_true, out=out)
        kms.append([k for k in s.kernel_times() if k["name"] == "localize"][0]["total_ms"])
    barrier()
    dt = time.perf_counter() - t0
    ms_kernel = float(np.sum(kms))
    if dist is not None:
        t = torch.tensor([dt, ms_kernel], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, ms_kernel = float(t[0]), float(t[1])
    clk = clocks.stop() if rank == 0 else None
    if rank == 0:
        nc = 4 * len(m.tag_idx)
        peak, how = read_peaks()
        err = np.abs(pose - m.cap_true[lo:hi])
        cb = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import pyoracle as po
            ns = min(n_loc, 50000)
            threads = len(os.sched_getaffinity(0))
            t1 = time.perf_counter()
            po.localize_batch(m.blk_offsets[:ns + 1], m.tag_idx[:m.blk_offsets[ns]], m.obs[:m.blk_offsets[ns]],
                              m.seed_block[:ns], m.cam_true, m.tag_true, num_threads=threads)
            dtc = time.perf_counter() - t1
            cb = {"value": 4 * int(m.blk_offsets[ns]) / dtc, "unit": "corners/s", "cores": threads, "kind": "port",
                  "captures_per_sec": ns / dtc,
                  "sample": "first %d captures of the same batch, restated localizeOne (not Ceres), OpenMP x%d" % (ns, threads)}
        alg = 17.0 * 4 * len(tag_idx) + 116.0 * (hi - lo)
        line = {"metric": "observation_corners_per_sec", "value": nc * args.steps / (ms_kernel * 1e-3), "unit": "corners/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_kernel / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "captures_per_sec": n_loc * args.steps / (ms_kernel * 1e-3),
                "config": {"workload": args.workload, "captures": n_loc, "tags": n_tag, "tags_per_capture": tpc,
                           "corners": nc, "mean_lm_iterations": float(its.mean()),
                           "median_abs_pose_error": float(np.median(err)), "l2_policy": "inputs larger than L2"},
                "clocks": clk, "gpu_launches": 2 * args.steps,
                "e2e": {"value": nc * args.steps / dt, "unit": "corners/s", "captures_per_sec": n_loc * args.steps / dt,
                        "h2d_bytes_per_step": int(obs.nbytes + tag_idx.nbytes + offs.nbytes + seed.nbytes),
                        "d2h_bytes_per_step": int(pose.nbytes + its.nbytes + cost.nbytes + term.nbytes),
                        "what": "arslam_localize_batch from pinned host arrays (H2D, kernel, D2H), wall clock"},
                "roofline": {"bound": "hbm", "kernel": "localize", "achieved": alg * args.steps / (ms_kernel * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg * args.steps / (ms_kernel * 1e-3) / 1e9 / peak,
                             "traffic": None, "peak_source": how,
                             "note": "whole LM solve per capture in registers: FP64-latency bound, not HBM bound"},
                "cpu_baseline": cb}
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="kernel-variant switch passed to arslam_set_tuning (A/B measurements), e.g. accum_pipe=0")
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"],
                    help="weak: every GPU gets the workload's captures (the map, i.e. the tags, is shared); strong: the workload is split")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_cap, n_tag, tpc, cfg_id = WORKLOADS[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb, dt, its = cpu_baseline(args, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "observation_corners_per_sec", "value": cb["value"], "unit": "corners/s",
                "n_gpus": args.gpus, "steps": its, "warmup": args.warmup, "ms_per_step": 1e3 * dt / its,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "lm_iters_per_sec": cb["lm_iters_per_sec"],
                "config": {"workload": args.workload, "captures": n_cap, "tags": n_tag, "tags_per_capture": tpc,
                           "note": "CPU arm runs the bounded sample described in cpu_baseline.sample"},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "corners/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import ar_slam_b200 as ar
    from ar_slam_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload.startswith("loc_"):
        return bench_localization(args, ar, synth, torch, dist, rank, world, local_rank)
    if args.scaling == "weak":
        n_cap *= world
    m = synth.make_map(n_cap, n_tag, tpc, seed=0xA55A0000 + cfg_id,
                       distortion=(-0.05, 0.01) if args.num_intrinsics == 3 else (0.0, 0.0))
    cap_idx, tag_idx, obs = shard(m, rank, world)
    n_corner_total = 4 * len(m.cap_idx)

    opts = bench_options(ar, ITERS_PER_SOLVE, args)
    s = ar.Solver(device=local_rank, options=opts)
    for kv in args.tune:
        k, v = kv.split("=")
        s.set_tuning(k, int(v))
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(ar.Solver.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        s.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    s.set_problem(m.n_cap, m.n_tag, cap_idx, tag_idx, obs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # every solve restarts from the same initial state, re-sent from pinned host memory (the only
    # host -> device traffic inside the timed region: 4.9 MB per 5 LM iterations)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.ascontiguousarray(a[:0])).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    keep_start = [pinned(a) for a in (m.cam0, m.cap0, m.tag0)]
    start = tuple(t.numpy() for t in keep_start)
    # ---- warm-up
    if args.warmup > 0:
        run_solves(s, m, args.warmup, start)
    # ---- timed region: exactly --steps LM iterations, CUDA events on the solver's stream
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        summaries = run_solves(s, m, args.steps, start)
        ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    launches = int(sum(x["gpu_launches"] for x in summaries))
    pcg_its = int(sum(x["linear_solver_iterations"] for x in summaries))
    final_cost = summaries[-1]["final_cost"]

    # ---- end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # what ArSlamSolver::optimize does per call (ar_slam_util.cpp:1001-1018 with the blocks of
        # :720-727): the whole problem and the parameters go host -> device from pinned host memory,
        # ITERS_PER_SOLVE LM iterations run, the parameters come back.  One untimed call first.
        keep = [pinned(a) for a in (cap_idx, tag_idx, obs, m.cam0, m.cap0, m.tag0)]
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
        n_solves = max(1, -(-args.steps // ITERS_PER_SOLVE))
        barrier()
        t0 = time.perf_counter()
        e2e_iters = 0
        for _ in range(n_solves):
            e2e_iters += one_call()["iterations"]
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        param_bytes = 8 * (3 + 6 * m.n_cap + 6 * m.n_tag)
        h2d = n_solves * (obs.nbytes + cap_idx.nbytes + tag_idx.nbytes + param_bytes)
        d2h = n_solves * param_bytes + 8 * 40 * e2e_iters
        e2e = {"value": n_corner_total * e2e_iters / dt, "unit": "corners/s",
               "h2d_bytes_per_step": int(h2d / e2e_iters), "d2h_bytes_per_step": int(d2h / e2e_iters),
               "lm_iters_per_sec": e2e_iters / dt, "lm_iterations": e2e_iters,
               "what": "%d x (arslam_set_problem, set_params, solve of %d LM iterations, get_params) from pinned host "
                       "arrays, wall clock; a step is one LM iteration" % (n_solves, ITERS_PER_SOLVE)}

    # ---- per-kernel device times (separate profiled solve; events around every launch)
    roofline = None
    kernels = None
    if rank == 0:
        s.set_profiling(True)
    if True:
        run_solves(s, m, ITERS_PER_SOLVE, start)
    if rank == 0:
        kt = s.kernel_times()
        s.set_profiling(False)
        kernels = {k["name"]: {"ms_per_launch": k["total_ms"] / max(1, k["launches"]), "launches": k["launches"],
                               "algorithmic_bytes": k["algorithmic_bytes"]} for k in kt}
        peak, how = read_peaks()
        cand = [k for k in kt if k["name"] in ("accum_E", "accum_F") and k["launches"] > 0]
        if cand:
            top = max(cand, key=lambda k: k["total_ms"])
            sec = top["total_ms"] / top["launches"] * 1e-3
            ach = top["algorithmic_bytes"] / sec / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
            if os.path.exists(tp):   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture
                with open(tp) as f:
                    traffic = json.load(f).get(args.workload, {}).get(top["name"])
            roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": traffic, "peak_source": how,
                        "us_per_launch": sec * 1e6, "algorithmic_bytes": top["algorithmic_bytes"],
                        "note": "fused evaluation+accumulation, J never in HBM: 18 B/corner in + 72 B/corner W out + "
                                "264 B/pose; the kernel is also FP64-pipe bound (800 FP64 instructions per block)"}
            tot = sum(k["total_ms"] for k in kt)
            roofline["share_of_step"] = top["total_ms"] / tot if tot else None
            by_time = sorted(kt, key=lambda k: -k["total_ms"])[:3]
            roofline["largest_kernels"] = [{"kernel": k["name"], "share_of_step": k["total_ms"] / tot} for k in by_time]
        # dense path: the blocked Cholesky dominates; its roofline is the FP64 tensor (DMMA) pipe
        chol = [k for k in kt if k["name"] == "dense_cholesky" and k["launches"] > 0]
        if chol and cand and chol[0]["total_ms"] > max(k["total_ms"] for k in cand):
            n = summaries[0]["reduced_dim"]
            sec = chol[0]["total_ms"] / chol[0]["launches"] * 1e-3
            flops = n ** 3 / 3.0 + 2.0 * n * n
            dg_peak, dg_how = 35.5, "measured cuBLAS DGEMM on this pool's B200 (profiles/r1_fp64_peaks.json)"
            fp = os.path.join(ROOT, "profiles", "r1_fp64_peaks.json")
            if os.path.exists(fp):
                with open(fp) as f:
                    dg_peak = max(json.load(f).values())
            roofline = {"bound": "tensor", "kernel": "dense_cholesky", "achieved": flops / sec / 1e12, "peak": dg_peak,
                        "unit": "TFLOP/s", "frac": flops / sec / 1e12 / dg_peak, "traffic": None, "peak_source": dg_how,
                        "us_per_launch": sec * 1e6, "algorithmic_flops": flops,
                        "note": "FP64 DMMA (mma.sync.m8n8k4.f64) blocked Cholesky + solve, n^3/3 + 2 n^2 flop, n = %d; "
                                "the peak is FP64, not the bf16 figure of MEASURED_PEAKS.json" % n}

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb, _, _ = cpu_baseline(args, 2, 1)
        value = n_corner_total * args.steps / (ms * 1e-3)
        line = {"metric": "observation_corners_per_sec", "value": value, "unit": "corners/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "lm_iters_per_sec": args.steps / (ms * 1e-3),
                "config": {"workload": args.workload, "captures": m.n_cap, "tags": m.n_tag, "tags_per_capture": tpc,
                           "blocks": int(len(m.cap_idx)), "corners": int(n_corner_total),
                           "linear_solver": {1: "dense_cholesky_dmma", 2: "pcg"}[summaries[0]["linear_solver"]],
                           "intrinsics": "f, l1, l2 (radial model)" if args.num_intrinsics == 3 else "f (focal only, the reference's live model)",
                           "eliminated": {1: "tags", 2: "captures"}[summaries[0]["eliminated_side"]],
                           "reduced_dim": summaries[0]["reduced_dim"], "iters_per_solve": ITERS_PER_SOLVE,
                           "l2_policy": "inputs larger than L2 (W + observations > 126 MB) for ba_100k_5k; "
                                        "no explicit flush", "final_cost": final_cost,
                           "pcg_iterations": pcg_its},
                "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cb,
                "kernels": kernels}
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
