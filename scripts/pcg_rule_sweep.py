"""Time to converge (function_tolerance 1e-6, the reference's stopping rule) under different PCG stopping rules:
residual based ||S y - b|| <= r ||b|| and Ceres' quadratic-model rule i (Q_i - Q_{i-1}) / Q_i < q."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ar_slam_b200 as ar  # noqa: E402
from ar_slam_b200 import synth  # noqa: E402

out = {}
for name, (nc, nt, seed) in (("ba_100k_5k", (100000, 5000, 0xA55A0003)), ("ba_20k_2k", (20000, 2000, 0xA55A0005)), ("ba_1k_200", (1000, 200, 0xA55A0002))):
    m = synth.make_map(nc, nt, 8, seed=seed)
    s = ar.Solver(options=ar.default_options(linear_solver=ar.LINSOLVE_PCG))
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    rows = []
    for r_tol, q_tol in ((0.1, 0.0), (0.0, 0.1), (0.0, 0.3), (0.0, 0.03), (0.0, 0.01), (0.03, 0.0), (0.3, 0.0)):
        o = ar.default_options(linear_solver=ar.LINSOLVE_PCG, pcg_tolerance=r_tol, pcg_q_tolerance=q_tol, pcg_max_iterations=2000)
        s.set_options(o)
        best = None
        for rep in range(2):
            s.set_params(m.cam0, m.cap0, m.tag0)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                ev0.record(stream)
                summ, _ = s.solve(log=False)
                ev1.record(stream)
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            best = ms if best is None else min(best, ms)
        rows.append({"r_tol": r_tol, "q_tol": q_tol, "ms": best, "lm_iterations": summ["iterations"], "pcg_iterations": int(summ["linear_solver_iterations"]),
                     "final_cost": summ["final_cost"], "reason": summ["reason_name"]})
        print(name, rows[-1], flush=True)
    out[name] = rows
    s.close()
print(json.dumps(out))
