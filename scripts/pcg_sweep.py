"""Experiment (not part of the product): LM convergence and time versus PCG tolerance."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import ar_slam_b200 as ar
from ar_slam_b200 import synth

m = synth.make_map(100000, 5000, seed=0xA55A0003)
for tol in (1e-1, 3e-2, 1e-2, 1e-3, 1e-4, 1e-6, 1e-8, 1e-10):
    o = ar.default_options(pcg_tolerance=tol, pcg_max_iterations=2000)
    s = ar.Solver(options=o)
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    summ, log = s.solve()
    s.set_params(m.cam0, m.cap0, m.tag0)
    t = time.perf_counter(); summ, log = s.solve(); dt = time.perf_counter() - t
    cam, cap, tag = s.get_params()
    print(json.dumps({"tol": tol, "lm_iters": summ["iterations"], "final_cost": summ["final_cost"], "reason": summ["reason_name"],
                      "pcg_iters": summ["linear_solver_iterations"], "solve_ms": dt * 1e3, "focal": cam[0],
                      "costs": [float("%.6g" % c) for c in log[:, 0]]}))
    s.close()
