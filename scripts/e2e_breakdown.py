"""Where the time of one optimize() call through the C-ABI goes at config 3 (wall clock, device synchronised)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ar_slam_b200 as ar
from ar_slam_b200 import synth

m = synth.make_map(100000, 5000, 8, seed=0xA55A0003)
o = ar.default_options(max_num_iterations=5, function_tolerance=0.0, parameter_tolerance=0.0, gradient_tolerance=0.0,
                       linear_solver=ar.LINSOLVE_PCG)
s = ar.Solver(options=o)
pin = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0)]
def t(f, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
for rep in range(3):
    a, _ = t(lambda: s.set_problem(m.n_cap, m.n_tag, pin[0], pin[1], pin[2]))
    b, _ = t(lambda: s.set_params(pin[3], pin[4], pin[5]))
    c, _ = t(lambda: s.solve(log=False))
    d, _ = t(lambda: s.get_params())
    s.set_params(pin[3], pin[4], pin[5])
    e, _ = t(lambda: s.solve(log=False))
    print("rep %d: set_problem %.2f ms, set_params %.2f, first solve(5) %.2f, get_params %.2f | second solve(5) on the same problem %.2f" % (rep, a, b, c, d, e))
s.close()
