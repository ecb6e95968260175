"""Wall-clock breakdown of the host-facing calls (set_problem / set_params / solve / get_params)
on BASELINE config 3, with pageable and with pinned host arrays.  Run on a GPU box."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ar_slam_b200 as ar  # noqa: E402
from ar_slam_b200 import synth  # noqa: E402


def pinned_like(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
    t.numpy()[...] = a
    return t


def main():
    n_cap, n_tag = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    m = synth.make_map(n_cap, n_tag, 8, seed=0xA55A0003)
    opts = ar.default_options(max_num_iterations=5, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    s = ar.Solver(options=opts)
    keep = []
    for label, conv in (("pageable", lambda a: a), ("pinned", lambda a: keep.append(pinned_like(a)) or keep[-1].numpy())):
        ci, ti, ob = conv(m.cap_idx), conv(m.tag_idx), conv(m.obs)
        cam0, cap0, tag0 = conv(m.cam0), conv(m.cap0), conv(m.tag0)
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s.set_problem(m.n_cap, m.n_tag, ci, ti, ob)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            s.set_params(cam0, cap0, tag0)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            summ, _ = s.solve(log=False)
            t3 = time.perf_counter()
            s.get_params()
            t4 = time.perf_counter()
            s.set_params(cam0, cap0, tag0)
            summ2, _ = s.solve(log=False)
            t5 = time.perf_counter()
            print("%-8s rep %d: set_problem %.2f ms  set_params %.2f  solve#1 %.2f (%d it)  get_params %.2f  set_params+solve#2 %.2f"
                  % (label, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), summ["iterations"], 1e3 * (t4 - t3), 1e3 * (t5 - t4)))
    s.close()


if __name__ == "__main__":
    main()
