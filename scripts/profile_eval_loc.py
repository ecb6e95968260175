"""Runs kernel (1) eval_jacobian_kernel at config 3 (3.2 M corners, Jacobians materialised) and kernel (5) localize_kernel at
config 4 (1 M captures, one chunk) a few times -- the command ncu wraps for profiles/r2_ncu_eval_localize.*"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ar_slam_b200 as ar  # noqa: E402
from ar_slam_b200 import synth  # noqa: E402

m = synth.make_map(100000, 5000, 8, seed=0xA55A0003)
s = ar.Solver()
s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
s.set_params(m.cam0, m.cap0, m.tag0)
s.set_profiling(True)
for _ in range(3):
    s.evaluate(jacobians=True)
    kt = {k["name"]: k for k in s.kernel_times()}
print("eval_jacobian us", 1e3 * kt["eval_jacobian"]["total_ms"], "bytes", kt["eval_jacobian"]["algorithmic_bytes"],
      "GB/s", kt["eval_jacobian"]["algorithmic_bytes"] / (kt["eval_jacobian"]["total_ms"] * 1e-3) / 1e9)
for _ in range(2):
    s.evaluate(jacobians=False)
    kt = {k["name"]: k for k in s.kernel_times()}
print("eval residuals only us", 1e3 * kt["eval_jacobian"]["total_ms"], "GB/s", kt["eval_jacobian"]["algorithmic_bytes"] / (kt["eval_jacobian"]["total_ms"] * 1e-3) / 1e9)
n = int(os.environ.get("LOC_N", "1000000"))
lm = synth.make_localization_batch(n, 5000, 8, seed=0xA55A0004)
s.set_tuning("loc_chunk", n)
for _ in range(3):
    t = time.perf_counter()
    s.localize_batch(lm.blk_offsets, lm.tag_idx, lm.obs, lm.seed_block, lm.cam_true, lm.tag_true)
    kt = {k["name"]: k for k in s.kernel_times()}
    print("localize us", 1e3 * kt["localize"]["total_ms"], "wall ms", 1e3 * (time.perf_counter() - t))
s.close()
