"""Parity of the LM solve (kernels 2-4 + control loop) with the oracle's
restated ceres::Solve (ar_slam_util.cpp:1001-1018), through the C-ABI.

The gauge is free, so poses are compared after a rigid
alignment; cost, focal length and the iteration trajectory are gauge
invariant and compared directly.
"""
import os

import numpy as np
import pytest

from util import align_rigid

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_lm(Solver, **opt_kw):
    import ar_slam_b200

    def lm(m, blocks, options=None, cam_const=False, tags_const=False):
        assert not cam_const and not tags_const
        s = Solver(options=ar_slam_b200.default_options(**opt_kw))
        s.set_problem(len(m.cap_uid), len(m.tag_id), [m.blk_cap[b] for b in blocks],
                      [m.blk_tag[b] for b in blocks], np.array([m.blk_rect[b] for b in blocks]))
        s.set_params(m.cam, np.array(m.cap_pose), np.array(m.tag_pose))
        summ, log = s.solve()
        cam, cap, tag = s.get_params()
        s.close()
        summ["log"] = log
        return cam, cap, tag, summ
    return lm


@pytest.mark.parametrize("elim", [0, 1, 2])
def test_demo_map_build_matches_oracle(gpu_solver_cls, oracle, elim):
    """BASELINE config 1: img1-3, the CLI's BFS schedule (ar_slam_util.cpp:744-866)."""
    from oracle import schedule
    mo, mg = schedule.MapData(), schedule.MapData()
    for m in (mo, mg):
        m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    schedule.Scheduler(mo).solve()
    schedule.Scheduler(mg, lm=gpu_lm(gpu_solver_cls, elimination=elim)).solve()
    assert len(mo.solve_log) == len(mg.solve_log) == 3
    for so, sg in zip(mo.solve_log, mg.solve_log):
        assert sg["termination"] == so["termination"] == 0
        assert abs(sg["iterations"] - so["iterations"]) <= 1
        # the first solve starts from identical numbers; later ones start from the previous
        # solve's end point, which was cut off at function_tolerance 1e-6 on a gauge-free problem
        tol = 1e-9 if so is mo.solve_log[0] else 1e-6
        assert abs(sg["initial_cost"] - so["initial_cost"]) <= tol * so["initial_cost"]
        assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-5 * so["final_cost"]
    assert abs(mg.cam[0] - mo.cam[0]) <= 1e-4 * mo.cam[0]
    assert abs(mg.cam[0] - 758.66) < 0.05 and mg.cam[1] == 0.0 and mg.cam[2] == 0.0
    # poses up to the free gauge: align tag centres, then compare
    To, Tg = np.array(mo.tag_pose)[:, :3], np.array(mg.tag_pose)[:, :3]
    R, t = align_rigid(Tg, To)
    assert np.abs((Tg @ R.T + t) - To).max() < 2e-3


def test_synthetic_1k_200_trajectory(gpu_solver_cls, oracle):
    """BASELINE config 2: same LM trajectory as the oracle, iteration by iteration."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(1000, 200)
    cam_o, cap_o, tag_o, so, log_o = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0,
                                                  m.tag0, options=oracle.default_options(num_threads=4))
    s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_DENSE))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    sg, log_g = s.solve()
    cam_g, cap_g, tag_g = s.get_params()
    s.close()
    assert sg["linear_solver"] == ar_slam_b200.LINSOLVE_DENSE and sg["eliminated_side"] == ar_slam_b200.ELIM_CAPTURES
    assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"]
    assert sg["reason"] == so["reason"]
    n = so["iterations"] + 1
    # cost, |step|, rho, radius per iteration
    assert np.allclose(log_g[:n, 0], log_o[:n, 0], rtol=1e-7, atol=0)
    assert np.allclose(log_g[1:n, 3], log_o[1:n, 3], rtol=1e-5)
    assert np.allclose(log_g[1:n - 1, 4], log_o[1:n - 1, 4], rtol=1e-5)
    assert np.allclose(log_g[:n, 5], log_o[:n, 5], rtol=1e-5)
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"]
    assert abs(cam_g[0] - cam_o[0]) <= 1e-7 * cam_o[0]
    used = np.unique(m.tag_idx)
    assert np.abs(tag_g[used] - tag_o[used]).max() < 1e-6
    assert np.abs(cap_g - cap_o).max() < 1e-6


def test_elimination_sides_agree(gpu_solver_cls):
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(300, 80, seed=7)
    out = []
    for elim in (ar_slam_b200.ELIM_TAGS, ar_slam_b200.ELIM_CAPTURES):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(elimination=elim, linear_solver=ar_slam_b200.LINSOLVE_DENSE))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        out.append((summ, log, s.get_params()))
        s.close()
    (sa, la, pa), (sb, lb, pb) = out
    assert sa["eliminated_side"] == 1 and sb["eliminated_side"] == 2
    assert sa["iterations"] == sb["iterations"]
    assert np.allclose(la[:, 0], lb[:, 0], rtol=1e-8)
    assert abs(pa[0][0] - pb[0][0]) < 1e-6 * pb[0][0]


def test_solve_is_deterministic(gpu_solver_cls):
    """No global atomics in the accumulation: evaluation-side results are bit-reproducible."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(500, 100, seed=11)
    costs = []
    for _ in range(2):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(max_num_iterations=0))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, _ = s.solve()
        costs.append(summ["initial_cost"])
        s.close()
    assert costs[0] == costs[1]


def test_pcg_matches_dense(gpu_solver_cls, oracle):
    """The sparse reduced system + PCG (used where the reduced system is large) must walk the
    same LM trajectory as the dense Cholesky path and the oracle when solved tightly."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(1000, 200)
    res = {}
    for name, ls in (("dense", ar_slam_b200.LINSOLVE_DENSE), ("pcg", ar_slam_b200.LINSOLVE_PCG)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ls, pcg_tolerance=1e-12,
                                                                 pcg_max_iterations=2000))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        res[name] = (summ, log, s.get_params())
        s.close()
    (sd, ld, pd), (sp, lp, pp) = res["dense"], res["pcg"]
    assert sp["linear_solver"] == ar_slam_b200.LINSOLVE_PCG and sp["linear_solver_iterations"] > 0
    assert sp["iterations"] == sd["iterations"] and sp["reason"] == sd["reason"]
    assert np.allclose(lp[:, 0], ld[:, 0], rtol=1e-9)
    assert np.allclose(lp[1:, 3], ld[1:, 3], rtol=1e-6)
    assert abs(pp[0][0] - pd[0][0]) <= 1e-8 * pd[0][0]
    assert np.abs(pp[1] - pd[1]).max() < 1e-7 and np.abs(pp[2] - pd[2]).max() < 1e-7
    _, _, _, so, log_o = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                      options=oracle.default_options(num_threads=4))
    assert sp["iterations"] == so["iterations"]
    assert abs(sp["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"]


def test_pipelined_pcg_matches_dense(gpu_solver_cls):
    """pcg_tolerance >= 1e-6 selects the one-barrier (pipelined) recurrence: at 1e-6 the LM
    trajectory still follows the dense Cholesky path closely."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(4000, 600, seed=33)
    res = {}
    for name, ls in (("dense", ar_slam_b200.LINSOLVE_DENSE), ("pcg", ar_slam_b200.LINSOLVE_PCG)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ls, pcg_tolerance=1e-6, pcg_max_iterations=2000))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        res[name] = (summ, log, s.get_params())
        s.close()
    (sd, ld, pd), (sp, lp, pp) = res["dense"], res["pcg"]
    assert sp["linear_solver_iterations"] > 0
    assert sp["iterations"] == sd["iterations"] and sp["reason"] == sd["reason"]
    assert np.allclose(lp[:, 0], ld[:, 0], rtol=1e-6)
    assert abs(sp["final_cost"] - sd["final_cost"]) <= 1e-7 * sd["final_cost"]
    assert abs(pp[0][0] - pd[0][0]) <= 1e-6 * pd[0][0]


def test_pcg_default_tolerance_converges_like_dense(gpu_solver_cls):
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(3000, 400, seed=21)
    out = {}
    for name, ls in (("dense", ar_slam_b200.LINSOLVE_DENSE), ("pcg", ar_slam_b200.LINSOLVE_PCG)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ls))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        out[name], _ = s.solve()
        s.close()
    assert out["pcg"]["termination"] == 0 and out["dense"]["termination"] == 0
    # default pcg_tolerance 0.1 (inexact Newton, Ceres' eta): a couple more LM iterations, same minimum
    assert 0 <= out["pcg"]["iterations"] - out["dense"]["iterations"] <= 3
    assert abs(out["pcg"]["final_cost"] - out["dense"]["final_cost"]) <= 1e-5 * out["dense"]["final_cost"]


@pytest.mark.parametrize("elim,ls", [(2, 1), (1, 1), (2, 2)])
def test_ragged_visibility_matches_oracle(gpu_solver_cls, oracle, elim, ls):
    """Captures seeing 1..12 tags, tags seen by 1..~150 captures: pose segments of every length,
    straddling warps and CTAs in both sorted copies; unused tags and captures; a tag seen twice by
    one capture.  Same LM trajectory as the oracle with either side eliminated and with PCG."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(700, 90, tags_per_capture=12, seed=17)
    rng = np.random.default_rng(5)
    keep = rng.random(len(m.cap_idx)) < rng.uniform(0.15, 1.0, m.n_cap)[m.cap_idx]
    keep[np.unique(m.cap_idx, return_index=True)[1]] = True          # at least one block per capture ...
    keep[m.cap_idx == 5] = False                                     # ... except capture 5: no blocks at all
    ci, ti, ob = m.cap_idx[keep], m.tag_idx[keep], m.obs[keep]
    ci = np.concatenate([ci, [7]]).astype(np.int32)                  # capture 7 sees its first tag twice
    ti = np.concatenate([ti, [ti[np.nonzero(ci == 7)[0][0]]]]).astype(np.int32)
    ob = np.concatenate([ob, ob[np.nonzero(ci == 7)[0][0]][None] + 0.5])
    perm = rng.permutation(len(ci))                                  # blocks arrive in arbitrary order
    ci, ti, ob = ci[perm], ti[perm], ob[perm]
    opts = oracle.default_options(num_threads=4, elimination=elim)
    cam_o, cap_o, tag_o, so, log_o = oracle.solve(m.n_cap, m.n_tag, ci, ti, ob, m.cam0, m.cap0, m.tag0, options=opts)
    s = gpu_solver_cls(options=ar_slam_b200.default_options(elimination=elim, linear_solver=ls, pcg_tolerance=1e-13,
                                                             pcg_max_iterations=5000))
    s.set_problem(m.n_cap, m.n_tag, ci, ti, ob)
    s.set_params(m.cam0, m.cap0, m.tag0)
    sg, log_g = s.solve()
    cam_g, cap_g, tag_g = s.get_params()
    cost_g = s.evaluate(jacobians=False)[0] if False else None
    s.close()
    assert sg["eliminated_side"] == elim and sg["linear_solver"] == ls
    assert sg["iterations"] == so["iterations"] and sg["reason"] == so["reason"]
    n = so["iterations"] + 1
    assert np.allclose(log_g[:n, 0], log_o[:n, 0], rtol=1e-7)
    assert np.allclose(log_g[:n, 5], log_o[:n, 5], rtol=1e-4)
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"]
    assert abs(cam_g[0] - cam_o[0]) <= 1e-6 * cam_o[0]
    assert np.array_equal(cap_g[5], m.cap0[5])                       # the capture without blocks is untouched
    used_t = np.unique(ti)
    unused_t = np.setdiff1d(np.arange(m.n_tag), used_t)
    assert np.array_equal(tag_g[unused_t], m.tag0[unused_t])
    assert np.abs(tag_g[used_t] - tag_o[used_t]).max() < 1e-5
    used_c = np.unique(ci)
    assert np.abs(cap_g[used_c] - cap_o[used_c]).max() < 1e-5


def test_auto_picks_dense_only_where_the_reduced_matrix_is_dense(gpu_solver_cls):
    import ar_slam_b200
    from ar_slam_b200 import synth
    picks = {}
    for name, (nc, nt, tpc) in {"sparse": (1000, 200, 8), "dense": (400, 24, 8)}.items():
        m = synth.make_map(nc, nt, tpc, seed=23)
        s = gpu_solver_cls()
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        picks[name], _ = s.solve()
        s.close()
        assert picks[name]["termination"] == 0
    assert picks["sparse"]["linear_solver"] == ar_slam_b200.LINSOLVE_PCG
    assert picks["dense"]["linear_solver"] == ar_slam_b200.LINSOLVE_DENSE


def _radial_map(oracle, n_cap, n_tag, seed, l1=0.08, l2=-0.03):
    """Synthetic map re-rendered through the radial model (camera = f, l1, l2)."""
    from ar_slam_b200 import synth
    m = synth.make_map(n_cap, n_tag, seed=seed)
    m.cam_true = np.array([m.cam_true[0], l1, l2])
    _, uv, _, _, _ = oracle.evaluate(m.cap_idx, m.tag_idx, np.zeros_like(m.obs), m.cam_true, m.cap_true, m.tag_true,
                                     model=1, jacobians=False)
    rng = np.random.default_rng(seed)
    m.obs = (uv + rng.normal(0, 0.3, uv.shape)).astype(np.float32).astype(np.float64)
    return m


@pytest.mark.parametrize("elim", [1, 2])
def test_radial_model_trajectory(gpu_solver_cls, oracle, elim):
    """num_intrinsics = 3 (f, l1, l2 all free; ar_slam_util.cpp:164-171 TODO, BASELINE config 5)."""
    import ar_slam_b200
    m = _radial_map(oracle, 400, 100, seed=21)
    cam_o, cap_o, tag_o, so, log_o = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0,
                                                  m.tag0, options=oracle.default_options(num_threads=4), model=1)
    s = gpu_solver_cls(options=ar_slam_b200.default_options(num_intrinsics=3, elimination=elim,
                                                             linear_solver=ar_slam_b200.LINSOLVE_DENSE))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    sg, log_g = s.solve()
    cam_g, cap_g, tag_g = s.get_params()
    s.close()
    assert sg["linear_solver"] == ar_slam_b200.LINSOLVE_DENSE and sg["eliminated_side"] == elim
    assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"] == 0
    n = so["iterations"] + 1
    assert np.allclose(log_g[:n, 0], log_o[:n, 0], rtol=1e-7, atol=0)
    assert np.allclose(log_g[1:n, 3], log_o[1:n, 3], rtol=1e-5)
    assert np.allclose(log_g[:n, 5], log_o[:n, 5], rtol=1e-5)
    assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"]
    assert np.abs(cam_g - cam_o).max() <= 1e-6 * max(1.0, np.abs(cam_o).max())
    # and the distortion is actually recovered
    assert abs(cam_g[1] - 0.08) < 0.02 and abs(cam_g[0] - m.cam_true[0]) < 2.0
    # the focal-only model cannot explain these observations
    s = gpu_solver_cls(options=ar_slam_b200.default_options(elimination=elim))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    s1, _ = s.solve()
    s.close()
    assert s1["final_cost"] > 1.5 * sg["final_cost"]


@pytest.mark.parametrize("elim", [1, 2])
def test_radial_model_pcg_matches_dense(gpu_solver_cls, oracle, elim):
    """The radial model on the sparse path: three border columns (f, l1, l2) and a 3 x 3 intrinsics block in the
    block-sparse reduced system (SparseTarget::add_border_x, pcg_finalize_kernel, pcg_kernel<3>).  Solved tightly it
    walks the dense Cholesky trajectory."""
    import ar_slam_b200
    m = _radial_map(oracle, 1200, 300, seed=29)
    res = {}
    for name, ls, tol in (("dense", ar_slam_b200.LINSOLVE_DENSE, 0.1), ("pcg", ar_slam_b200.LINSOLVE_PCG, 1e-12),
                          ("pipelined", ar_slam_b200.LINSOLVE_PCG, 1e-6)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(num_intrinsics=3, elimination=elim, linear_solver=ls,
                                                                 pcg_tolerance=tol, pcg_max_iterations=3000))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        res[name] = (summ, log, s.get_params())
        s.close()
    (sd, ld, pd), (sp, lp, pp) = res["dense"], res["pcg"]
    assert sp["linear_solver"] == ar_slam_b200.LINSOLVE_PCG and sp["linear_solver_iterations"] > 0
    assert sd["termination"] == sp["termination"] == 0
    assert sp["iterations"] == sd["iterations"]
    assert np.allclose(lp[:, 0], ld[:, 0], rtol=1e-8)
    assert np.allclose(lp[1:, 3], ld[1:, 3], rtol=1e-5)
    assert np.abs(pp[0] - pd[0]).max() <= 1e-7 * max(1.0, np.abs(pd[0]).max())
    assert np.abs(pp[1] - pd[1]).max() < 1e-6 and np.abs(pp[2] - pd[2]).max() < 1e-6
    assert abs(pp[0][1] - 0.08) < 0.02
    # the one-barrier kernel (pcg_pipe_kernel<3>, chosen for pcg_tolerance >= 1e-6) reaches the same minimum
    sq, lq, pq = res["pipelined"]
    assert sq["termination"] == 0 and abs(sq["iterations"] - sd["iterations"]) <= 1
    assert abs(sq["final_cost"] - sd["final_cost"]) <= 1e-6 * sd["final_cost"]
    assert np.abs(pq[0] - pd[0]).max() <= 1e-4 * max(1.0, np.abs(pd[0]).max())


def test_parameter_round_trip_and_continued_solve(gpu_solver_cls):
    """set_params / get_params move data straight between the caller's arrays and HBM: what was set
    comes back bit-exactly, a solve updates it in place (as Ceres updates the caller's parameter
    blocks, ar_slam_util.hpp:72,208,237), and a second solve continues from the result."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(500, 120, seed=5)
    s = gpu_solver_cls(options=ar_slam_b200.default_options(max_num_iterations=3))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    with pytest.raises(ar_slam_b200.ArslamError):
        s.get_params()                      # nothing set yet for this problem
    s.set_params(m.cam0, m.cap0, m.tag0)
    cam, cap, tag = s.get_params()
    assert np.array_equal(cam, m.cam0) and np.array_equal(cap, m.cap0) and np.array_equal(tag, m.tag0)
    s1, _ = s.solve()
    cam1, cap1, tag1 = s.get_params()
    assert s1["iterations"] == 3 and s1["final_cost"] < s1["initial_cost"]
    assert not np.array_equal(cap1, m.cap0)
    c1, _, _, _, _ = s.evaluate(jacobians=False)
    assert abs(c1 - s1["final_cost"]) <= 1e-12 * c1          # evaluate sees the solved state
    s2, _ = s.solve()                                         # continues, does not restart
    assert abs(s2["initial_cost"] - s1["final_cost"]) <= 1e-12 * s1["final_cost"]
    assert s2["final_cost"] <= s1["final_cost"]
    # caller-owned output arrays
    out_cap, out_tag = np.zeros((m.n_cap, 6)), np.zeros((m.n_tag, 6))
    cam2, cap2, tag2 = s.get_params(out=(out_cap, out_tag))
    assert cap2 is out_cap and tag2 is out_tag and np.abs(out_cap).max() > 0
    # a new problem invalidates the parameters
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    with pytest.raises(ar_slam_b200.ArslamError):
        s.solve()
    s.close()


def test_append_blocks_equals_full_problem(gpu_solver_cls):
    """arslam_append_blocks (the AddResidualBlock loop of the incremental schedules,
    ar_slam_util.cpp:723, :832) leaves the device in the state arslam_set_problem builds from
    all blocks: evaluation is bit-identical, the solve agrees."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(600, 150, seed=17)
    cut1 = int(np.searchsorted(m.cap_idx, 250))    # blocks of captures 0..249
    cut2 = int(np.searchsorted(m.cap_idx, 251))    # one more capture, as solveIncremental adds them
    opts = ar_slam_b200.default_options(max_num_iterations=4, linear_solver=ar_slam_b200.LINSOLVE_DENSE)
    full = gpu_solver_cls(options=opts)
    full.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    full.set_params(m.cam0, m.cap0, m.tag0)
    inc = gpu_solver_cls(options=opts)
    inc.append_blocks(250, m.n_tag, m.cap_idx[:cut1], m.tag_idx[:cut1], m.obs[:cut1])   # no problem yet: == set_problem
    inc.append_blocks(251, m.n_tag, m.cap_idx[cut1:cut2], m.tag_idx[cut1:cut2], m.obs[cut1:cut2])
    with pytest.raises(ar_slam_b200.ArslamError):
        inc.solve()                                   # parameters must be set again
    with pytest.raises(ar_slam_b200.ArslamError):
        inc.append_blocks(100, m.n_tag, m.cap_idx[:1], m.tag_idx[:1], m.obs[:1])         # counts cannot shrink
    with pytest.raises(ar_slam_b200.ArslamError):
        inc.append_blocks(251, m.n_tag, [251], [0], np.zeros((1, 8)))                    # index out of range
    inc.append_blocks(m.n_cap, m.n_tag, m.cap_idx[cut2:], m.tag_idx[cut2:], m.obs[cut2:])
    assert inc.n_blk == full.n_blk == len(m.cap_idx)
    inc.set_params(m.cam0, m.cap0, m.tag0)
    cf, rf, jcf, jpf, jaf = full.evaluate()
    ci, ri, jci, jpi, jai = inc.evaluate()
    assert cf == ci and np.array_equal(rf, ri) and np.array_equal(jpf, jpi) and np.array_equal(jaf, jai)
    sf, lf = full.solve()
    si, li = inc.solve()
    assert si["iterations"] == sf["iterations"] and si["initial_cost"] == sf["initial_cost"]
    assert abs(si["final_cost"] - sf["final_cost"]) <= 1e-10 * sf["final_cost"]
    pf, pi = full.get_params(), inc.get_params()
    assert np.abs(pf[1] - pi[1]).max() < 1e-9 and np.abs(pf[2] - pi[2]).max() < 1e-9
    full.close()
    inc.close()


def test_pipelined_accumulation_is_bit_identical(gpu_solver_cls):
    """accum_e/f_pipe_kernel (cross-block cp.async pipeline) must write the same records and W as the
    thread-per-block accum_kernel: same summation order, so the dense LM trajectory is bit-identical.
    Ragged visibility, more chunks than CTAs (grid-stride loop) and a partial last chunk."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(6000, 300, seed=11)
    logs = []
    for pipe in (0, 1):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_DENSE,
                                                                 max_num_iterations=4))
        s.set_tuning("accum_pipe", pipe)
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        cam, cap, tag = s.get_params()
        s.close()
        logs.append((log, cam, cap, tag))
    # the camera/cost sums are grouped per CTA, and the persistent kernel has fewer CTAs: costs agree
    # to rounding, everything per-pose is bit-identical on the first linearisation
    assert logs[0][0].shape == logs[1][0].shape
    assert np.allclose(logs[0][0][:, 0], logs[1][0][:, 0], rtol=1e-11, atol=0), (logs[0][0][:, 0] - logs[1][0][:, 0])
    assert np.allclose(logs[0][2], logs[1][2], rtol=0, atol=1e-9), np.abs(logs[0][2] - logs[1][2]).max()
    assert np.allclose(logs[0][3], logs[1][3], rtol=0, atol=1e-9), np.abs(logs[0][3] - logs[1][3]).max()


def test_locality_order_and_local_elimination_equal_default(gpu_solver_cls):
    """The capture-sorted copy stored in locality order (captures sorted by their smallest tag, segments with
    explicit ends) and the elimination kernel that pre-reduces products inside the CTA (schur_local.cuh) are
    parked experiments (measured slower), but they stay correct: same LM trajectory as the default path."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(3000, 150, seed=13)
    logs = []
    for locality, local in ((0, 0), (1, 0), (1, 1), (0, 1)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_PCG, pcg_tolerance=1e-12,
                                                                 pcg_max_iterations=3000, max_num_iterations=4))
        s.set_tuning("locality", locality)
        s.set_tuning("schur_local", local)
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        cam, cap, tag = s.get_params()
        s.close()
        logs.append((log, cam, cap, tag))
    for other in logs[1:]:
        assert np.allclose(other[0][:, 0], logs[0][0][:, 0], rtol=1e-10, atol=0)
        assert np.abs(other[2] - logs[0][2]).max() <= 1e-8 and np.abs(other[3] - logs[0][3]).max() <= 1e-8


def test_pcg_quadratic_model_termination(gpu_solver_cls):
    """pcg_q_tolerance: Ceres' ConjugateGradientsSolver rule (stop when i (Q_i - Q_{i-1}) / Q_i < q, Q(x) = x'Sx - 2b'x; what
    its trust-region strategies use for inexact steps with q = eta = 0.1).  Same minimum as the tightly solved
    problem, far fewer PCG iterations."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(4000, 400, seed=17)
    res = {}
    for name, kw in (("tight", dict(pcg_tolerance=1e-10)), ("q", dict(pcg_tolerance=0.0, pcg_q_tolerance=0.1))):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_PCG, pcg_max_iterations=3000, **kw))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        res[name], _ = s.solve()
        s.close()
    assert res["q"]["termination"] == 0 and res["tight"]["termination"] == 0
    assert abs(res["q"]["final_cost"] - res["tight"]["final_cost"]) <= 1e-5 * res["tight"]["final_cost"]
    assert res["q"]["linear_solver_iterations"] < 0.5 * res["tight"]["linear_solver_iterations"]
    # the rule actually stops the solves: nowhere near the iteration cap
    assert res["q"]["linear_solver_iterations"] < 0.5 * 3000 * res["q"]["iterations"]


def test_dense_cholesky_variants_agree(gpu_solver_cls):
    """The asynchronous trailing update (cp.async stream across tiles, C updated through TMA bulk reductions) does
    the same arithmetic as the synchronous round-1 kernel: every C element receives one contribution per launch,
    so the factor -- and with it the whole LM trajectory -- is bit-identical.  The one-launch chained
    back-substitution sums in a different (fixed) order than the stepwise one: agreement to rounding.
    n = 6 * 420 + 1 = 2521: ten outer panels, a partial last 64-block (the tiling's edge sizes:
    test_gpu_normal_equations.py::test_dense_cholesky_sizes_match_numpy)."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(2500, 420, seed=23)
    runs = {}
    for name, tune in (("default", {}), ("sync", {"chol_big": 0}), ("wide_panel", {"chol_nb": 512}), ("stepwise", {"chol_chain": 0})):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_DENSE, max_num_iterations=4))
        for k, v in tune.items():
            s.set_tuning(k, v)
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        runs[name] = (summ, log, s.get_params())
        s.close()
    ref = runs["default"]
    assert ref[0]["termination"] in (0, 1)
    for name in ("sync", "wide_panel"):
        # the reduced system itself is summed with FP64 reductions in arrival order (schur_eliminate), so two runs
        # agree to rounding, not to the bit; the factorisation variants add nothing to that
        assert np.allclose(runs[name][1][:, 0], ref[1][:, 0], rtol=1e-11, atol=0), name
        assert np.allclose(runs[name][2][2], ref[2][2], rtol=0, atol=1e-9), name
    assert np.allclose(runs["stepwise"][1][:, 0], ref[1][:, 0], rtol=1e-10, atol=0)
    assert np.allclose(runs["stepwise"][2][2], ref[2][2], rtol=0, atol=1e-8)


@pytest.mark.parametrize("shape", [(60, 12), (500, 90)])
@pytest.mark.parametrize("cam_const", [False, True])
def test_radial_pcg_small_and_constant_camera(gpu_solver_cls, oracle, shape, cam_const):
    """Radial model on the sparse path at sizes where most CTAs of the persistent kernels own no row, and with the
    three intrinsics held constant (sigma = 0: their 3 x 3 block is the damping alone, the border columns vanish)."""
    import ar_slam_b200
    m = _radial_map(oracle, shape[0], shape[1], seed=41)
    res = {}
    for name, ls, tol in (("dense", ar_slam_b200.LINSOLVE_DENSE, 0.1), ("pcg", ar_slam_b200.LINSOLVE_PCG, 1e-12),
                          ("pipelined", ar_slam_b200.LINSOLVE_PCG, 1e-6)):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(num_intrinsics=3, linear_solver=ls, pcg_tolerance=tol,
                                                                 pcg_max_iterations=3000, max_num_iterations=6))
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_constant(camera=cam_const)
        s.set_params(m.cam_true if cam_const else m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        res[name] = (summ, log, s.get_params())
        s.close()
    sd, ld, pd = res["dense"]
    # (an inexact first step from a far start moves the intermediate costs by ~1e-4; the end point is the same)
    for name, rtol, rtol_end in (("pcg", 1e-8, 1e-8), ("pipelined", 1e-3, 1e-6)):
        sp, lp, pp = res[name]
        assert sp["linear_solver"] == ar_slam_b200.LINSOLVE_PCG
        assert len(lp) == len(ld) and np.allclose(lp[:, 0], ld[:, 0], rtol=rtol), name
        assert abs(lp[-1, 0] - ld[-1, 0]) <= rtol_end * ld[-1, 0], name
        if cam_const:
            assert np.array_equal(pp[0], m.cam_true)
