"""Pins the CPU oracle: 50-digit mpmath known-answer vectors, finite differences,
seed heuristics and the demo's converged numbers."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _kat():
    with open(os.path.join(GOLD, "kat_projection.json")) as f:
        return json.load(f)["cases"]


def test_oracle_matches_mpmath_known_answers(oracle):
    for c in _kat():
        cam, cap, tag = np.array(c["camera"]), np.array(c["capture"]), np.array(c["tag"])
        _, res, jc, jp, ja = oracle.evaluate([0], [0], np.zeros((1, 8)), cam, cap[None], tag[None],
                                             tag_size=c["tag_size"], model=c["model"])
        uv = np.array([float(x) for x in c["uv"]])
        J0 = np.array([[float(x) for x in row] for row in c["jacobian"]])
        J = np.concatenate([jc[0], jp[0], ja[0]], axis=1)
        assert np.abs(res[0] - uv).max() <= 1e-12 * max(1.0, np.abs(uv).max()), c["name"]
        # theta just above sqrt(DBL_EPSILON): (1 - cos theta) is quantised to ~1 ulp of 1.0 in double
        # Jets, an O(1e-9) error Ceres itself carries (documented, see tests/test_gpu_evaluate.py)
        tol = 5e-9 if c["name"] == "theta_just_above_eps" else 1e-12
        assert np.abs(J - J0).max() <= tol * np.abs(J0).max(), c["name"]


def test_appendix_d_values(oracle):
    uv = oracle.project_block([800, 0, 0], [0.1, -0.2, 1.5, 0.05, -0.1, 0.2], [0.3, 0.1, 0.2, -0.3, 0.2, 0.1])
    ref = [105.29456120018949947, -72.988777781197719918, 133.77292694025394604, -65.348499477179904203,
           126.5079739714648318, -38.657748339666287317, 97.787782098984175599, -46.473416430164075857]
    assert np.abs(uv - ref).max() < 1e-11


def test_oracle_jacobian_vs_central_differences(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        cam = np.array([rng.uniform(500, 1500), 0, 0])
        cap = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.5, 3)])
        tag = np.concatenate([rng.normal(0, 0.3, 3) + [0, 0, 2], rng.normal(0, 0.5, 3)])
        obs = np.zeros((1, 8))
        _, _, jc, jp, ja = oracle.evaluate([0], [0], obs, cam, cap[None], tag[None])
        x = np.concatenate([cam, cap, tag])
        J = np.concatenate([jc[0], jp[0], ja[0]], axis=1)

        def f(v):
            return oracle.evaluate([0], [0], obs, v[:3], v[None, 3:9], v[None, 9:15], jacobians=False)[1][0]
        for k in range(15):
            if k in (1, 2):
                assert np.all(J[:, k] == 0)   # l1, l2 inert
                continue
            h = 1e-6 * max(1.0, abs(x[k]))
            e = np.zeros(15)
            e[k] = h
            fd = (f(x + e) - f(x - e)) / (2 * h)
            assert np.abs(fd - J[:, k]).max() <= 1e-5 * max(1.0, np.abs(J[:, k]).max())


def test_small_angle_branch_is_first_order(oracle):
    # theta^2 <= DBL_EPSILON: R p = p + w x p exactly
    w = np.array([3e-9, -4e-9, 5e-9])
    p = np.array([0.3, -0.2, 1.1])
    assert np.array_equal(oracle.rotate_point(w, p), p + np.cross(w, p))
    w2 = w * 10
    assert not np.array_equal(oracle.rotate_point(w2, p), p + np.cross(w2, p))


def test_seed_heuristics(oracle):
    # a fronto-parallel tag at depth 2 m seen by an identity camera seeds back the same geometry
    cam = np.array([760.0, 0, 0])
    tag = np.array([0.1, -0.05, 2.0, 0, 0, 0.3])
    cap = np.zeros(6)
    rect = oracle.project_block(cam, cap, tag)
    seeded_cap = oracle.init_capture_pose(rect, cam, tag)
    assert np.abs(seeded_cap).max() < 2e-2
    seeded_tag = oracle.init_tag_pose(rect, cam, cap)
    assert np.abs(seeded_tag - tag).max() < 2e-2
    # composeAxisAngle of rotations about z adds the angles
    assert np.allclose(oracle.compose_axis_angle([0, 0, 0.2], [0, 0, 0.5]), [0, 0, 0.7])


def test_demo_map_build_and_localisation(oracle):
    """BASELINE config 1 with the restated CLI schedule: the sanity values of BASELINE.md section 6."""
    from oracle import schedule
    m = schedule.MapData()
    m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    assert len(m.cap_uid) == 3 and len(m.tag_id) == 6 and len(m.blk_cap) == 15
    sch = schedule.Scheduler(m)
    sch.solve()
    costs = [s["final_cost"] for s in m.solve_log]
    assert np.allclose(costs, [0.386, 5.359, 12.614], atol=2e-3)
    assert [s["reduced_dim"] for s in m.solve_log] == [9, 15, 21]      # tags eliminated, like Ceres' ordering
    assert abs(m.cam[0] - 758.7) < 0.1 and m.cam[1] == 0 and m.cam[2] == 0
    T = np.array(m.tag_pose)[:, :3]
    d = {t: np.linalg.norm(T[m.tag_map["aruco_4X4_50_%s" % t]] - T[m.tag_map["aruco_4X4_50_18"]])
         for t in ("19", "20", "21", "22", "23")}
    ref = {"19": 0.367, "20": 0.823, "21": 0.466, "22": 0.403, "23": 0.513}
    assert all(abs(d[k] - ref[k]) < 2e-3 for k in ref)
    cam = m.cam.copy()
    m.load_yaml(os.path.join(GOLD, "demo_loc_detections.yaml"))
    m.cam[:] = cam
    sch.localize_many(3)
    assert abs(m.solve_log[-1]["final_cost"] - 31.6) < 0.1 and m.solve_log[-1]["num_parameters"] == 6


def test_oracle_vs_scipy_least_squares(oracle):
    """Independent cross-check of the converged cost / focal length (different optimiser)."""
    from scipy.optimize import least_squares
    from oracle import schedule
    m = schedule.MapData()
    m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    schedule.Scheduler(m).solve()
    cap_idx, tag_idx, obs = np.array(m.blk_cap), np.array(m.blk_tag), np.array(m.blk_rect)
    x0 = np.concatenate([[m.cam[0]], np.array(m.cap_pose).ravel(), np.array(m.tag_pose).ravel()])

    def fun(x):
        cam = np.array([x[0], 0, 0])
        return oracle.evaluate(cap_idx, tag_idx, obs, cam, x[1:19].reshape(3, 6), x[19:].reshape(6, 6),
                               jacobians=False)[1].ravel()
    sol = least_squares(fun, x0, method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12)
    assert abs(sol.cost - m.solve_log[-1]["final_cost"]) <= 1e-4 * sol.cost
    assert abs(sol.x[0] - m.cam[0]) <= 2e-3 * sol.x[0]
    # run the oracle to tight tolerances from the same point: same minimum
    o = oracle.default_options(function_tolerance=1e-14, parameter_tolerance=1e-14, max_num_iterations=200)
    _, _, _, s, _ = oracle.solve(3, 6, cap_idx, tag_idx, obs, m.cam, np.array(m.cap_pose), np.array(m.tag_pose),
                                 options=o)
    assert abs(s["final_cost"] - sol.cost) <= 1e-8 * sol.cost


def test_elimination_choices_agree(oracle):
    from ar_slam_b200 import synth
    m = synth.make_map(60, 20, tags_per_capture=6, seed=9)
    out = []
    for elim in (0, 1, 2, 3):
        _, _, _, s, log = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                       options=oracle.default_options(elimination=elim))
        out.append((s, log))
    for s, log in out[1:]:
        assert s["iterations"] == out[0][0]["iterations"]
        assert np.allclose(log[:, 0], out[0][1][:, 0], rtol=1e-9)
