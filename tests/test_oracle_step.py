"""Pins the oracle's restated LevenbergMarquardtStrategy + SchurEliminator step (oracle/ar_oracle.cpp) against
an independent dense numpy solve of the damped, Jacobi-scaled normal equations (SURVEY Appendix B items
2-3), with and without constant parameter blocks.  CPU only."""
import numpy as np

from test_gpu_normal_equations import numpy_lm_step


def test_oracle_first_step_equals_dense_numpy_solve(oracle):
    from ar_slam_b200 import synth
    m = synth.make_map(150, 30, seed=9)
    cost, res, jc, jp, ja = oracle.evaluate(m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0)
    o = oracle.default_options(max_num_iterations=1, function_tolerance=0.0, parameter_tolerance=0.0)
    cc = np.zeros(m.n_cap, dtype=np.uint8)
    tc = np.zeros(m.n_tag, dtype=np.uint8)
    cc[:5] = 1
    tc[3] = 1
    for cam_const, cap_const, tag_const in ((False, None, None), (True, cc, tc), (False, cc, None)):
        for elim in (0, 1, 2):
            o.elimination = elim
            want = numpy_lm_step(m, res, jc, jp, ja, 1e4, cam_const, cap_const, tag_const)
            cam, cap, tag, s, log = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                                 options=o, cam_const=cam_const, cap_const=cap_const, tag_const=tag_const)
            got = np.concatenate([[cam[0] - m.cam0[0]], (cap - m.cap0).ravel(), (tag - m.tag0).ravel()])
            assert s["iterations"] == 1
            assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()


def test_oracle_radial_first_step_equals_numpy_schur(oracle):
    """The numpy Schur reference used by the full-shape config-5 GPU test, pinned to the oracle's own solve
    (model = 1: f, l1, l2 live) on a small map."""
    from ar_slam_b200 import synth
    from test_gpu_full_shapes import numpy_radial_first_step
    m = synth.make_map(120, 40, seed=4, distortion=(-0.05, 0.01))
    cost, d_cam, d_cap, d_tag = numpy_radial_first_step(oracle, m)
    o = oracle.default_options(max_num_iterations=1, function_tolerance=0.0, parameter_tolerance=0.0, elimination=2)
    cam, cap, tag, s, log = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                         options=o, model=1)
    assert s["iterations"] == 1 and abs(s["initial_cost"] - cost) <= 1e-12 * cost
    scale = max(np.abs(d_cap).max(), np.abs(d_tag).max())
    assert np.abs((cap - m.cap0) - d_cap).max() <= 1e-9 * scale
    assert np.abs((tag - m.tag0) - d_tag).max() <= 1e-9 * scale
    assert np.abs((cam - m.cam0) - d_cam).max() <= 1e-9 * np.abs(d_cam).max()
