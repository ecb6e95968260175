"""Direct parity of the hot-path kernels (not via the LM trajectory), through the C-ABI:

* arslam_get_normal_equations -- what accum_kernel / accum_e_pipe_kernel / accum_f_pipe_kernel
  write (W blocks, per-pose H | g | H_pose,f records, camera sums) against J^T J / J^T r formed in
  numpy from the ORACLE's Jet Jacobians (AutoDiffCostFunction<..,8,3,6,6>, ar_slam_util.cpp:722);
* one LM step (Schur elimination + reduced solve + back-substitution, kernels (3)/(4)) against a
  dense numpy solve of the damped, Jacobi-scaled normal equations Ceres' LevenbergMarquardtStrategy
  would form;
* arslam_set_constant against the oracle's SetParameterBlockConstant handling
  (ar_slam_util.cpp:965, :972 and the gauge fix of :697-700).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tri_unpack(rec21):
    H = np.zeros((6, 6))
    k = 0
    for i in range(6):
        for j in range(i, 6):
            H[i, j] = H[j, i] = rec21[k]
            k += 1
    return H


def oracle_pieces(oracle, m, cam, cap, tag):
    cost, res, jc, jp, ja = oracle.evaluate(m.cap_idx, m.tag_idx, m.obs, cam, cap, tag, num_threads=4)
    return cost, res, jc, jp, ja


@pytest.mark.parametrize("elim", [2, 1])
@pytest.mark.parametrize("pipe", [0, 3])
def test_normal_equations_match_oracle_jtj(gpu_solver_cls, oracle, elim, pipe):
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(700, 60, seed=5)          # 5.6 k blocks: several chunks per CTA and ragged segments
    s = gpu_solver_cls(options=ar_slam_b200.default_options(elimination=elim))
    s.set_tuning("accum_pipe", pipe)
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    side, bc, bt, W, Hc, Ht, cam4 = s.normal_equations()
    s.close()
    assert side == elim
    cost, res, jc, jp, ja = oracle_pieces(oracle, m, m.cam0, m.cap0, m.tag0)
    # blocks come back sorted by the E pose: match them to the oracle's by (capture, tag)
    key_o = m.cap_idx.astype(np.int64) * m.n_tag + m.tag_idx
    assert len(np.unique(key_o)) == len(key_o)
    order = np.argsort(key_o)
    pos = order[np.searchsorted(key_o[order], bc.astype(np.int64) * m.n_tag + bt)]
    assert np.array_equal(m.cap_idx[pos], bc) and np.array_equal(m.tag_idx[pos], bt)
    je, jf = (jp, ja) if elim == 2 else (ja, jp)
    W_o = np.einsum("bri,brj->bij", je[pos], jf[pos])
    assert np.abs(W - W_o).max() <= 1e-11 * np.abs(W_o).max()
    for H_g, idx, j, n in ((Hc, m.cap_idx, jp, m.n_cap), (Ht, m.tag_idx, ja, m.n_tag)):
        JtJ = np.zeros((n, 6, 6))
        np.add.at(JtJ, idx, np.einsum("bri,brj->bij", j, j))
        g = np.zeros((n, 6))
        np.add.at(g, idx, np.einsum("bri,br->bi", j, res))
        hf = np.zeros((n, 6))
        np.add.at(hf, idx, np.einsum("bri,br->bi", j, jc[:, :, 0]))
        H_u = np.array([tri_unpack(r[:21]) for r in H_g])
        assert np.abs(H_u - JtJ).max() <= 1e-11 * np.abs(JtJ).max()
        assert np.abs(H_g[:, 21:27] - g).max() <= 1e-11 * np.abs(g).max()
        assert np.abs(H_g[:, 27:33] - hf).max() <= 1e-11 * np.abs(hf).max()
    k = jc[:, :, 0]
    want = np.array([(k * k).sum(), (k * res).sum(), (res * res).sum()])
    assert np.allclose(cam4[:3], want, rtol=1e-12, atol=0)
    assert abs(0.5 * cam4[2] - cost) <= 1e-12 * cost


def numpy_lm_step(m, res, jc, jp, ja, radius, cam_const=False, cap_const=None, tag_const=None):
    """delta of one Levenberg-Marquardt step, dense: columns [f | captures | tags] (focal-only model)."""
    nb = len(m.cap_idx)
    ncol = 1 + 6 * m.n_cap + 6 * m.n_tag
    J = np.zeros((8 * nb, ncol))
    rows = np.arange(8 * nb).reshape(nb, 8)
    J[:, 0] = jc[:, :, 0].ravel()
    for b in range(nb):
        c0 = 1 + 6 * m.cap_idx[b]
        t0 = 1 + 6 * m.n_cap + 6 * m.tag_idx[b]
        J[rows[b][:, None], np.arange(c0, c0 + 6)[None, :]] = jp[b]
        J[rows[b][:, None], np.arange(t0, t0 + 6)[None, :]] = ja[b]
    live = np.ones(ncol, dtype=bool)
    if cam_const:
        live[0] = False
    if cap_const is not None:
        live[1:1 + 6 * m.n_cap] = ~np.repeat(np.asarray(cap_const, dtype=bool), 6)
    if tag_const is not None:
        live[1 + 6 * m.n_cap:] = ~np.repeat(np.asarray(tag_const, dtype=bool), 6)
    Jl = J[:, live]
    sigma = 1.0 / (1.0 + np.sqrt((Jl * Jl).sum(0)))
    Js = Jl * sigma
    H = Js.T @ Js
    D2 = np.clip(np.diag(H), 1e-6, 1e32) / radius
    y = np.linalg.solve(H + np.diag(D2), Js.T @ res.ravel())
    delta = np.zeros(ncol)
    delta[live] = -sigma * y
    return delta


@pytest.mark.parametrize("elim", [2, 1])
@pytest.mark.parametrize("lin", ["dense", "pcg"])
def test_single_lm_step_matches_numpy_normal_equations(gpu_solver_cls, oracle, elim, lin):
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(150, 30, seed=9)
    cost, res, jc, jp, ja = oracle_pieces(oracle, m, m.cam0, m.cap0, m.tag0)
    delta = numpy_lm_step(m, res, jc, jp, ja, 1e4)
    kw = dict(elimination=elim, max_num_iterations=1, function_tolerance=0.0, parameter_tolerance=0.0)
    if lin == "dense":
        kw["linear_solver"] = ar_slam_b200.LINSOLVE_DENSE
    else:
        kw.update(linear_solver=ar_slam_b200.LINSOLVE_PCG, pcg_tolerance=1e-13, pcg_max_iterations=5000)
    s = gpu_solver_cls(options=ar_slam_b200.default_options(**kw))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    summ, log = s.solve()
    cam, cap, tag = s.get_params()
    s.close()
    assert summ["iterations"] == 1 and summ["num_successful_steps"] == 2      # the step was taken
    got = np.concatenate([[cam[0] - m.cam0[0]], (cap - m.cap0).ravel(), (tag - m.tag0).ravel()])
    tol = 2e-9 if lin == "dense" else 2e-8
    assert np.abs(got - delta).max() <= tol * np.abs(delta).max(), np.abs(got - delta).max() / np.abs(delta).max()
    assert abs(log[1, 3] - np.linalg.norm(delta)) <= 1e-8 * np.linalg.norm(delta)


@pytest.mark.parametrize("elim", [2, 1])
def test_set_constant_matches_oracle(gpu_solver_cls, oracle, elim):
    """Camera + a few captures and tags held constant (ar_slam_util.cpp:965, :972; gauge fix :697-700)."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(200, 40, seed=21)
    rng = np.random.default_rng(3)
    cap_const = np.zeros(m.n_cap, dtype=np.uint8)
    tag_const = np.zeros(m.n_tag, dtype=np.uint8)
    cap_const[0] = 1                                     # the disabled gauge fix of the reference
    cap_const[rng.choice(m.n_cap, 7, replace=False)] = 1
    tag_const[rng.choice(m.n_tag, 5, replace=False)] = 1
    for cam_const in (False, True):
        # one step first, against the dense numpy solve of the reduced program
        cost, res, jc, jp, ja = oracle_pieces(oracle, m, m.cam0, m.cap0, m.tag0)
        delta = numpy_lm_step(m, res, jc, jp, ja, 1e4, cam_const, cap_const, tag_const)
        o1 = ar_slam_b200.default_options(elimination=elim, linear_solver=ar_slam_b200.LINSOLVE_DENSE, max_num_iterations=1,
                                          function_tolerance=0.0, parameter_tolerance=0.0)
        s = gpu_solver_cls(options=o1)
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_constant(camera=cam_const, cap=cap_const, tag=tag_const)
        s.set_params(m.cam0, m.cap0, m.tag0)
        s.solve()
        cam, cap, tag = s.get_params()
        got = np.concatenate([[cam[0] - m.cam0[0]], (cap - m.cap0).ravel(), (tag - m.tag0).ravel()])
        assert np.abs(got - delta).max() <= 2e-9 * np.abs(delta).max()
        # then the whole solve against the oracle's restated ceres::Solve with the same constant blocks
        s.set_options(ar_slam_b200.default_options(elimination=elim, linear_solver=ar_slam_b200.LINSOLVE_DENSE))
        s.set_params(m.cam0, m.cap0, m.tag0)
        sg, log_g = s.solve()
        cam_g, cap_g, tag_g = s.get_params()
        s.close()
        cam_o, cap_o, tag_o, so, log_o = oracle.solve(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                                      options=oracle.default_options(num_threads=4), cam_const=cam_const,
                                                      cap_const=cap_const, tag_const=tag_const)
        assert sg["iterations"] == so["iterations"] and sg["termination"] == so["termination"] and sg["reason"] == so["reason"]
        n = so["iterations"] + 1
        assert np.allclose(log_g[:n, 0], log_o[:n, 0], rtol=1e-7, atol=0)
        assert np.allclose(log_g[:n, 5], log_o[:n, 5], rtol=1e-5, atol=0)          # radius
        # gradient max norm over the live blocks only (row 0: the GPU's later rows lag by one evaluation)
        assert abs(log_g[0, 2] - log_o[0, 2]) <= 1e-9 * log_o[0, 2]
        # constant blocks keep their values bit for bit; the rest follows the oracle
        assert np.array_equal(cap_g[cap_const == 1], m.cap0[cap_const == 1])
        assert np.array_equal(tag_g[tag_const == 1], m.tag0[tag_const == 1])
        if cam_const:
            assert cam_g[0] == m.cam0[0]
        assert abs(cam_g[0] - cam_o[0]) <= 1e-6 * cam_o[0]
        assert np.abs(cap_g - cap_o).max() <= 1e-6 and np.abs(tag_g - tag_o).max() <= 1e-6


@pytest.mark.parametrize("n_tag", [10, 11, 21, 22, 42, 43, 64, 107, 150])
def test_dense_cholesky_sizes_match_numpy(gpu_solver_cls, oracle, n_tag):
    """The dense DMMA Cholesky at the sizes where its tiling changes shape: one 64-block (n_pad = 64), exactly one
    128-tile, a partial last 128-tile (n_pad = 192, 320, 448, 704), exactly one 256-wide outer panel and a second
    panel of a single block; reduced dimension n = 6 n_tag + 1.  One LM step against a dense numpy solve of the
    full normal equations; chained and stepwise back-substitution."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(4 * n_tag, n_tag, seed=100 + n_tag)
    cost, res, jc, jp, ja = oracle_pieces(oracle, m, m.cam0, m.cap0, m.tag0)
    delta = numpy_lm_step(m, res, jc, jp, ja, 1e4)
    for chain in (1, 0):
        s = gpu_solver_cls(options=ar_slam_b200.default_options(elimination=ar_slam_b200.ELIM_CAPTURES,
                                                                 linear_solver=ar_slam_b200.LINSOLVE_DENSE, max_num_iterations=1,
                                                                 function_tolerance=0.0, parameter_tolerance=0.0))
        s.set_tuning("chol_chain", chain)
        s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
        s.set_params(m.cam0, m.cap0, m.tag0)
        summ, log = s.solve()
        cam, cap, tag = s.get_params()
        s.close()
        assert summ["reduced_dim"] == 6 * n_tag + 1 and summ["iterations"] == 1 and summ["num_successful_steps"] == 2
        got = np.concatenate([[cam[0] - m.cam0[0]], (cap - m.cap0).ravel(), (tag - m.tag0).ravel()])
        assert np.abs(got - delta).max() <= 2e-9 * np.abs(delta).max(), (chain, np.abs(got - delta).max() / np.abs(delta).max())
