"""Parity at the BASELINE shapes themselves (configs 3, 4, 5), through the C-ABI.

The LM trajectory tests run at sizes the oracle's dense Schur solver finishes in seconds; here the
FULL shapes are gated with what stays cheap at that size:

* config 4 (1 M captures x 5 k tags): every capture against the oracle's restated localizeOne;
* config 3 (100 k x 5 k, 800 k blocks): evaluation cost and the block pieces of J^T J / J^T r against the
  oracle's Jet Jacobians, then the PCG path (tight) against the dense DMMA Cholesky path (n = 30 001)
  over several LM iterations -- two independent linear solvers on the same normal equations;
* config 5 (20 k x 2 k, radial model, all three intrinsics live): the first LM step against a dense
  scipy Cholesky solve of the Schur complement assembled in numpy from the oracle's Jacobians (model = 1).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config4_localization_1m_matches_oracle(gpu_solver_cls, oracle):
    from ar_slam_b200 import synth
    m = synth.make_localization_batch(1000000, 5000, 8, seed=0xA55A0004)
    s = gpu_solver_cls()
    pose, its, cost, term = s.localize_batch(m.blk_offsets, m.tag_idx, m.obs, m.seed_block, m.cam_true, m.tag_true)
    s.close()
    po, io, co, to = oracle.localize_batch(m.blk_offsets, m.tag_idx, m.obs, m.seed_block, m.cam_true, m.tag_true,
                                           num_threads=oracle.max_threads())
    same = its == io
    assert same.mean() >= 0.999, same.mean()
    assert np.array_equal(term[same], to[same])
    assert np.allclose(cost[same], co[same], rtol=1e-9)
    assert np.abs(pose[same] - po[same]).max() <= 1e-8
    # the few captures whose trajectory length differs (a convergence test decided by the last bits) still agree
    if (~same).any():
        assert np.abs(pose[~same] - po[~same]).max() <= 1e-5
    err = np.abs(pose - m.cap_true)
    assert np.median(err[:, :3]) < 5e-3 and np.median(err[:, 3:]) < 5e-3


def test_config3_accumulation_and_solvers_at_full_shape(gpu_solver_cls, oracle):
    import ar_slam_b200
    from ar_slam_b200 import synth
    from test_gpu_normal_equations import tri_unpack  # noqa: F401  (same record layout)
    m = synth.make_map(100000, 5000, 8, seed=0xA55A0003)
    assert len(m.cap_idx) == 800000
    s = gpu_solver_cls(options=ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_PCG, pcg_tolerance=1e-12,
                                                             pcg_max_iterations=4000, max_num_iterations=4,
                                                             function_tolerance=0.0, parameter_tolerance=0.0))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    # ---- evaluation + accumulation kernels vs the oracle's Jets at 3.2 M corners
    cost_g, res_g, _, _, _ = s.evaluate(jacobians=False)
    cost_o, res_o, jc, jp, ja = oracle.evaluate(m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0,
                                                num_threads=oracle.max_threads())
    assert abs(cost_g - cost_o) <= 1e-12 * cost_o
    assert np.abs(res_g - res_o).max() <= 1e-9 * np.abs(res_o).max()
    side, bc, bt, W, Hc, Ht, cam4 = s.normal_equations()
    assert side == ar_slam_b200.ELIM_CAPTURES
    key_o = m.cap_idx.astype(np.int64) * m.n_tag + m.tag_idx
    order = np.argsort(key_o)
    pos = order[np.searchsorted(key_o[order], bc.astype(np.int64) * m.n_tag + bt)]
    assert np.array_equal(m.cap_idx[pos], bc) and np.array_equal(m.tag_idx[pos], bt)
    W_o = np.einsum("bri,brj->bij", jp[pos], ja[pos])
    assert np.abs(W - W_o).max() <= 1e-11 * np.abs(W_o).max()
    iu = np.triu_indices(6)
    for H_g, idx, j, n in ((Hc, m.cap_idx, jp, m.n_cap), (Ht, m.tag_idx, ja, m.n_tag)):
        JtJ = np.zeros((n, 6, 6))
        np.add.at(JtJ, idx, np.einsum("bri,brj->bij", j, j))
        g = np.zeros((n, 6))
        np.add.at(g, idx, np.einsum("bri,br->bi", j, res_o))
        hf = np.zeros((n, 6))
        np.add.at(hf, idx, np.einsum("bri,br->bi", j, jc[:, :, 0]))
        assert np.abs(H_g[:, :21] - JtJ[:, iu[0], iu[1]]).max() <= 1e-11 * np.abs(JtJ).max()
        assert np.abs(H_g[:, 21:27] - g).max() <= 1e-11 * np.abs(g).max()
        assert np.abs(H_g[:, 27:33] - hf).max() <= 1e-11 * np.abs(hf).max()
    assert abs(0.5 * cam4[2] - cost_o) <= 1e-12 * cost_o
    del W, W_o, jp, ja, jc
    # ---- 4 LM iterations: block-sparse PCG (tight) vs the dense DMMA Cholesky (n = 30 001, 7.2 GB)
    sp, log_p = s.solve()
    cam_p, cap_p, tag_p = s.get_params()
    s.set_options(ar_slam_b200.default_options(linear_solver=ar_slam_b200.LINSOLVE_DENSE, dense_max_dim=1 << 20,
                                               max_num_iterations=4, function_tolerance=0.0, parameter_tolerance=0.0))
    s.set_params(m.cam0, m.cap0, m.tag0)
    sd, log_d = s.solve()
    cam_d, cap_d, tag_d = s.get_params()
    s.close()
    assert sp["linear_solver"] == ar_slam_b200.LINSOLVE_PCG and sd["linear_solver"] == ar_slam_b200.LINSOLVE_DENSE
    assert sd["reduced_dim"] == 30001 and sp["iterations"] == sd["iterations"] == 4
    assert np.allclose(log_p[:, 0], log_d[:, 0], rtol=1e-8, atol=0), (log_p[:, 0], log_d[:, 0])
    assert np.allclose(log_p[:, 5], log_d[:, 5], rtol=1e-6, atol=0)      # trust-region radius
    assert abs(cam_p[0] - cam_d[0]) <= 1e-7 * cam_d[0]
    assert np.abs(cap_p - cap_d).max() <= 1e-6 and np.abs(tag_p - tag_d).max() <= 1e-6
    assert log_d[-1, 0] < 1e-3 * log_d[0, 0]                             # and the cost really went down


def numpy_radial_first_step(oracle, m, radius=1e4):
    """First LM step of the radial model (f, l1, l2 live), captures eliminated: Jacobians from the oracle's Jets
    (model = 1), Schur complement assembled in numpy, reduced system solved by LAPACK.  Returns
    (cost, d_cam[3], d_cap[n_cap,6], d_tag[n_tag,6])."""
    import scipy.linalg
    nb, nc, nt = len(m.cap_idx), m.n_cap, m.n_tag
    cost_o, res, jc, jp, ja = oracle.evaluate(m.cap_idx, m.tag_idx, m.obs, m.cam0, m.cap0, m.tag0, model=1,
                                              num_threads=oracle.max_threads())
    r = res                                                    # [nb, 8]
    # Jacobi scaling sigma = 1 / (1 + column norm), Ceres' LevenbergMarquardtStrategy damping on the scaled columns
    n2c = np.zeros((nc, 6)); np.add.at(n2c, m.cap_idx, (jp * jp).sum(1))
    n2t = np.zeros((nt, 6)); np.add.at(n2t, m.tag_idx, (ja * ja).sum(1))
    n2k = (jc * jc).sum((0, 1))
    sc, st, sk = 1 / (1 + np.sqrt(n2c)), 1 / (1 + np.sqrt(n2t)), 1 / (1 + np.sqrt(n2k))
    Jc = jp * sc[m.cap_idx][:, None, :]
    Jt = ja * st[m.tag_idx][:, None, :]
    Jk = jc * sk[None, None, :]
    # E = captures (block diagonal), F = tags + 3 intrinsics (dense)
    Hcc = np.zeros((nc, 6, 6)); np.add.at(Hcc, m.cap_idx, np.einsum("bri,brj->bij", Jc, Jc))
    gc = np.zeros((nc, 6)); np.add.at(gc, m.cap_idx, np.einsum("bri,br->bi", Jc, r))
    d = np.clip(np.einsum("cii->ci", Hcc), 1e-6, 1e32) / radius
    Hcc[:, np.arange(6), np.arange(6)] += d
    nF = 6 * nt + 3
    Wct = np.einsum("bri,brj->bij", Jc, Jt)                    # per block: capture x its tag
    Wck = np.zeros((nc, 6, 3)); np.add.at(Wck, m.cap_idx, np.einsum("bri,brj->bij", Jc, Jk))
    HF = np.zeros((nF, nF))
    gF = np.zeros(nF)
    tt = np.einsum("bri,brj->bij", Jt, Jt)
    tk = np.einsum("bri,brj->bij", Jt, Jk)
    for b in range(nb):   # F-side diagonal blocks and borders
        t0 = 6 * m.tag_idx[b]
        HF[t0:t0 + 6, t0:t0 + 6] += tt[b]
        HF[t0:t0 + 6, 6 * nt:] += tk[b]
    HF[6 * nt:, :6 * nt] = HF[:6 * nt, 6 * nt:].T
    HF[6 * nt:, 6 * nt:] = np.einsum("bri,brj->ij", Jk, Jk)
    np.add.at(gF[:6 * nt].reshape(nt, 6), m.tag_idx, np.einsum("bri,br->bi", Jt, r))
    gF[6 * nt:] = np.einsum("bri,br->i", Jk, r)
    dF = np.clip(np.diag(HF), 1e-6, 1e32) / radius
    HF[np.arange(nF), np.arange(nF)] += dF
    # Schur complement over the captures
    Hinv = np.linalg.inv(Hcc)
    starts = np.concatenate([[0], np.cumsum(np.bincount(m.cap_idx, minlength=nc))])
    assert np.all(np.diff(m.cap_idx) >= 0)                     # blocks grouped by capture
    S, rhs = HF, gF
    for c in range(nc):
        bl = np.arange(starts[c], starts[c + 1])
        if len(bl) == 0:
            continue
        cols = np.concatenate([(6 * m.tag_idx[bl][:, None] + np.arange(6)[None, :]).ravel(), 6 * nt + np.arange(3)])
        Wc = np.concatenate([Wct[bl].transpose(1, 0, 2).reshape(6, -1), Wck[c]], axis=1)    # 6 x (6 k + 3)
        HW = Hinv[c] @ Wc
        S[np.ix_(cols, cols)] -= Wc.T @ HW
        rhs[cols] -= HW.T @ gc[c]
    yF = scipy.linalg.cho_solve(scipy.linalg.cho_factor(S, lower=True, overwrite_a=True, check_finite=False), rhs)
    # back-substitution
    Wy = np.zeros((nc, 6))
    np.add.at(Wy, m.cap_idx, np.einsum("bij,bj->bi", Wct, yF[:6 * nt].reshape(nt, 6)[m.tag_idx]))
    Wy += np.einsum("cij,j->ci", Wck, yF[6 * nt:])
    yc = np.einsum("cij,cj->ci", Hinv, gc - Wy)
    return cost_o, -sk * yF[6 * nt:], -sc * yc, -st * yF[:6 * nt].reshape(nt, 6)


@pytest.mark.parametrize("solver", ["dense", "pcg"])
def test_config5_radial_first_step_at_full_shape(gpu_solver_cls, oracle, solver):
    """Config 5 at its full shape, first LM step of the radial model: the dense DMMA Cholesky and the block-sparse
    system with three border columns solved tightly by PCG, both against a scipy solve of the same Schur complement."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_map(20000, 2000, 8, seed=0xA55A0005, distortion=(-0.05, 0.01))
    cost_o, d_cam, d_cap, d_tag = numpy_radial_first_step(oracle, m)
    ls = ar_slam_b200.LINSOLVE_DENSE if solver == "dense" else ar_slam_b200.LINSOLVE_PCG
    s = gpu_solver_cls(options=ar_slam_b200.default_options(num_intrinsics=3, max_num_iterations=1, dense_max_dim=1 << 20,
                                                             linear_solver=ls, pcg_tolerance=1e-13, pcg_max_iterations=5000,
                                                             function_tolerance=0.0, parameter_tolerance=0.0))
    s.set_problem(m.n_cap, m.n_tag, m.cap_idx, m.tag_idx, m.obs)
    s.set_params(m.cam0, m.cap0, m.tag0)
    cost_g, res_g, _, _, _ = s.evaluate(jacobians=False)
    assert abs(cost_g - cost_o) <= 1e-12 * cost_o
    summ, log = s.solve()
    cam, cap, tag = s.get_params()
    s.close()
    assert summ["linear_solver"] == ls and summ["reduced_dim"] == 12003
    assert summ["iterations"] == 1 and summ["num_successful_steps"] == 2
    scale = max(np.abs(d_cap).max(), np.abs(d_tag).max())
    assert np.abs((cap - m.cap0) - d_cap).max() <= 1e-7 * scale
    assert np.abs((tag - m.tag0) - d_tag).max() <= 1e-7 * scale
    assert np.abs((cam - m.cam0) - d_cam).max() <= 1e-7 * np.abs(d_cam).max()
