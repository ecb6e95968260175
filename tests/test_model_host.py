"""The closed-form Jacobian code of the CUDA kernels (ar_slam_b200/csrc/model.cuh), compiled
as host code by a test-only shim, against the oracle's Jet autodiff.  Runs without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libmodel_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out,
                           os.path.join(ROOT, "tests", "_shim", "model_host_shim.cpp")],
                          env={k: v for k, v in os.environ.items() if k not in ("CXX", "CC")})
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _eval(shim, cam, cap, tag, rect, ts=0.0635):
    res, jc, jp, ja = np.zeros(8), np.zeros((8, 3)), np.zeros((8, 6)), np.zeros((8, 6))
    shim.shim_eval_block(_p(cam), _p(cap), _p(tag), C.c_double(ts), _p(rect), _p(res), _p(jc), _p(jp), _p(ja))
    return res, jc, jp, ja


def test_closed_form_matches_jets(shim, oracle):
    rng = np.random.default_rng(0)
    worst_r = worst_j = worst_band = 0.0
    for t in range(4000):
        cam = np.array([rng.uniform(300, 3000), 0, 0])
        cap = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.8, 3)])
        tag = np.concatenate([rng.normal(0, 0.5, 3) + [0, 0, 2.0], rng.normal(0, 0.8, 3)])
        if t % 5 == 0:
            cap[3:] = 0
        if t % 7 == 0:
            tag[3:] = 0
        if t % 11 == 0:
            cap[3:] = rng.normal(0, 1, 3) * 10.0 ** rng.uniform(-9, -3)
        if t % 13 == 0:
            tag[3:] = rng.normal(0, 1, 3) * 10.0 ** rng.uniform(-9, -3)
        rect = rng.normal(0, 200, 8)
        _, r0, c0, p0, a0 = oracle.evaluate([0], [0], rect[None], cam, cap[None], tag[None])
        r, jc, jp, ja = _eval(shim, cam, cap, tag, rect)
        J0 = np.concatenate([c0[0], p0[0], a0[0]], 1)
        J = np.concatenate([jc, jp, ja], 1)
        e = (np.abs(J - J0).max(axis=1) / np.linalg.norm(J0, axis=1)).max()
        theta = np.linalg.norm(cap[3:])
        if 1.49e-8 < theta < 1e-6:
            worst_band = max(worst_band, e)
        else:
            worst_j = max(worst_j, e)
        worst_r = max(worst_r, (np.abs(r - r0[0]) / np.maximum(1, np.abs(r0[0]))).max())
    assert worst_r <= 1e-12
    assert worst_j <= 1e-9          # north_star tolerance
    assert worst_band <= 5e-9       # documented band, see tests/test_gpu_evaluate.py


def test_device_seed_matches_reference_heuristic(shim, oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        rect = rng.normal(0, 150, 8)
        tag = np.concatenate([rng.normal(0, 1, 3), rng.normal(0, 0.7, 3)])
        f = rng.uniform(400, 3000)
        out = np.zeros(6)
        shim.shim_seed_capture_pose(_p(rect), C.c_double(f), _p(tag), C.c_double(0.0635), _p(out))
        ref = oracle.init_capture_pose(rect, [f, 0, 0], tag)
        assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
        cap = np.concatenate([rng.normal(0, 1, 3), rng.normal(0, 0.7, 3)])
        shim.shim_seed_tag_pose(_p(rect), C.c_double(f), _p(cap), C.c_double(0.0635), _p(out))
        ref = oracle.init_tag_pose(rect, [f, 0, 0], cap)
        assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_register_cholesky(shim):
    rng = np.random.default_rng(4)
    for _ in range(50):
        A = rng.normal(size=(6, 6))
        H = A @ A.T + 0.1 * np.eye(6)
        b = rng.normal(size=6)
        x = np.zeros(6)
        ok = shim.shim_chol6_solve(_p(np.ascontiguousarray(H)), _p(b), _p(x))
        assert ok == 1 and np.allclose(x, np.linalg.solve(H, b), rtol=1e-9)
    H = np.eye(6)
    H[3, 3] = -1.0
    assert shim.shim_chol6_solve(_p(H), _p(b), _p(x)) == 0   # Eigen::LLT-style failure


def test_radial_model_matches_jets(shim, oracle):
    """The TODO model of ar_slam_util.cpp:164-171 (config 5): closed form vs the oracle's Jets."""
    rng = np.random.default_rng(8)
    worst = 0.0
    for t in range(1500):
        cam = np.array([rng.uniform(300, 3000), rng.normal(0, 0.08), rng.normal(0, 0.03)])
        cap = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 0.8, 3)])
        tag = np.concatenate([rng.normal(0, 0.5, 3) + [0, 0, 2.0], rng.normal(0, 0.8, 3)])
        if t % 5 == 0:
            cap[3:] = 0
        rect = rng.normal(0, 200, 8)
        _, r0, c0, p0, a0 = oracle.evaluate([0], [0], rect[None], cam, cap[None], tag[None], model=1)
        res, jc, jp, ja = np.zeros(8), np.zeros((8, 3)), np.zeros((8, 6)), np.zeros((8, 6))
        shim.shim_eval_block_dist(_p(cam), _p(cap), _p(tag), C.c_double(0.0635), _p(rect), _p(res), _p(jc), _p(jp),
                                  _p(ja))
        J0 = np.concatenate([c0[0], p0[0], a0[0]], 1)
        J = np.concatenate([jc, jp, ja], 1)
        worst = max(worst, (np.abs(J - J0).max(axis=1) / np.linalg.norm(J0, axis=1)).max(),
                    (np.abs(res - r0[0]) / np.maximum(1, np.abs(r0[0]))).max())
    assert worst <= 1e-9
