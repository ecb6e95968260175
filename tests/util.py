"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def random_blocks(rng, n_cap, n_tag, n_blk, special=True):
    """Random but geometrically sane blocks: tags ~2 m in front of the cameras."""
    cam = np.array([rng.uniform(500, 1500), 0.0, 0.0])
    cap = np.concatenate([rng.normal(0, 0.3, (n_cap, 3)), rng.normal(0, 0.4, (n_cap, 3))], axis=1)
    tag = np.concatenate([rng.normal(0, 0.4, (n_tag, 3)) + [0, 0, 2.5], rng.normal(0, 0.6, (n_tag, 3))], axis=1)
    if special:
        cap[0, 3:] = 0.0                                  # exact identity (first capture of every map)
        tag[0, 3:] = 0.0
        if n_cap > 3:
            cap[1, 3:] = [3e-9, -4e-9, 5e-9]              # theta^2 below DBL_EPSILON
            cap[2, 3:] = [1e-5, -2e-5, 1.5e-5]
            cap[3, 3:] = [0.4, 3.0, -0.3]                 # close to pi
        if n_tag > 2:
            tag[1, 3:] = [2e-9, 1e-9, -3e-9]
            tag[2, 3:] = [2.0, -1.5, 0.8]
    cap_idx = rng.integers(0, n_cap, n_blk).astype(np.int32)
    tag_idx = rng.integers(0, n_tag, n_blk).astype(np.int32)
    obs = rng.normal(0, 250, (n_blk, 8))
    return cam, cap, tag, cap_idx, tag_idx, obs


def jac_rel_err(J, J0):
    """max over rows of |dJ|_inf / ||J0_row||_2 for Jacobians stacked [n_blk, 8, cols]."""
    num = np.abs(J - J0).max(axis=2)
    den = np.maximum(np.linalg.norm(J0, axis=2), 1e-300)
    return float((num / den).max())


def align_rigid(P, Q):
    """Rigid transform (R, t) minimising |R P + t - Q| for point sets [n,3]."""
    mp, mq = P.mean(0), Q.mean(0)
    H = (P - mp).T @ (Q - mq)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1, 1, d]) @ U.T
    return R, mq - R @ mp
