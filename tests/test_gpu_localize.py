"""Parity of kernel (5), batched localisation, with the oracle's restated
localizeOne (ar_slam_util.cpp:903-979), through the C-ABI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_demo_img4(gpu_solver_cls, oracle):
    """BASELINE config 1b: ar_loc of img4 against the map built from img1-3."""
    from oracle import schedule
    m = schedule.MapData()
    m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    sch = schedule.Scheduler(m)
    sch.solve()
    cam = m.cam.copy()
    m.load_yaml(os.path.join(GOLD, "demo_loc_detections.yaml"))
    m.cam[:] = cam   # loadYaml of the detections overwrote the camera (ar_slam_util.cpp:357-367)
    blocks = m.cap_blocks[3]
    tag_idx = [m.blk_tag[b] for b in blocks]
    obs = np.array([m.blk_rect[b] for b in blocks])
    s = gpu_solver_cls()
    pose, its, cost, term = s.localize_batch([0, len(blocks)], tag_idx, obs, [0], m.cam, np.array(m.tag_pose))
    s.close()
    sch.localize_many(3)
    so = m.solve_log[-1]
    assert its[0] == so["iterations"] and term[0] == 0
    assert abs(cost[0] - so["final_cost"]) <= 1e-9 * so["final_cost"]
    assert abs(cost[0] - 31.6121425) < 1e-5
    assert np.abs(pose[0] - m.cap_pose[3]).max() < 1e-9


def test_batch_matches_oracle(gpu_solver_cls, oracle):
    from ar_slam_b200 import synth
    m = synth.make_localization_batch(20000, 500, seed=5)
    # ragged input: drop a few blocks, mark some captures as not localisable
    seed = m.seed_block.copy()
    seed[5::13] = 3
    seed[::97] = -1
    s = gpu_solver_cls()
    pose, its, cost, term = s.localize_batch(m.blk_offsets, m.tag_idx, m.obs, seed, m.cam_true, m.tag_true)
    s.close()
    po, io, co, to = oracle.localize_batch(m.blk_offsets, m.tag_idx, m.obs, seed, m.cam_true, m.tag_true,
                                           num_threads=4)
    assert np.array_equal(its < 0, io < 0) and np.all(its[::97] == -1)
    ok = io >= 0
    same = its[ok] == io[ok]
    assert same.mean() > 0.999            # identical LM trajectory length
    assert np.array_equal(term[ok], to[ok])
    assert np.allclose(cost[ok][same], co[ok][same], rtol=1e-9)
    assert np.abs(pose[ok][same] - po[ok][same]).max() < 1e-8
    # and the answer is right: close to the ground-truth pose
    err = np.abs(pose[ok] - m.cap_true[ok])
    assert np.median(err[:, :3]) < 5e-3 and np.median(err[:, 3:]) < 5e-3


def test_batch_radial_model(gpu_solver_cls, oracle):
    """Localisation against a map whose camera carries radial distortion (num_intrinsics = 3)."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_localization_batch(4000, 300, seed=9)
    cam = np.array([m.cam_true[0], 0.08, -0.03])
    n_blk = len(m.tag_idx)
    cap_idx = np.repeat(np.arange(len(m.blk_offsets) - 1), np.diff(m.blk_offsets)).astype(np.int32)
    _, uv, _, _, _ = oracle.evaluate(cap_idx, m.tag_idx, np.zeros((n_blk, 8)), cam, m.cap_true, m.tag_true,
                                     model=1, jacobians=False)
    obs = (uv + np.random.default_rng(9).normal(0, 0.3, uv.shape)).astype(np.float32).astype(np.float64)
    s = gpu_solver_cls(options=ar_slam_b200.default_options(num_intrinsics=3))
    pose, its, cost, term = s.localize_batch(m.blk_offsets, m.tag_idx, obs, m.seed_block, cam, m.tag_true)
    s.close()
    po, io, co, to = oracle.localize_batch(m.blk_offsets, m.tag_idx, obs, m.seed_block, cam, m.tag_true,
                                           model=1, num_threads=4)
    same = its == io
    assert same.mean() > 0.995
    assert np.array_equal(term[same], to[same])
    assert np.allclose(cost[same], co[same], rtol=1e-9)
    assert np.abs(pose[same] - po[same]).max() < 1e-8
    err = np.abs(pose - m.cap_true)
    assert np.median(err[:, :3]) < 5e-3 and np.median(err[:, 3:]) < 5e-3


def test_chunked_pipeline_equals_one_chunk_and_rejects_bad_indices(gpu_solver_cls):
    """arslam_localize_batch uploads / solves / downloads in chunks on three streams: any chunk size gives
    the same bits; indices are validated on the device and a bad batch returns ARSLAM_ERR_INVALID."""
    import ar_slam_b200
    from ar_slam_b200 import synth
    m = synth.make_localization_batch(5000, 200, seed=6)
    outs = []
    for chunk in (0, 700, 4999, 1):
        s = gpu_solver_cls()
        s.set_tuning("loc_chunk", chunk if chunk != 1 else 333)
        outs.append(s.localize_batch(m.blk_offsets, m.tag_idx, m.obs, m.seed_block, m.cam_true, m.tag_true))
        s.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)
    s = gpu_solver_cls()
    bad_tag = m.tag_idx.copy()
    bad_tag[1234] = 200                      # == n_tag
    with pytest.raises(ar_slam_b200.ArslamError) as e:
        s.localize_batch(m.blk_offsets, bad_tag, m.obs, m.seed_block, m.cam_true, m.tag_true)
    assert e.value.code == -1
    bad_seed = m.seed_block.copy()
    bad_seed[77] = 50                        # beyond the capture's blocks
    with pytest.raises(ar_slam_b200.ArslamError):
        s.localize_batch(m.blk_offsets, m.tag_idx, m.obs, bad_seed, m.cam_true, m.tag_true)
    bad_off = m.blk_offsets.copy()
    bad_off[100] = bad_off[101] + 1          # not monotone
    with pytest.raises(ar_slam_b200.ArslamError):
        s.localize_batch(bad_off, m.tag_idx, m.obs, m.seed_block, m.cam_true, m.tag_true)
    # the handle is still usable
    pose, its, cost, term = s.localize_batch(m.blk_offsets, m.tag_idx, m.obs, m.seed_block, m.cam_true, m.tag_true)
    assert np.array_equal(pose, outs[0][0])
    s.close()
