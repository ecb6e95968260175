"""The C++ host facade (ar_slam_b200/host): the drop-in for the reference's ArSlamSolver class
and its CLIs.  CPU part: data store, addDetections semantics, map.yaml format.  GPU part: the
CLIs end to end on the demo fixtures, against the oracle running the same schedules."""
import os
import subprocess

import numpy as np
import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "ar_slam_b200", "lib")


@pytest.fixture(scope="module")
def host_tools():
    from ar_slam_b200 import build
    build.build()
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "ar_slam_b200", "host"), "-s"], env=env)
    return LIB


def test_yaml_round_trip_is_byte_stable(host_tools, tmp_path):
    src = os.path.join(GOLD, "demo_map_detections.yaml")
    out1 = subprocess.check_output([os.path.join(host_tools, "host_selftest"), "roundtrip", src], text=True)
    p = tmp_path / "a.yaml"
    p.write_text(out1)
    out2 = subprocess.check_output([os.path.join(host_tools, "host_selftest"), "roundtrip", str(p)], text=True)
    assert out1 == out2
    doc = yaml.safe_load(out1)
    assert list(doc.keys()) == ["blocks", "captures", "arucos", "camera"]
    assert len(doc["blocks"]) == 15 and list(doc["blocks"][0].keys()) == ["capture", "aruco", "aruco_rect"]
    assert doc["camera"] == {"params": [3000, 0, 0], "width": 1020, "height": 768}
    ref = yaml.safe_load(open(src))
    assert doc == ref
    # doubles are written with 17 significant digits (yaml-cpp's max_digits10)
    p.write_text(out1.replace("params: [3000, 0, 0]", "params: [0.1, 758.66221424000003, 1e-300]"))
    out3 = subprocess.check_output([os.path.join(host_tools, "host_selftest"), "roundtrip", str(p)], text=True)
    assert "params: [0.10000000000000001, 758.66221424000003, 1e-300]" in out3


def test_add_detections_semantics(host_tools):
    r = subprocess.run([os.path.join(host_tools, "host_selftest"), "detections"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    doc = yaml.safe_load(r.stdout)
    assert list(doc["captures"].keys()) == ["img1"] and doc["captures"]["img1"]["img_fn"] == "/tmp/img1.jpg"
    assert doc["blocks"][0]["aruco_rect"][0] == float(np.float32(0.1))   # Point32 widened to double


def test_cli_usage_errors(host_tools):
    assert subprocess.run([os.path.join(host_tools, "ar_slam_cli")], capture_output=True).returncode == 1
    assert subprocess.run([os.path.join(host_tools, "ar_loc")], capture_output=True).returncode == 1
    r = subprocess.run([os.path.join(host_tools, "ar_slam_cli"), "img1.jpg"], capture_output=True, text=True)
    assert r.returncode == 2 and "image ingest" in r.stderr


@pytest.mark.gpu
def test_cli_map_build_and_ar_loc_match_oracle(host_tools, tmp_path, oracle):
    """BASELINE config 1 through the drop-in CLIs: ar_slam_cli -> map.yaml, ar_loc -> localize.yaml."""
    from oracle import schedule
    r = subprocess.run([os.path.join(host_tools, "ar_slam_cli"), os.path.join(GOLD, "demo_map_detections.yaml")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = schedule.MapData()
    got.load_yaml(str(tmp_path / "map.yaml"))
    ref = schedule.MapData()
    ref.load_yaml(os.path.join(GOLD, "demo_map_detections.yaml"))
    schedule.Scheduler(ref).solve()
    assert abs(got.cam[0] - ref.cam[0]) <= 1e-4 * ref.cam[0] and got.cam[1] == 0 and got.cam[2] == 0
    cost_got = oracle.evaluate(got.blk_cap, got.blk_tag, np.array(got.blk_rect), got.cam, np.array(got.cap_pose),
                               np.array(got.tag_pose), jacobians=False)[0]
    assert abs(cost_got - ref.solve_log[-1]["final_cost"]) <= 1e-5 * cost_got
    assert [got.cap_uid, got.tag_id] == [ref.cap_uid, ref.tag_id]
    # ar_loc: the detections yaml must carry the map's camera (loadYaml overwrites it, :357-367)
    loc = yaml.safe_load(open(os.path.join(GOLD, "demo_loc_detections.yaml")))
    loc["camera"]["params"] = [float(v) for v in got.cam]
    (tmp_path / "loc.yaml").write_text(yaml.safe_dump(loc, sort_keys=False, default_flow_style=None))
    r = subprocess.run([os.path.join(host_tools, "ar_loc"), "map.yaml", "loc.yaml"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = schedule.MapData()
    out.load_yaml(str(tmp_path / "localize.yaml"))
    assert out.cap_uid == ["cap_0", "cap_1", "cap_2", "cap_3"]
    m2 = schedule.MapData()
    m2.load_yaml(str(tmp_path / "map.yaml"))
    m2.load_yaml(str(tmp_path / "loc.yaml"))
    schedule.Scheduler(m2).localize_many(3)
    assert np.abs(np.array(out.cap_pose[3]) - np.array(m2.cap_pose[3])).max() < 1e-8
    assert np.array_equal(np.array(out.cap_pose[:3]), np.array(m2.cap_pose[:3]))   # map captures untouched
