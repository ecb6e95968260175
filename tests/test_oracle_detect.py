"""The CPU restatement of the marker detector (oracle/aruco_detect.py) against the real thing.

Stage by stage against the cv2 function each stage restates (cv2 is in the image, here and on the GPU box; the
stage tests skip without it) and end to end against golden corners that cv2.aruco produced for the reference's own
demo frames and for rendered scenes (tests/golden/make_marker_golden.py -> marker_golden.json, demo_gray.npz).
Reference call sites: aruco_detector.cpp:106, ar_slam_util.cpp:249-268."""
import json
import os

import numpy as np
import pytest

from ar_slam_b200 import synth
from oracle import aruco_detect as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden():
    with open(os.path.join(GOLD, "marker_golden.json")) as f:
        return json.load(f)


def scene_image(sc):
    return synth.render_marker_scene(sc["h"], sc["w"], synth.dict_4x4_50_bits(), sc["n_markers"], sc["seed"],
                                     noise=sc["noise"])[0]


def as_pairs(ids, corners):
    return [(int(i), np.asarray(c, np.float32).reshape(4, 2).tolist()) for i, c in zip(ids, corners)]


def test_dictionary_table_is_cv2s():
    cv2 = pytest.importorskip("cv2")
    d = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50)
    bits = synth.dict_4x4_50_bits()
    assert (bits == A.dictionary_bits(d.bytesList, 4)).all()
    for m in range(50):
        assert (bits[m] == cv2.aruco.Dictionary.getBitsFromByteList(d.bytesList[m:m + 1], 4)).all()
        for r in range(4):      # identify() restated: same marker and rotation for every turned code
            inner = np.ascontiguousarray(np.rot90(bits[m], r))
            ok, idx, rot = d.identify(inner, 0.6)
            full = np.zeros((6, 6), np.uint8)
            full[1:5, 1:5] = inner
            assert ok and (idx, rot) == A.identify(full, bits, d.maxCorrectionBits, A.REFERENCE_PARAMS)


def test_embedded_tables_of_the_other_dictionaries_are_cv2s():
    """csrc/aruco_dictionaries.inc (arslam_detector_set_predefined_dictionary "5X5_100" / "6X6_250") against cv2."""
    cv2 = pytest.importorskip("cv2")
    import re
    text = open(os.path.join(os.path.dirname(GOLD), "..", "ar_slam_b200", "csrc", "aruco_dictionaries.inc")).read()
    for name, size in (("DICT_5X5_100", 5), ("DICT_6X6_250", 6)):
        d = cv2.aruco.getPredefinedDictionary(getattr(cv2.aruco, name))
        body = re.search(r"k%s\[(\d+)\] = \{(.*?)\};" % name, text, re.S)
        codes = [int(c, 16) for c in re.findall(r"0x([0-9a-f]+)ull", body.group(2))]
        assert len(codes) == int(body.group(1)) == d.bytesList.shape[0]
        assert int(re.search(r"k%s_max_correction = (\d+)" % name, text).group(1)) == d.maxCorrectionBits
        for m, c in enumerate(codes):
            want = cv2.aruco.Dictionary.getBitsFromByteList(d.bytesList[m:m + 1], size).ravel()
            assert c == int("".join(str(int(b)) for b in want), 2)


def test_grey_threshold_contours_polygons_match_cv2():
    cv2 = pytest.importorskip("cv2")
    sc = golden()["scenes"][2]
    img = scene_image(sc)
    g = A.to_gray(img)
    assert (g == cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)).all()
    n_poly = 0
    for win in (3, 13, 23):
        t = A.adaptive_threshold(g, win, 7.0)
        assert (t == cv2.adaptiveThreshold(g, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, win, 7)).all()
        mine = A.find_contours(t)
        ref, _ = cv2.findContours(t, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        assert len(mine) == len(ref)
        for a, b in zip(mine, ref):
            b = b.reshape(-1, 2)
            assert a.shape == b.shape and (a == b).all()
            if len(a) >= 8:
                for rate in (0.03, 0.08):
                    eps = len(a) * rate
                    pa = A.approx_poly_dp(a, eps)
                    pb = cv2.approxPolyDP(b.reshape(-1, 1, 2), eps, True).reshape(-1, 2)
                    assert pa.shape == pb.shape and (pa == pb).all()
                    if len(pa) >= 3:
                        assert A.is_contour_convex(pa) == cv2.isContourConvex(pb.reshape(-1, 1, 2))
                    n_poly += 1
    assert n_poly > 100


def test_random_bitmaps_contours_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for dens in (0.25, 0.5, 0.75):
        for _ in range(4):
            t = ((rng.random((37, 51)) < dens) * 255).astype(np.uint8)
            ref, _ = cv2.findContours(t, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
            mine = A.find_contours(t)
            assert len(mine) == len(ref)
            assert all((a == b.reshape(-1, 2)).all() for a, b in zip(mine, ref))


def test_perspective_removal_and_otsu_match_cv2():
    cv2 = pytest.importorskip("cv2")
    sc = golden()["scenes"][0]
    g = A.to_gray(scene_image(sc))
    dst = np.array([[0, 0], [23, 0], [23, 23], [0, 23]], np.float32)
    for quad in sc["corners"]:
        q = np.array(quad, np.float32)
        m = cv2.getPerspectiveTransform(q, dst)
        ref = cv2.warpPerspective(g, m, (24, 24), flags=cv2.INTER_NEAREST)
        mine = A.warp_nearest(g, A.perspective_transform(q, dst), 24)
        assert (ref == mine).all()
        assert A.otsu_threshold(mine) == int(cv2.threshold(ref, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)[0])


def test_pipeline_reproduces_cv2_on_the_demo_frames():
    gold = golden()
    frames = np.load(os.path.join(GOLD, "demo_gray.npz"))
    bits = synth.dict_4x4_50_bits()
    for name in ("img1", "img4"):
        corners, ids = A.detect_markers(frames[name], bits, 1)
        assert as_pairs(ids, corners) == as_pairs(gold["demo"][name]["ids"], gold["demo"][name]["corners"])


def test_pipeline_reproduces_cv2_on_rendered_scenes():
    bits = synth.dict_4x4_50_bits()
    n = 0
    for sc in golden()["scenes"][:6]:
        corners, ids = A.detect_markers(scene_image(sc), bits, 1)
        assert as_pairs(ids, corners) == as_pairs(sc["ids"], sc["corners"]), sc["seed"]
        n += len(ids)
    assert n >= 40


def test_quad_too_near_the_border_takes_its_group_with_it():
    """Golden scene 14: marker 49 sits at the top edge; the quad around its quiet zone is too near the image border,
    and cv2 4.13 drops that quad only AFTER it has absorbed the marker's own quads, so nothing is reported there.
    Dropping it before the grouping (as the 4.5.4 sources read) would report the marker."""
    sc = [s for s in golden()["scenes"] if s["seed"] == 14][0]
    assert 49 in sc["placed"] and 49 not in sc["ids"]
    bits = synth.dict_4x4_50_bits()
    corners, ids = A.detect_markers(scene_image(sc), bits, 1)
    assert as_pairs(ids, corners) == as_pairs(sc["ids"], sc["corners"])


def test_other_dictionaries_against_live_cv2():
    """DICT_5X5_100 and DICT_6X6_250 (aruco_detector.cpp:148-152): 28 and 32 pixel canonical images, error correction
    allowed (maxCorrectionBits 3 and 5 at errorCorrectionRate 0.6)."""
    cv2 = pytest.importorskip("cv2")
    for name, size in (("DICT_5X5_100", 5), ("DICT_6X6_250", 6)):
        d = cv2.aruco.getPredefinedDictionary(getattr(cv2.aruco, name))
        bits = A.dictionary_bits(d.bytesList, size)
        for m in range(0, bits.shape[0], 7):
            assert (bits[m] == cv2.aruco.Dictionary.getBitsFromByteList(d.bytesList[m:m + 1], size)).all()
        img = synth.render_marker_scene(360, 500, bits, 5, 300 + size, noise=3.0)[0]
        r, i, _ = cv2.aruco.ArucoDetector(d, cv2.aruco.DetectorParameters()).detectMarkers(img)
        corners, ids = A.detect_markers(img, bits, d.maxCorrectionBits, A.DEFAULTS, marker_size=size)
        assert len(ids) >= 3
        assert as_pairs(ids, corners) == as_pairs(i.ravel(), r)


def test_pipeline_against_live_cv2_with_default_parameters():
    """Not the reference's 0.1 but cv2's default minCornerDistanceRate, on scenes that are not in the golden file."""
    cv2 = pytest.importorskip("cv2")
    bits = synth.dict_4x4_50_bits()
    det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), cv2.aruco.DetectorParameters())
    for seed in (101, 102):
        img = synth.render_marker_scene(360, 500, bits, 6, seed, noise=6.0)[0]
        r, i, _ = det.detectMarkers(img)
        corners, ids = A.detect_markers(img, bits, 1, A.DEFAULTS)
        assert as_pairs(ids, corners) == as_pairs([] if i is None else i.ravel(), r)
