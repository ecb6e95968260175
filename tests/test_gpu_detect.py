"""GPU marker detector (arslam_detect_markers, csrc/detect.cu) against the CPU restatement and cv2's golden corners.

Through the C-ABI (ctypes).  Stage by stage: grey + threshold bits, traced borders (the parallel 'raster-first start
survives' formulation against the sequential Suzuki-Abe scan), candidate quads with their identification; then the
detections themselves against what cv2.aruco returned for the reference's demo frames and for rendered scenes
(tests/golden/marker_golden.json).  Integer / index work: bit-exact.  Reference: aruco_detector.cpp:106,
ar_slam_util.cpp:249-268."""
import json
import os

import numpy as np
import pytest

from ar_slam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WINDOWS = (3, 13, 23)


def golden():
    with open(os.path.join(GOLD, "marker_golden.json")) as f:
        return json.load(f)


def scene_image(sc):
    return synth.render_marker_scene(sc["h"], sc["w"], synth.dict_4x4_50_bits(), sc["n_markers"], sc["seed"],
                                     noise=sc["noise"])[0]


def as_pairs(ids, corners):
    return [(int(i), np.asarray(c, np.float32).reshape(4, 2).tolist()) for i, c in zip(ids, corners)]


@pytest.fixture(scope="module")
def capi():
    import ar_slam_b200
    ar_slam_b200.load_library()
    from ar_slam_b200 import capi
    return capi


@pytest.fixture(scope="module")
def A():
    from oracle import aruco_detect
    return aruco_detect


def reference_params(capi, **kw):
    return capi.default_detect_params(min_corner_distance_rate=0.1, **kw)      # ar_slam_util.cpp:250


def test_grey_and_threshold_bits_equal_the_restatement(capi, A):
    sc = golden()["scenes"][1]                      # 768 x 1020: not a multiple of the 64 x 32 tile
    img = scene_image(sc)
    det = capi.Detector(2, sc["w"], sc["h"])
    grey_in = A.to_gray(img[::-1].copy())
    det.detect(np.stack([img, img[::-1]]), reference_params(capi))
    g = det.read_stage(0).reshape(2, sc["h"], sc["w"])
    m = det.read_stage(1).reshape(2, sc["h"], sc["w"])
    for b, grey in enumerate((A.to_gray(img), grey_in)):
        assert (g[b] == grey).all()
        for k, win in enumerate(WINDOWS):
            want = A.adaptive_threshold(grey, win, 7.0) != 0
            got = (m[b] >> k & 1) != 0
            assert (want == got).all(), (b, win, int((want != got).sum()))
    det.detect(grey_in[None], reference_params(capi))              # one-channel input
    assert (det.read_stage(0).reshape(sc["h"], sc["w"]) == grey_in).all()
    assert (det.read_stage(1).reshape(sc["h"], sc["w"]) == m[1]).all()


def oracle_borders(A, grey, params):
    """{(window, first point, kind-free key): points} for the borders long enough to matter, from the sequential scan."""
    big = max(grey.shape)
    lo, hi = int(params["minMarkerPerimeterRate"] * big), int(params["maxMarkerPerimeterRate"] * big)
    out = []
    for k, win in enumerate(WINDOWS):
        for c in A.find_contours(A.adaptive_threshold(grey, win, 7.0)):
            if lo <= len(c) <= hi:
                out.append((k, c))
    return out


def gpu_borders(det):
    table, pts = det.read_stage(2), det.read_stage(3)
    out = []
    for img, win, disc, ln, off in table:
        p = pts[off:off + ln]
        out.append((int(img), int(win), int(disc), np.stack([p & 0xffff, p >> 16], axis=1).astype(np.int32)))
    return out


def test_traced_borders_equal_the_sequential_scan(capi, A):
    """Same set of borders, same first point, same direction, same order of discovery, on a noisy scene and on
    random bitmaps' worth of clutter (window 3 of a noisy frame is mostly clutter)."""
    sc = golden()["scenes"][2]                      # 360 x 500, noise 9
    img = scene_image(sc)
    grey = A.to_gray(img)
    det = capi.Detector(1, sc["w"], sc["h"])
    det.detect(img[None], reference_params(capi))
    got = sorted(gpu_borders(det), key=lambda t: (t[1], -t[2]))
    want = oracle_borders(A, grey, A.REFERENCE_PARAMS)
    assert len(got) == len(want)
    for (_, win, _, p), (k, c) in zip(got, want):
        assert win == k and p.shape == c.shape and (p == c).all()


def oracle_candidates(A, grey, params, bits):
    out = []
    for k, win in enumerate(WINDOWS):
        quads, _, flags = A.find_marker_contours(A.adaptive_threshold(grey, win, 7.0), params)
        for q, f in zip(quads, flags):
            q = A.reorder_corners(q)
            r = A.identify(A.extract_bits(grey, q, params), bits, 1, params)
            out.append((k, q.tolist(), bool(f), -1 if r is None else r[0], 0 if r is None else r[1]))
    return out


def test_candidate_quads_and_identification_equal_the_restatement(capi, A):
    bits = synth.dict_4x4_50_bits()
    for sc in golden()["scenes"][:3]:
        img = scene_image(sc)
        det = capi.Detector(1, sc["w"], sc["h"])
        det.detect(img[None], reference_params(capi))
        c = det.candidates()
        got = [(int(w), q.tolist(), bool(n), int(i), int(r) if i >= 0 else 0)
               for w, q, n, i, r in zip(c["window"], c["quad"], c["near_border"], c["id"], c["rotation"])]
        want = oracle_candidates(A, A.to_gray(img), A.REFERENCE_PARAMS, bits)
        assert got == want, sc["seed"]


def test_demo_frames_give_cv2s_corners(capi):
    gold = golden()
    frames = np.load(os.path.join(GOLD, "demo_gray.npz"))
    det = capi.Detector(2, 1020, 768)
    res = det.detect(np.stack([frames["img1"], frames["img4"]]), reference_params(capi))
    for (ids, corners), name in zip(res, ("img1", "img4")):
        assert as_pairs(ids, corners) == as_pairs(gold["demo"][name]["ids"], gold["demo"][name]["corners"])
    t = det.times()
    assert t["launches"] == 5 and t["total_ms"] > 0


def test_rendered_scenes_give_cv2s_corners_in_batches(capi):
    gold = golden()
    by_shape = {}
    for sc in gold["scenes"]:
        by_shape.setdefault((sc["h"], sc["w"]), []).append(sc)
    n = 0
    for (h, w), scs in by_shape.items():
        det = capi.Detector(len(scs), w, h)
        res = det.detect(np.stack([scene_image(sc) for sc in scs]), reference_params(capi))
        for (ids, corners), sc in zip(res, scs):
            assert as_pairs(ids, corners) == as_pairs(sc["ids"], sc["corners"]), sc["seed"]
            n += len(ids)
    assert n >= 60


def test_live_cv2_and_other_dictionaries(capi):
    """cv2 on the box itself (same image as the build container): default parameters, and DICT_5X5_100 /
    DICT_6X6_250 of aruco_detector.cpp:148-152 through arslam_detector_set_dictionary."""
    cv2 = pytest.importorskip("cv2")
    for name, size in (("DICT_4X4_50", 4), ("DICT_5X5_100", 5), ("DICT_6X6_250", 6)):
        d = cv2.aruco.getPredefinedDictionary(getattr(cv2.aruco, name))
        bits = np.array([cv2.aruco.Dictionary.getBitsFromByteList(d.bytesList[m:m + 1], size)
                         for m in range(d.bytesList.shape[0])], np.uint8)
        det = capi.Detector(2, 640, 480)
        det.set_dictionary(bits, d.maxCorrectionBits)
        imgs = np.stack([synth.render_marker_scene(480, 640, bits, 7, 200 + size + s, noise=3.0)[0] for s in (0, 10)])
        res = det.detect(imgs, capi.default_detect_params())
        ref = cv2.aruco.ArucoDetector(d, cv2.aruco.DetectorParameters())
        found = 0
        for (ids, corners), img in zip(res, imgs):
            r, i, _ = ref.detectMarkers(img)
            assert as_pairs(ids, corners) == as_pairs([] if i is None else i.ravel(), r), name
            found += len(ids)
        assert found >= 8, name
        det2 = capi.Detector(2, 640, 480)                  # the same dictionary from the tables the library embeds
        det2.set_predefined_dictionary(name[5:])
        res2 = det2.detect(imgs, capi.default_detect_params())
        assert [as_pairs(*a) for a in res2] == [as_pairs(*a) for a in res], name
    with pytest.raises(capi.ArslamError):
        det2.set_predefined_dictionary("7X7_1000")


def test_large_frame_with_long_borders_matches_live_cv2(capi):
    """1600 x 1200: borders of up to 6400 points are admissible, so the checkpoints of the border following are
    256 steps apart instead of 128; big markers give borders of a few thousand points."""
    cv2 = pytest.importorskip("cv2")
    bits = synth.dict_4x4_50_bits()
    img = synth.render_marker_scene(1200, 1600, bits, 6, 77, noise=2.0)[0]
    big = np.kron(bits[7], np.ones((150, 150), np.uint8))                       # one marker 900 pixels wide
    img[100:1300 - 100, 200:1300] = 235
    img[250:1150, 300:1200] = 30
    img[400:1000, 450:1050] = np.where(big[..., None] > 0, 230, 30)
    det = capi.Detector(1, 1600, 1200)
    ids, corners = det.detect(img[None], capi.default_detect_params())[0]
    r, i, _ = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50),
                                      cv2.aruco.DetectorParameters()).detectMarkers(img)
    assert 7 in list(ids)
    assert as_pairs(ids, corners) == as_pairs(i.ravel(), r)
    table = det.read_stage(2)
    assert table[:, 3].max() > 3000                                             # the big marker's border


def test_frame_sizes_that_are_not_multiples_of_the_vector_widths(capi, A):
    """333 x 251 and 131 x 97: the threshold kernel's 4-pixel stores and the start scan's 8-pixel words end inside
    the zero padding; grey, threshold bits, ids and corners against the restatement (cv2's default parameters)."""
    bits = synth.dict_4x4_50_bits()
    for h, w in ((251, 333), (97, 131)):
        img = synth.render_marker_scene(h, w, bits, 3, 5, noise=3.0)[0]
        det = capi.Detector(1, w, h)
        ids, corners = det.detect(img[None], capi.default_detect_params())[0]
        grey = A.to_gray(img)
        assert (det.read_stage(0).reshape(h, w) == grey).all()
        m = det.read_stage(1).reshape(h, w)
        for k, win in enumerate(WINDOWS):
            assert (((m >> k & 1) != 0) == (A.adaptive_threshold(grey, win, 7.0) != 0)).all()
        ocorners, oids = A.detect_markers(img, bits, 1, A.DEFAULTS)
        assert len(oids) == 3 and as_pairs(ids, corners) == as_pairs(oids, ocorners)


def test_frames_already_on_the_device_and_errors(capi):
    import torch
    sc = golden()["scenes"][0]
    img = scene_image(sc)
    det = capi.Detector(1, sc["w"], sc["h"])
    host = det.detect(img[None], reference_params(capi))[0]
    dev = torch.from_numpy(img[None].copy()).cuda()
    torch.cuda.synchronize()
    on_dev = det.detect(None, reference_params(capi), device_ptr=dev.data_ptr(), shape=dev.shape)[0]
    assert as_pairs(*host) == as_pairs(*on_dev) == as_pairs(sc["ids"], sc["corners"])
    blank = np.full((1, sc["h"], sc["w"]), 128, np.uint8)
    assert len(det.detect(blank, reference_params(capi))[0][0]) == 0
    with pytest.raises(capi.ArslamError):
        det.detect(np.zeros((2, sc["h"], sc["w"]), np.uint8))                  # more frames than reserved
    with pytest.raises(capi.ArslamError):
        det.detect(np.zeros((1, sc["h"], sc["w"], 2), np.uint8))               # two channels
    with pytest.raises(capi.ArslamError):
        det.detect(img[None], reference_params(capi), max_markers=2)           # more markers than the caller allows
    with pytest.raises(capi.ArslamError):
        det.detect(img[None], capi.default_detect_params(adaptive_thresh_win_size_min=4))
    assert as_pairs(*det.detect(img[None], reference_params(capi))[0]) == as_pairs(sc["ids"], sc["corners"])
