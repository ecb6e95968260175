#!/usr/bin/env python
"""Regenerate the demo detections (BASELINE config 1) from the reference's own
images with cv2.aruco, and freeze them in the reference's map.yaml `blocks`
format (reader: ar_slam/src/ar_slam_util.cpp:304-368).

Mirrors ArSlamSolver::loadImages (ar_slam_util.cpp:247-286): DICT_4X4_50,
minCornerDistanceRate = 0.1, ids named aruco_4X4_50_<n>, capture uids cap_<k>,
corners centred on the image (ar_slam_util.hpp:257-263).  Needs
/root/reference and OpenCV, so it only runs in the build container; the two
yaml files it writes are committed:
  demo_map_detections.yaml   img1..img3 (map build)
  demo_loc_detections.yaml   img4       (ar_loc input)
"""
import os
import sys

import cv2

REF = "/root/reference/ar_slam/resources/images"
HERE = os.path.dirname(os.path.abspath(__file__))


def detect(path):
    img = cv2.imread(path)
    assert img is not None, path
    params = cv2.aruco.DetectorParameters()
    params.minCornerDistanceRate = 0.1
    dictionary = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50)
    det = cv2.aruco.ArucoDetector(dictionary, params)
    rects, ids, _ = det.detectMarkers(img)
    h, w = img.shape[:2]
    out = []
    for r, i in zip(rects, ids.flatten()):
        pts = r.reshape(4, 2)
        flat = []
        for x, y in pts:
            flat += [float(x) - 0.5 * w, float(y) - 0.5 * h]
        out.append((int(i), flat))
    return out, w, h


def fmt(v):
    return repr(float(v))


def write(fn, images, first_idx):
    lines = ["blocks:"]
    caps = []
    tags = []
    w = h = None
    for k, img in enumerate(images):
        uid = "cap_%d" % (first_idx + k)
        dets, w, h = detect(os.path.join(REF, img))
        caps.append((uid, img))
        for tid, flat in dets:
            name = "aruco_4X4_50_%d" % tid
            if name not in tags:
                tags.append(name)
            lines += ["  - capture: %s" % uid, "    aruco: %s" % name,
                      "    aruco_rect: [%s]" % ", ".join(fmt(v) for v in flat)]
    lines.append("captures:")
    for uid, img in caps:
        lines += ["  %s:" % uid, "    inv_pose: [0, 0, 0, 0, 0, 0]",
                  "    img_fn: ar_slam/resources/images/%s" % img]
    lines.append("arucos:")
    for name in tags:
        lines += ["  %s:" % name, "    pose: [0, 0, 0, 0, 0, 0]"]
    lines += ["camera:", "  params: [3000, 0, 0]", "  width: %d" % w, "  height: %d" % h]
    with open(os.path.join(HERE, fn), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", fn, "captures", len(caps), "tags", len(tags))


if __name__ == "__main__":
    write("demo_map_detections.yaml", ["img1.jpg", "img2.jpg", "img3.jpg"], 0)
    write("demo_loc_detections.yaml", ["img4.jpg"], 3)
    sys.exit(0)
