#!/usr/bin/env python
"""Golden vectors for the marker detector, made with the real cv2.aruco.

Runs only in the build container (needs /root/reference and OpenCV).  Writes, next to this file:
  demo_gray.npz        grey versions (cv2.cvtColor) of the reference's demo images img1.jpg and img4.jpg
                       (ar_slam/resources/images), so the GPU box can run the detector on real frames
  marker_golden.json   ids and corners from cv2.aruco.ArucoDetector.detectMarkers with the reference's settings
                       (DICT_4X4_50, minCornerDistanceRate = 0.1: ar_slam_util.cpp:249-252) for those frames and
                       for rendered scenes (ar_slam_b200.synth.render_marker_scene, seeds listed in the file)
"""
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from ar_slam_b200 import synth  # noqa: E402

REF = "/root/reference/ar_slam/resources/images"
SCENES = [dict(seed=s, h=[480, 768, 360][s % 3], w=[640, 1020, 500][s % 3], n_markers=8,
               noise=[2.0, 5.0, 9.0][s % 3]) for s in list(range(9)) + [14, 17, 20]]
# seed 14: a marker whose quiet zone touches the image border (the enclosing quad is too near the border and takes the
# marker's own quads with it: cv2 reports nothing there)


def detector(rate=0.1):
    p = cv2.aruco.DetectorParameters()
    p.minCornerDistanceRate = rate
    return cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), p)


def run(det, img):
    r, i, _ = det.detectMarkers(img)
    if i is None:
        return dict(ids=[], corners=[])
    return dict(ids=[int(k) for k in i.ravel()], corners=[x.reshape(4, 2).tolist() for x in r])


def main():
    det = detector()
    gray = {}
    out = dict(opencv=cv2.__version__, params=dict(minCornerDistanceRate=0.1), demo={}, scenes=[])
    for name in ("img1", "img4"):
        g = cv2.cvtColor(cv2.imread(os.path.join(REF, name + ".jpg")), cv2.COLOR_BGR2GRAY)
        gray[name] = g
        out["demo"][name] = run(det, g)
    np.savez_compressed(os.path.join(HERE, "demo_gray.npz"), **gray)
    bits = synth.dict_4x4_50_bits()
    for sc in SCENES:
        img, placed = synth.render_marker_scene(sc["h"], sc["w"], bits, sc["n_markers"], sc["seed"], noise=sc["noise"])
        out["scenes"].append(dict(sc, placed=placed, **run(det, img)))
    with open(os.path.join(HERE, "marker_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: v["ids"] for k, v in out["demo"].items()}, [len(s["ids"]) for s in out["scenes"]])


if __name__ == "__main__":
    main()
