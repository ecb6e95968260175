"""Stage-by-stage comparison of the GPU marker detector with the CPU restatement; prints where they part.
Run on the GPU box: python scripts/detect_debug.py  (writes gpurun_out/detect_debug.txt)."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ar_slam_b200 import capi, synth  # noqa: E402
from oracle import aruco_detect as A  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_detect as T  # noqa: E402

out = open(os.path.join(ROOT, "gpurun_out", "detect_debug.txt"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


gold = T.golden()
bits = synth.dict_4x4_50_bits()
for si in (2, 0, 1):
    sc = gold["scenes"][si]
    img = T.scene_image(sc)
    grey = A.to_gray(img)
    log("=== scene", sc["seed"], sc["h"], sc["w"])
    try:
        det = capi.Detector(1, sc["w"], sc["h"])
        t0 = time.time()
        res = det.detect(img[None], capi.default_detect_params(min_corner_distance_rate=0.1))
        log("detect ok, wall %.1f ms" % (1e3 * (time.time() - t0)), det.times())
        g = det.read_stage(0).reshape(sc["h"], sc["w"])
        m = det.read_stage(1).reshape(sc["h"], sc["w"])
        log("grey mismatches", int((g != grey).sum()))
        for k, win in enumerate(T.WINDOWS):
            want = A.adaptive_threshold(grey, win, 7.0) != 0
            got = (m >> k & 1) != 0
            bad = np.argwhere(want != got)
            log("window", win, "mask mismatches", len(bad), bad[:5].tolist())
        gb = sorted(T.gpu_borders(det), key=lambda t: (t[1], -t[2]))
        ob = T.oracle_borders(A, grey, A.REFERENCE_PARAMS)
        log("borders gpu", len(gb), "oracle", len(ob))
        okey = {(k, tuple(c[0]), len(c)): c for k, c in ob}
        gkey = {(w, tuple(p[0]), len(p)): p for _, w, _, p in gb}
        log("  only gpu", len(set(gkey) - set(okey)), sorted(set(gkey) - set(okey))[:8])
        log("  only oracle", len(set(okey) - set(gkey)), sorted(set(okey) - set(gkey))[:8])
        same = sum(1 for k2 in set(gkey) & set(okey) if (gkey[k2] == okey[k2]).all())
        log("  common", len(set(gkey) & set(okey)), "identical points", same)
        order_ok = [(w, tuple(p[0])) for _, w, _, p in gb] == [(k, tuple(c[0])) for k, c in ob]
        log("  order identical", order_ok)
        c = det.candidates()
        got = [(int(w), q.tolist(), bool(n), int(i), int(r) if i >= 0 else 0)
               for w, q, n, i, r in zip(c["window"], c["quad"], c["near_border"], c["id"], c["rotation"])]
        want = T.oracle_candidates(A, grey, A.REFERENCE_PARAMS, bits)
        log("candidates gpu", len(got), "oracle", len(want), "equal", got == want)
        if got != want:
            gq = {(w, json.dumps(q)): (n, i, r) for w, q, n, i, r in got}
            wq = {(w, json.dumps(q)): (n, i, r) for w, q, n, i, r in want}
            log("  quads only gpu", [k2 for k2 in gq if k2 not in wq][:6])
            log("  quads only oracle", [k2 for k2 in wq if k2 not in gq][:6])
            log("  differing id/rot", [(k2, gq[k2], wq[k2]) for k2 in gq if k2 in wq and gq[k2] != wq[k2]][:6])
        ids, corners = res[0]
        final_ok = T.as_pairs(ids, corners) == T.as_pairs(sc["ids"], sc["corners"])
        log("final", ids.tolist(), "golden", sc["ids"], "equal", final_ok)
    except Exception:
        log(traceback.format_exc())
out.close()
