#!/usr/bin/env python
"""Generate tests/golden/kat_projection.json: 50-digit known-answer vectors for
the reprojection model (reference: ar_slam/src/ar_slam_util.cpp:131-216, with
Ceres 2.0 AngleAxisRotatePoint semantics incl. the theta^2 <= DBL_EPSILON
branch).  Values AND the 8x15 Jacobian come from forward-mode duals evaluated
in mpmath at 50 digits, i.e. an implementation independent of the C++ oracle.

Run here (needs mpmath); output is committed.  Case 0 is SURVEY Appendix D.
"""
import json
import os
import sys

import mpmath as mp

mp.mp.dps = 50
EPS = mp.mpf(2) ** -52
DIRS = [(-1, -1), (1, -1), (1, 1), (-1, 1)]


class Dual:
    __slots__ = ("a", "v")

    def __init__(self, a, v=None, n=15):
        self.a = mp.mpf(a)
        self.v = list(v) if v is not None else [mp.mpf(0)] * n

    @staticmethod
    def var(a, k, n=15):
        d = Dual(a, n=n)
        d.v[k] = mp.mpf(1)
        return d

    def _c(self, o):
        return o if isinstance(o, Dual) else Dual(o, n=len(self.v))

    def __add__(self, o):
        o = self._c(o)
        return Dual(self.a + o.a, [x + y for x, y in zip(self.v, o.v)])

    __radd__ = __add__

    def __sub__(self, o):
        o = self._c(o)
        return Dual(self.a - o.a, [x - y for x, y in zip(self.v, o.v)])

    def __rsub__(self, o):
        return self._c(o) - self

    def __neg__(self):
        return Dual(-self.a, [-x for x in self.v])

    def __mul__(self, o):
        o = self._c(o)
        return Dual(self.a * o.a, [self.a * y + x * o.a for x, y in zip(self.v, o.v)])

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._c(o)
        q = self.a / o.a
        return Dual(q, [(x - q * y) / o.a for x, y in zip(self.v, o.v)])

    def __rtruediv__(self, o):
        return self._c(o) / self


def dsqrt(x):
    s = mp.sqrt(x.a)
    return Dual(s, [v / (2 * s) for v in x.v])


def dsin(x):
    return Dual(mp.sin(x.a), [mp.cos(x.a) * v for v in x.v])


def dcos(x):
    return Dual(mp.cos(x.a), [-mp.sin(x.a) * v for v in x.v])


def rotate(aa, p):
    t2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]
    if t2.a > EPS:
        th = dsqrt(t2)
        c, s = dcos(th), dsin(th)
        w = [a / th for a in aa]
        wxp = [w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]]
        tmp = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (1 - c)
        return [p[i] * c + wxp[i] * s + w[i] * tmp for i in range(3)]
    wxp = [aa[1] * p[2] - aa[2] * p[1], aa[2] * p[0] - aa[0] * p[2], aa[0] * p[1] - aa[1] * p[0]]
    return [p[i] + wxp[i] for i in range(3)]


def project(cam, cap, tag, idx, s, model):
    m = [Dual(mp.mpf(s) / 2 * DIRS[idx][0]), Dual(mp.mpf(s) / 2 * DIRS[idx][1]), Dual(0)]
    pw = rotate(tag[3:], m)
    q = [pw[i] + tag[i] + cap[i] for i in range(3)]
    pc = rotate(cap[3:], q)
    xp, yp = pc[0] / pc[2], pc[1] / pc[2]
    if model == 0:
        return cam[0] * xp, cam[0] * yp
    r2 = xp * xp + yp * yp
    dist = r2 * (cam[1] + cam[2] * r2) + 1
    return cam[0] * dist * xp, cam[0] * dist * yp


def case(name, cam, cap, tag, model=0, s="0.0635"):
    # inputs are given as decimal strings / floats; they are first rounded to
    # binary64 (what the solver sees), then lifted to 50 digits.
    camf = [float(x) for x in cam]
    capf = [float(x) for x in cap]
    tagf = [float(x) for x in tag]
    sf = float(s)
    dc = [Dual.var(camf[i], i) for i in range(3)]
    dp = [Dual.var(capf[i], 3 + i) for i in range(6)]
    da = [Dual.var(tagf[i], 9 + i) for i in range(6)]
    uv, jac = [], []
    for idx in range(4):
        u, v = project(dc, dp, da, idx, sf, model)
        for d in (u, v):
            uv.append(mp.nstr(d.a, 30))
            jac.append([mp.nstr(x, 30) for x in d.v])
    return {"name": name, "model": model, "tag_size": sf, "camera": camf, "capture": capf,
            "tag": tagf, "uv": uv, "jacobian": jac}


def main():
    cases = [
        case("survey_appendix_d", [800, 0, 0], [0.1, -0.2, 1.5, 0.05, -0.1, 0.2],
             [0.3, 0.1, 0.2, -0.3, 0.2, 0.1]),
        case("capture_identity_rotation", [3000, 0, 0], [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
             [0.05, -0.02, 1.1, 0.0, 0.0, 0.4]),
        case("both_zero_rotation", [760, 0, 0], [0.02, 0.01, 0.0, 0, 0, 0], [0.1, 0.2, 2.0, 0, 0, 0]),
        case("theta_below_eps", [760, 0, 0], [0.1, 0.1, 1.0, 6e-9, -5e-9, 7e-9],
             [0.0, 0.3, 0.5, 1e-9, 2e-9, -1e-9]),
        case("theta_just_above_eps", [760, 0, 0], [0.1, 0.1, 1.0, 1.2e-8, -0.9e-8, 1.1e-8],
             [0.0, 0.3, 0.5, 2e-8, 1e-8, -2e-8]),
        case("theta_small_1e-4", [760, 0, 0], [-0.3, 0.2, 0.8, 1e-4, -2e-4, 1.5e-4],
             [0.2, -0.1, 0.9, -3e-4, 1e-4, 2e-4]),
        case("large_rotation_near_pi", [900, 0, 0], [0.4, -0.1, 2.0, 0.3, 3.0, -0.2],
             [-0.2, 0.5, 0.7, 2.0, -1.5, 0.8]),
        case("demo_like", [758.7, 0, 0], [-0.12, 0.33, 0.05, 0.21, -0.4, 1.3],
             [0.31, -0.22, 1.42, 0.1, 0.15, -1.2]),
        case("distortion_model", [760, -0.05, 0.01], [0.1, -0.2, 1.5, 0.05, -0.1, 0.2],
             [0.3, 0.1, 0.2, -0.3, 0.2, 0.1], model=1),
        case("distortion_zero_coeffs", [760, 0.0, 0.0], [0.0, 0.1, 1.2, 0.2, 0.1, -0.3],
             [0.1, 0.1, 0.4, 0.3, -0.2, 0.5], model=1),
    ]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_projection.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_kat.py (mpmath %s, 50 digits)" % mp.__version__,
                   "columns": "camera[3] | capture[6] | tag[6]", "cases": cases}, f, indent=1)
    print("wrote", out, len(cases), "cases")
    # Appendix D spot check
    u0 = mp.mpf(cases[0]["uv"][0])
    # (the survey value used exact decimals; here inputs are rounded to binary64 first)
    assert abs(u0 - mp.mpf("105.29456120018949947")) < mp.mpf("1e-12"), u0


if __name__ == "__main__":
    sys.exit(main())
