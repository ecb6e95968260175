"""ctypes binding of oracle/_build/liboracle.so (CPU ORACLE, test infrastructure).

Parity status: unpinned against real Ceres (absent here); pinned by the mpmath
known-answer vectors, finite differences and scipy (tests/test_oracle_*.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

TAG_SIZE = 0.0635  # ar_slam_util.hpp:319

CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2
REASONS = {1: "gradient", 2: "parameter", 3: "function", 4: "min_radius", 5: "max_iterations",
           6: "invalid_steps"}


class Options(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("jacobi_scaling", C.c_int32), ("num_threads", C.c_int32), ("elimination", C.c_int32),
                ("_pad", C.c_int32),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("termination", C.c_int32), ("reason", C.c_int32),
                ("n_e_blocks", C.c_int32), ("reduced_dim", C.c_int32), ("num_parameters", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double),
                ("gradient_max_norm", C.c_double), ("total_seconds", C.c_double),
                ("jacobian_seconds", C.c_double), ("linear_solver_seconds", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++, no external deps)."""
    srcs = [os.path.join(_HERE, f) for f in ("ar_oracle.cpp", "ar_oracle.hpp", "ar_oracle_capi.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"], env={k: v for k, v in os.environ.items()
                                                           if k not in ("CXX", "CXXFLAGS")})
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def default_options(**kw):
    o = Options()
    lib().oracle_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def max_threads():
    return int(lib().oracle_max_threads())


def project_block(cam, cap, tag, tag_size=TAG_SIZE, model=0):
    out = np.zeros(8)
    lib().oracle_project_block(_p(_f64(cam)), _p(_f64(cap)), _p(_f64(tag)), C.c_double(tag_size),
                               C.c_int(model), _p(out))
    return out


def evaluate(cap_idx, tag_idx, obs, cam, cap, tag, tag_size=TAG_SIZE, model=0, jacobians=True,
             num_threads=1):
    """Residuals [n_blk,8] and Ceres-layout Jacobians [n_blk,8,3], [n_blk,8,6], [n_blk,8,6]."""
    cap_idx, tag_idx = _i32(cap_idx), _i32(tag_idx)
    obs, cam, cap, tag = _f64(obs), _f64(cam), _f64(cap), _f64(tag)
    nb = len(cap_idx)
    res = np.zeros((nb, 8))
    jc = np.zeros((nb, 8, 3)) if jacobians else None
    jp = np.zeros((nb, 8, 6)) if jacobians else None
    ja = np.zeros((nb, 8, 6)) if jacobians else None
    cost = C.c_double(0)
    lib().oracle_evaluate(C.c_int(nb), _p(cap_idx, C.c_int32), _p(tag_idx, C.c_int32), _p(obs), _p(cam),
                          _p(cap), _p(tag), C.c_double(tag_size), C.c_int(model), C.c_int(num_threads),
                          C.byref(cost), _p(res), _p(jc), _p(jp), _p(ja))
    return cost.value, res, jc, jp, ja


def init_capture_pose(rect8, cam, tag_pose, tag_size=TAG_SIZE):
    out = np.zeros(6)
    lib().oracle_init_capture_pose(_p(_f64(rect8)), _p(_f64(cam)), _p(_f64(tag_pose)),
                                   C.c_double(tag_size), _p(out))
    return out


def init_tag_pose(rect8, cam, cap_pose, tag_size=TAG_SIZE):
    out = np.zeros(6)
    lib().oracle_init_tag_pose(_p(_f64(rect8)), _p(_f64(cam)), _p(_f64(cap_pose)), C.c_double(tag_size),
                               _p(out))
    return out


def compose_axis_angle(r1, r2):
    out = np.zeros(3)
    lib().oracle_compose_axis_angle(_p(_f64(r1)), _p(_f64(r2)), _p(out))
    return out


def rotate_point(aa, pt):
    out = np.zeros(3)
    lib().oracle_rotate_point(_p(_f64(aa)), _p(_f64(pt)), _p(out))
    return out


def solve(n_cap, n_tag, cap_idx, tag_idx, obs, cam, cap, tag, options=None, tag_size=TAG_SIZE, model=0,
          cam_const=False, cap_const=None, tag_const=None):
    """Run the restated ceres::Solve.  Returns (cam, cap, tag, summary dict, iteration log)."""
    o = options or default_options()
    cap_idx, tag_idx, obs = _i32(cap_idx), _i32(tag_idx), _f64(obs)
    cam = _f64(cam).copy()
    cap = _f64(cap).reshape(-1).copy()
    tag = _f64(tag).reshape(-1).copy()
    assert cap.size == 6 * n_cap and tag.size == 6 * n_tag
    cc = np.ascontiguousarray(cap_const, dtype=np.uint8) if cap_const is not None else None
    tc = np.ascontiguousarray(tag_const, dtype=np.uint8) if tag_const is not None else None
    s = Summary()
    log_cap = o.max_num_iterations + 2
    log = np.full((log_cap, 8), np.nan)
    lib().oracle_solve(C.c_int(n_cap), C.c_int(n_tag), C.c_int(len(cap_idx)), _p(cap_idx, C.c_int32),
                       _p(tag_idx, C.c_int32), _p(obs), C.c_double(tag_size), C.c_int(model),
                       C.c_int(1 if cam_const else 0), _p(cc, C.c_uint8), _p(tc, C.c_uint8), C.byref(o),
                       _p(cam), _p(cap), _p(tag), C.byref(s), _p(log), C.c_int(log_cap))
    d = s.as_dict()
    d["reason_name"] = REASONS.get(s.reason, "?")
    return cam, cap.reshape(-1, 6), tag.reshape(-1, 6), d, log[: s.iterations + 1]


def localize_batch(blk_offsets, tag_idx, obs, seed_block, cam, tag, options=None, tag_size=TAG_SIZE,
                   model=0, num_threads=1):
    o = options or default_options()
    blk_offsets, tag_idx, seed_block = _i32(blk_offsets), _i32(tag_idx), _i32(seed_block)
    obs, cam, tag = _f64(obs), _f64(cam), _f64(tag).reshape(-1)
    n = len(blk_offsets) - 1
    pose = np.zeros((n, 6))
    its = np.zeros(n, dtype=np.int32)
    cost = np.zeros(n)
    term = np.zeros(n, dtype=np.int32)
    lib().oracle_localize_batch(C.c_int(n), _p(blk_offsets, C.c_int32), _p(tag_idx, C.c_int32), _p(obs),
                                _p(seed_block, C.c_int32), C.c_int(tag.size // 6), _p(cam), _p(tag),
                                C.c_double(tag_size), C.c_int(model), C.byref(o), C.c_int(num_threads),
                                _p(pose), _p(its, C.c_int32), _p(cost), _p(term, C.c_int32))
    return pose, its, cost, term
