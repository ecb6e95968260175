"""ctypes mirror of include/ar_slam_b200.h (the C-ABI a ROS2 / CLI host binds).

Argument meaning and error behaviour follow the header; Python only moves numpy
arrays across the boundary.  If the shared library is missing this module
raises -- there is deliberately no fallback implementation.
"""
import ctypes as C
import os

import numpy as np

ELIM_AUTO, ELIM_TAGS, ELIM_CAPTURES = 0, 1, 2
LINSOLVE_AUTO, LINSOLVE_DENSE, LINSOLVE_PCG = 0, 1, 2
CONVERGENCE, NO_CONVERGENCE, FAILURE = 0, 1, 2
REASONS = {1: "gradient", 2: "parameter", 3: "function", 4: "min_radius", 5: "max_iterations",
           6: "invalid_steps"}
ERR_NO_DEVICE = -3

_HERE = os.path.dirname(os.path.abspath(__file__))


class ArslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("arslam error %d: %s" % (code, msg))
        self.code = code


class Options(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
                ("jacobi_scaling", C.c_int32), ("elimination", C.c_int32), ("linear_solver", C.c_int32),
                ("pcg_max_iterations", C.c_int32), ("num_intrinsics", C.c_int32), ("verbose", C.c_int32),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("pcg_tolerance", C.c_double), ("pcg_q_tolerance", C.c_double),
                ("tag_size", C.c_double),
                ("dense_max_dim", C.c_int64)]


class Summary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("termination", C.c_int32), ("reason", C.c_int32),
                ("eliminated_side", C.c_int32), ("linear_solver", C.c_int32), ("reduced_dim", C.c_int32),
                ("num_jacobian_evals", C.c_int64), ("num_cost_evals", C.c_int64),
                ("linear_solver_iterations", C.c_int64), ("gpu_launches", C.c_int64),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double),
                ("gradient_max_norm", C.c_double), ("total_ms", C.c_double), ("eval_ms", C.c_double),
                ("linsolve_ms", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["reason_name"] = REASONS.get(self.reason, "?")
        return d


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("total_ms", C.c_double), ("launches", C.c_int64),
                ("algorithmic_bytes", C.c_double)]


def library_path():
    return os.path.join(_HERE, "lib", "libar_slam_b200.so")


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise ArslamError(-2, "%s is missing: run `python -m ar_slam_b200.build` "
                                  "(there is no CPU fallback)" % path)
        lib = C.CDLL(path)
        lib.arslam_last_error.restype = C.c_char_p
        lib.arslam_last_error.argtypes = [C.c_void_p]
        lib.arslam_create.argtypes = [C.c_int, C.POINTER(Options), C.POINTER(C.c_void_p)]
        lib.arslam_destroy.argtypes = [C.c_void_p]
        lib.arslam_destroy.restype = None
        _lib = lib
    return _lib


def default_options(**kw):
    o = Options()
    load_library().arslam_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Solver:
    """One `arslam_solver` handle (== the `problem_` member of one ArSlamSolver)."""

    def __init__(self, device=0, options=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.options = options or default_options()
        rc = self._lib.arslam_create(C.c_int(device), C.byref(self.options), C.byref(self._h))
        if rc != 0:
            raise ArslamError(rc, (self._lib.arslam_last_error(None) or b"").decode())
        self.n_cap = self.n_tag = self.n_blk = 0

    def close(self):
        if self._h:
            self._lib.arslam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise ArslamError(rc, (self._lib.arslam_last_error(self._h) or b"").decode())
        return rc

    def set_options(self, options):
        self.options = options
        self._check(self._lib.arslam_set_options(self._h, C.byref(options)))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle (e.g. torch.cuda.current_stream().cuda_stream) or None."""
        self._check(self._lib.arslam_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_profiling(self, on):
        self._check(self._lib.arslam_set_profiling(self._h, C.c_int(1 if on else 0)))

    def set_tuning(self, key, value):
        """Kernel-variant switch for A/B measurements (see arslam_set_tuning in the header)."""
        self._lib.arslam_set_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        self._check(self._lib.arslam_set_tuning(self._h, key.encode(), C.c_int64(int(value))))

    def kernel_times(self):
        buf = (KernelTime * 64)()
        n = self._check(self._lib.arslam_kernel_times(self._h, buf, C.c_int32(64)))
        return [{"name": buf[i].name.decode(), "total_ms": buf[i].total_ms, "launches": buf[i].launches,
                 "algorithmic_bytes": buf[i].algorithmic_bytes} for i in range(n)]

    def set_problem(self, n_cap, n_tag, cap_idx, tag_idx, rect8):
        cap_idx, tag_idx, rect8 = _i32(cap_idx), _i32(tag_idx), _f64(rect8).reshape(-1)
        if rect8.size != 8 * len(cap_idx) or len(tag_idx) != len(cap_idx):
            raise ValueError("cap_idx, tag_idx and rect8 disagree on the number of blocks")
        self._check(self._lib.arslam_set_problem(self._h, C.c_int64(n_cap), C.c_int64(n_tag),
                                                 C.c_int64(len(cap_idx)), _p(cap_idx, C.c_int32),
                                                 _p(tag_idx, C.c_int32), _p(rect8)))
        self.n_cap, self.n_tag, self.n_blk = int(n_cap), int(n_tag), len(cap_idx)

    def append_blocks(self, n_cap, n_tag, cap_idx, tag_idx, rect8):
        """Adds blocks to the problem on the device (n_cap / n_tag: new totals); set_params must follow."""
        cap_idx, tag_idx, rect8 = _i32(cap_idx), _i32(tag_idx), _f64(rect8).reshape(-1)
        if rect8.size != 8 * len(cap_idx) or len(tag_idx) != len(cap_idx):
            raise ValueError("cap_idx, tag_idx and rect8 disagree on the number of blocks")
        had = getattr(self, "n_blk", 0)
        self._check(self._lib.arslam_append_blocks(self._h, C.c_int64(n_cap), C.c_int64(n_tag),
                                                   C.c_int64(len(cap_idx)), _p(cap_idx, C.c_int32),
                                                   _p(tag_idx, C.c_int32), _p(rect8)))
        self.n_cap, self.n_tag, self.n_blk = int(n_cap), int(n_tag), had + len(cap_idx)

    def set_params(self, cam, cap, tag):
        cam, cap, tag = _f64(cam), _f64(cap).reshape(-1), _f64(tag).reshape(-1)
        if cam.size != 3 or cap.size != 6 * self.n_cap or tag.size != 6 * self.n_tag:
            raise ValueError("parameter array sizes do not match the problem")
        self._check(self._lib.arslam_set_params(self._h, _p(cam), _p(cap), _p(tag)))

    # device-resident parameters (incremental schedules)
    def set_camera(self, cam):
        cam = _f64(cam)
        self._lib.arslam_set_camera.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._check(self._lib.arslam_set_camera(self._h, _p(cam)))

    def set_poses(self, which, first, pose6):
        pose6 = _f64(pose6).reshape(-1, 6)
        self._lib.arslam_set_poses.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
        self._check(self._lib.arslam_set_poses(self._h, which, first, len(pose6), _p(pose6)))

    def get_poses(self, which, first, count):
        out = np.zeros((count, 6))
        self._lib.arslam_get_poses.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
        self._check(self._lib.arslam_get_poses(self._h, which, first, count, _p(out)))
        return out

    def seed_captures(self, cap_idx, tag_idx, rect8):
        cap_idx, tag_idx, rect8 = _i32(cap_idx), _i32(tag_idx), _f64(rect8).reshape(-1)
        self._lib.arslam_seed_captures.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        self._check(self._lib.arslam_seed_captures(self._h, len(cap_idx), _p(cap_idx, C.c_int32), _p(tag_idx, C.c_int32), _p(rect8)))

    def seed_tags(self, tag_idx, cap_idx, rect8):
        cap_idx, tag_idx, rect8 = _i32(cap_idx), _i32(tag_idx), _f64(rect8).reshape(-1)
        self._lib.arslam_seed_tags.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        self._check(self._lib.arslam_seed_tags(self._h, len(tag_idx), _p(tag_idx, C.c_int32), _p(cap_idx, C.c_int32), _p(rect8)))

    def get_params(self, out=None):
        """(camera3, cap_pose [n_cap, 6], tag_pose [n_tag, 6]); out = caller-owned (cap, tag) arrays, e.g. pinned.
        In a multi-GPU solve only the rank's own capture range of cap_pose is written."""
        cam = np.zeros(3)
        cap, tag = out if out is not None else (np.zeros((self.n_cap, 6)), np.zeros((self.n_tag, 6)))
        if cap.shape != (self.n_cap, 6) or tag.shape != (self.n_tag, 6) or cap.dtype != np.float64 or tag.dtype != np.float64 \
                or not cap.flags.c_contiguous or not tag.flags.c_contiguous:
            raise ValueError("out = (cap [n_cap, 6] f64, tag [n_tag, 6] f64), C-contiguous")
        self._check(self._lib.arslam_get_params(self._h, _p(cam), _p(cap), _p(tag)))
        return cam, cap, tag

    def set_constant(self, camera=False, cap=None, tag=None):
        """SetParameterBlockConstant: camera (all intrinsics), per-capture and per-tag masks (bool arrays or None)."""
        cm = np.ascontiguousarray(cap, dtype=np.uint8) if cap is not None else None
        tm = np.ascontiguousarray(tag, dtype=np.uint8) if tag is not None else None
        if (cm is not None and cm.size != self.n_cap) or (tm is not None and tm.size != self.n_tag):
            raise ValueError("mask sizes do not match the problem")
        self._lib.arslam_set_constant.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        self._check(self._lib.arslam_set_constant(self._h, C.c_int(1 if camera else 0), _p(cm, C.c_uint8), _p(tm, C.c_uint8)))

    def normal_equations(self):
        """(eliminated_side, blk_cap, blk_tag, W [n_blk,6,6], H_cap [n_cap,33], H_tag [n_tag,33], camera4)."""
        nb = self.n_blk
        side = C.c_int32(0)
        bc, bt = np.zeros(nb, dtype=np.int32), np.zeros(nb, dtype=np.int32)
        W = np.zeros((nb, 6, 6))
        Hc, Ht, cam4 = np.zeros((self.n_cap, 33)), np.zeros((self.n_tag, 33)), np.zeros(4)
        self._check(self._lib.arslam_get_normal_equations(self._h, C.byref(side), _p(bc, C.c_int32), _p(bt, C.c_int32),
                                                          _p(W), _p(Hc), _p(Ht), _p(cam4)))
        return side.value, bc, bt, W, Hc, Ht, cam4

    def evaluate(self, jacobians=True, residuals=True):
        nb = self.n_blk
        res = np.zeros((nb, 8)) if residuals else None
        jc = np.zeros((nb, 8, 3)) if jacobians else None
        jp = np.zeros((nb, 8, 6)) if jacobians else None
        ja = np.zeros((nb, 8, 6)) if jacobians else None
        cost = C.c_double(0)
        self._check(self._lib.arslam_evaluate(self._h, C.byref(cost), _p(res), _p(jc), _p(jp), _p(ja)))
        return cost.value, res, jc, jp, ja

    def solve(self, log=True):
        s = Summary()
        rows = self.options.max_num_iterations + 2
        lg = np.full((rows, 8), np.nan) if log else None
        self._check(self._lib.arslam_solve(self._h, C.byref(s), _p(lg), C.c_int32(rows if log else 0)))
        return s.as_dict(), (lg[: s.iterations + 1] if log else None)

    def localize_batch(self, blk_offsets, tag_idx, rect8, seed_block, cam, tag_pose, out=None):
        blk_offsets, tag_idx, seed_block = _i32(blk_offsets), _i32(tag_idx), _i32(seed_block)
        rect8, cam, tag_pose = _f64(rect8).reshape(-1), _f64(cam), _f64(tag_pose).reshape(-1)
        n = len(blk_offsets) - 1
        if out is not None:  # caller-owned result arrays (e.g. pinned memory)
            pose, its, cost, term = out
            if pose.shape != (n, 6) or pose.dtype != np.float64 or its.dtype != np.int32 or cost.dtype != np.float64 \
                    or term.dtype != np.int32 or len(its) != n or len(cost) != n or len(term) != n:
                raise ValueError("out = (pose [n,6] f64, iterations [n] i32, final_cost [n] f64, termination [n] i32)")
        else:
            pose = np.zeros((n, 6))
            its = np.zeros(n, dtype=np.int32)
            cost = np.zeros(n)
            term = np.zeros(n, dtype=np.int32)
        self._check(self._lib.arslam_localize_batch(
            self._h, C.c_int64(n), _p(blk_offsets, C.c_int32), _p(tag_idx, C.c_int32), _p(rect8),
            _p(seed_block, C.c_int32), C.c_int64(tag_pose.size // 6), _p(cam), _p(tag_pose), _p(pose),
            _p(its, C.c_int32), _p(cost), _p(term, C.c_int32)))
        return pose, its, cost, term

    # multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = load_library().arslam_comm_unique_id(buf)
        if rc != 0:
            raise ArslamError(rc, (load_library().arslam_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, rank, world, uid):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._check(self._lib.arslam_comm_init(self._h, C.c_int(rank), C.c_int(world), buf))


class DetectParams(C.Structure):
    """arslam_detect_params: the cv::aruco::DetectorParameters fields the path reads."""
    _fields_ = [("adaptive_thresh_win_size_min", C.c_int32), ("adaptive_thresh_win_size_max", C.c_int32),
                ("adaptive_thresh_win_size_step", C.c_int32), ("min_distance_to_border", C.c_int32),
                ("marker_border_bits", C.c_int32), ("perspective_remove_pixel_per_cell", C.c_int32),
                ("adaptive_thresh_constant", C.c_double), ("min_marker_perimeter_rate", C.c_double),
                ("max_marker_perimeter_rate", C.c_double), ("polygonal_approx_accuracy_rate", C.c_double),
                ("min_corner_distance_rate", C.c_double), ("min_marker_distance_rate", C.c_double),
                ("min_group_distance", C.c_double), ("perspective_remove_ignored_margin_per_cell", C.c_double),
                ("max_erroneous_bits_in_border_rate", C.c_double), ("min_otsu_std_dev", C.c_double),
                ("error_correction_rate", C.c_double)]


def default_detect_params(**kw):
    p = DetectParams()
    load_library().arslam_detect_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Detector:
    """One `arslam_detector` handle: cv::aruco::detectMarkers for batches of equally sized frames on the GPU
    (aruco_detector.cpp:106, ar_slam_util.cpp:268)."""

    def __init__(self, max_images, max_width, max_height, device=0):
        self._lib = lib = load_library()
        lib.arslam_detector_last_error.restype = C.c_char_p
        lib.arslam_detector_last_error.argtypes = [C.c_void_p]
        lib.arslam_detector_create.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        lib.arslam_detector_destroy.argtypes = [C.c_void_p]
        lib.arslam_detector_destroy.restype = None
        lib.arslam_detector_set_dictionary.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
        lib.arslam_detect_markers.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                              C.POINTER(DetectParams), C.c_int32, C.POINTER(C.c_int32),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        lib.arslam_detector_candidates.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                   C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                   C.POINTER(C.c_int32)]
        lib.arslam_detector_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        self._h = C.c_void_p()
        rc = lib.arslam_detector_create(device, max_images, max_width, max_height, C.byref(self._h))
        if rc != 0:
            raise ArslamError(rc, (lib.arslam_detector_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.arslam_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise ArslamError(rc, (self._lib.arslam_detector_last_error(self._h) or b"").decode())
        return rc

    def set_dictionary(self, bits, max_correction_bits):
        """bits: (n_markers, marker_size, marker_size) array of 0/1 (cv::aruco::Dictionary::getBitsFromByteList)."""
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        self._check(self._lib.arslam_detector_set_dictionary(self._h, b.shape[0], b.shape[1], int(max_correction_bits),
                                                             _p(b, C.c_uint8)))

    def set_predefined_dictionary(self, name):
        """"4X4_50" (default), "5X5_100" or "6X6_250" (aruco_detector.cpp:148-152)."""
        self._lib.arslam_detector_set_predefined_dictionary.argtypes = [C.c_void_p, C.c_char_p]
        self._check(self._lib.arslam_detector_set_predefined_dictionary(self._h, name.encode()))

    def detect(self, images, params=None, max_markers=256, device_ptr=None, shape=None):
        """images: (n, h, w, 3) BGR or (n, h, w) grey uint8 array (host), or device_ptr + shape for frames already
        in HBM.  Returns per frame (ids int32 array, corners (k, 4, 2) float32 array)."""
        if device_ptr is None:
            a = np.ascontiguousarray(images, dtype=np.uint8)
            shape = a.shape
            ptr, on_dev = a.ctypes.data, 0
        else:
            ptr, on_dev = int(device_ptr), 1
        n, h, w = shape[0], shape[1], shape[2]
        ch = shape[3] if len(shape) == 4 else 1
        n_found = np.zeros(n, np.int32)
        ids = np.zeros((n, max_markers), np.int32)
        corners = np.zeros((n, max_markers, 4, 2), np.float32)
        self._check(self._lib.arslam_detect_markers(self._h, C.c_void_p(ptr), n, w, h, ch, on_dev,
                                                    C.byref(params) if params is not None else None, max_markers,
                                                    _p(n_found, C.c_int32), _p(ids, C.c_int32), _p(corners, C.c_float)))
        return [(ids[i, :n_found[i]].copy(), corners[i, :n_found[i]].copy()) for i in range(n)]

    def candidates(self, cap=65536):
        """Candidate quads of the last detect() before grouping, in cv::aruco's order."""
        img = np.zeros(cap, np.int32)
        win = np.zeros(cap, np.int32)
        quad = np.zeros((cap, 4, 2), np.float32)
        near = np.zeros(cap, np.int32)
        mid = np.zeros(cap, np.int32)
        rot = np.zeros(cap, np.int32)
        n = self._check(self._lib.arslam_detector_candidates(self._h, cap, _p(img, C.c_int32), _p(win, C.c_int32),
                                                             _p(quad, C.c_float), _p(near, C.c_int32),
                                                             _p(mid, C.c_int32), _p(rot, C.c_int32)))
        n = min(n, cap)
        return dict(image=img[:n], window=win[:n], quad=quad[:n], near_border=near[:n], id=mid[:n], rotation=rot[:n])

    def read_stage(self, what):
        """0: grey frames, 1: threshold bits, 2: border table (n, 5) int32, 3: border points (x | y << 16) int32."""
        f = self._lib.arslam_detector_read_stage
        f.restype = C.c_int64
        f.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]
        need = self._check(f(self._h, what, None, 0))
        out = np.zeros(max(need, 1), np.uint8)
        self._check(f(self._h, what, C.c_void_p(out.ctypes.data), need))
        out = out[:need]
        return out if what < 2 else out.view(np.int32).reshape((-1, 5) if what == 2 else (-1,))

    def times(self):
        ms = (C.c_double * 6)()
        launches = C.c_int64()
        self._check(self._lib.arslam_detector_times(self._h, ms, C.byref(launches)))
        return dict(threshold_ms=ms[0], starts_ms=ms[1], follow_ms=ms[2], approx_ms=ms[3], identify_ms=ms[4],
                    total_ms=ms[5], launches=launches.value)
