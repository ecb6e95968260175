"""The C-ABI library loads on a CPU-only box, exports every symbol the header declares, agrees
with the ctypes mirror on struct layouts, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    with open(os.path.join(ROOT, "include", "ar_slam_b200.h")) as f:
        return f.read()


def test_library_builds_and_exports_header_symbols():
    from ar_slam_b200 import build, capi
    build.build()
    lib = capi.load_library()
    names = set(re.findall(r"\b(arslam_[a-z_0-9]+)\s*\(", _header()))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), "libar_slam_b200.so does not export %s" % n
    assert lib.arslam_abi_version() == int(re.search(r"#define ARSLAM_ABI_VERSION (\d+)", _header()).group(1))


def test_default_options_are_the_reference_settings():
    import ar_slam_b200 as ar
    o = ar.default_options()
    # ar_slam_util.cpp:1003-1012 + Ceres 2.0 defaults
    assert o.max_num_iterations == 50 and o.initial_trust_region_radius == 1e4
    assert o.function_tolerance == 1e-6 and o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.min_relative_decrease == 1e-3 and o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert o.jacobi_scaling == 1 and o.tag_size == 0.0635 and o.num_intrinsics == 1
    assert o.max_num_consecutive_invalid_steps == 5


def test_struct_layout_matches_header_field_order():
    from ar_slam_b200 import capi
    hdr = _header()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1]
            out += [n.strip().split("[")[0] for n in names.split(",")]
        return out
    assert fields("arslam_options") == [n for n, _ in capi.Options._fields_]
    assert fields("arslam_summary") == [n for n, _ in capi.Summary._fields_]
    assert fields("arslam_kernel_time") == [n for n, _ in capi.KernelTime._fields_]
    assert fields("arslam_detect_params") == [n for n, _ in capi.DetectParams._fields_]


def test_detector_defaults_are_opencvs():
    """cv::aruco::DetectorParameters defaults (the reference only overrides minCornerDistanceRate, ar_slam_util.cpp:250)."""
    from ar_slam_b200 import capi
    p = capi.default_detect_params()
    assert (p.adaptive_thresh_win_size_min, p.adaptive_thresh_win_size_max, p.adaptive_thresh_win_size_step) == (3, 23, 10)
    assert p.adaptive_thresh_constant == 7.0 and p.min_distance_to_border == 3 and p.marker_border_bits == 1
    assert (p.min_marker_perimeter_rate, p.max_marker_perimeter_rate, p.polygonal_approx_accuracy_rate) == (0.03, 4.0, 0.03)
    assert p.min_corner_distance_rate == 0.05 and p.min_marker_distance_rate == 0.125
    assert p.perspective_remove_pixel_per_cell == 4 and p.perspective_remove_ignored_margin_per_cell == 0.13
    assert p.max_erroneous_bits_in_border_rate == 0.35 and p.min_otsu_std_dev == 5.0 and p.error_correction_rate == 0.6
    try:
        import cv2
    except ImportError:
        return
    q = cv2.aruco.DetectorParameters()
    assert p.min_group_distance == q.minGroupDistance and p.min_marker_distance_rate == q.minMarkerDistanceRate
    assert p.min_corner_distance_rate == q.minCornerDistanceRate and p.min_otsu_std_dev == q.minOtsuStdDev


def test_no_cpu_fallback():
    import torch
    import ar_slam_b200 as ar
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ar.ArslamError) as e:
        ar.Solver()
    assert e.value.code == ar.capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "ar_slam_b200")):
        if os.sep + "lib" in base:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(base, fn)) as f:
                    assert not pat.search(f.read()), "%s refers to the oracle" % fn


def test_synthetic_generator_is_deterministic_and_sane():
    from ar_slam_b200 import synth
    a = synth.make_map(300, 80, seed=5)
    b = synth.make_map(300, 80, seed=5)
    assert np.array_equal(a.obs, b.obs) and np.array_equal(a.tag_idx, b.tag_idx)
    assert len(a.cap_idx) == 8 * 300 and np.all(np.bincount(a.cap_idx) == 8)
    assert np.all(np.diff(a.cap_idx) >= 0)                       # blocks of a capture are contiguous
    assert np.array_equal(a.obs, a.obs.astype(np.float32).astype(np.float64))   # Point32 round trip
    uv, z = synth.project(a.cam_true, a.cap_true[a.cap_idx], a.tag_true[a.tag_idx])
    assert (z > 0).all() and np.abs(uv[..., 0]).max() < 510 and np.abs(uv[..., 1]).max() < 384
    assert np.abs(uv.reshape(-1, 8) - a.obs).std() < 0.35          # 0.3 px noise
