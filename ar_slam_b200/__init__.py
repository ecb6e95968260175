"""ar_slam_b200 -- B200-native solver for ar_slam's optimisation hot path.

The product is the C-ABI library ``ar_slam_b200/lib/libar_slam_b200.so``
(sources in ``ar_slam_b200/csrc``, interface in ``include/ar_slam_b200.h``).
This package holds the build recipe, a ctypes mirror of that interface used by
the tests and bench.py, and the synthetic map generator.  There is no CPU
fallback: without the CUDA library, or without a GPU, every solver call raises.
"""
from .capi import (ELIM_AUTO, ELIM_CAPTURES, ELIM_TAGS, LINSOLVE_AUTO, LINSOLVE_DENSE, LINSOLVE_PCG,  # noqa: F401
                   ArslamError, Options, Solver, Summary, default_options, library_path, load_library)

__all__ = ["Solver", "Options", "Summary", "default_options", "load_library", "library_path", "ArslamError"]
