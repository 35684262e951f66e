"""Drop-in for the reference's Python package `satellite_approximation` (src/satellite_approximation/__init__.py) on
the fill path.  Uses the compiled pybind11 module `_core` (cpp/src/pybind_module.cpp, built by `make -C cpp pybind`,
needs Eigen headers) when it is present, else the ctypes mirror in satellite_approximation_b200 -- both sit on the same
C-ABI (libsatfill.so) and neither has a CPU fallback.  Cloud-detection symbols of the reference are out of scope."""
from __future__ import annotations

try:
    from ._core import LogLevel, Path, blend_images_poisson, filling_missing_portions_smooth_boundaries, set_log_level

    BACKEND = "pybind11"
except ImportError:
    from satellite_approximation_b200 import (
        LogLevel,
        Path,
        blend_images_poisson,
        filling_missing_portions_smooth_boundaries,
        set_log_level,
    )

    BACKEND = "ctypes"

__all__ = ["LogLevel", "Path", "set_log_level", "filling_missing_portions_smooth_boundaries", "blend_images_poisson"]


def __getattr__(name):
    if name in ("CloudParams", "SkipShadowDetection", "get_diagonal_distance", "detect"):
        raise NotImplementedError(f"satellite_approximation.{name}: cloud/shadow detection is outside the B200 fill path")
    raise AttributeError(name)
