"""ctypes binding of ``libsatfill.so`` (``include/satfill.h``), the C-ABI of the B200 fill path.

This module holds no arithmetic: it loads the shared library, declares the prototypes exactly as the header does and
converts numpy / torch buffers into (pointer, strides) pairs.  If the library has not been built, or there is no CUDA
device, every entry point fails loudly -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SATFILL_LIB") or os.path.join(_HERE, "lib", "libsatfill.so")  # SATFILL_LIB: tuning builds
ABI_VERSION = 5  # SATFILL_ABI_VERSION of include/satfill.h

SA_OK, SA_EMPTY_MASK, SA_NOT_CONVERGED, SA_SIZE_MISMATCH, SA_BAD_ARGUMENT, SA_CUDA_ERROR, SA_NCCL_ERROR, SA_OOM = range(8)
SA_LAPLACE, SA_POISSON = 0, 1
SA_PRECOND_JACOBI, SA_PRECOND_MULTIGRID = 0, 1
SA_MG_RB32, SA_MG_JACOBI64, SA_MG_RB32_CTA = 0, 1, 2

STATUS_NAMES = {
    0: "SA_OK", 1: "SA_EMPTY_MASK", 2: "SA_NOT_CONVERGED", 3: "SA_SIZE_MISMATCH", 4: "SA_BAD_ARGUMENT",
    5: "SA_CUDA_ERROR", 6: "SA_NCCL_ERROR", 7: "SA_OUT_OF_MEMORY",
}  # fmt: skip

# every symbol include/satfill.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTS = [
    "sa_create", "sa_destroy", "sa_last_error", "sa_abi_version", "sa_default_options", "sa_kernel_launches",
    "sa_mask_scan", "sa_unknown_numbering", "sa_label_components", "sa_laplace_fill", "sa_poisson_blend",
    "sa_scene_create", "sa_scene_destroy", "sa_scene_set_mask", "sa_scene_set_band", "sa_scene_set_guidance",
    "sa_scene_solve", "sa_scene_get_band", "sa_scene_info", "sa_scene_precondition", "sa_synchronize",
    "sa_dist_unique_id", "sa_dist_init", "sa_dist_partition", "sa_dist_levels", "sa_scene_set_distributed",
    "sa_scene_owned_rows", "sa_scene_allgather_band", "sa_apply_laplace_u8", "sa_morph_close_mask", "sa_last_fill_direct",
    "sa_scene_plane_elements", "sa_has_legacy_variants", "sa_dist_uses_peer_memory",
]  # fmt: skip


class SatfillError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Options(C.Structure):
    _fields_ = [
        ("tolerance", C.c_double),
        ("max_iterations", C.c_int64),
        ("precond", C.c_int32),
        ("check_every", C.c_int32),
        ("mg_levels", C.c_int32),
        ("mg_smooth", C.c_int32),
        ("profile", C.c_int32),
        ("mg_unfused", C.c_int32),
        ("mg_variant", C.c_int32),
        ("cg_variant", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("unknowns", C.c_int64),
        ("iterations", C.c_int64),
        ("max_iterations", C.c_int64),
        ("tolerance", C.c_double),
        ("error", C.c_double),
        ("solve_ms", C.c_double),
        ("setup_ms", C.c_double),
        ("status", C.c_int32),
        ("active_tiles", C.c_int32),
        ("kernel_ms", C.c_double * 8),
        ("kernel_launches", C.c_int64 * 8),
        ("kernel_units", C.c_int64 * 8),
    ]

    def as_dict(self) -> dict:
        d = {name: getattr(self, name) for name, _ in self._fields_}
        d["kernel_ms"] = list(self.kernel_ms)
        d["kernel_launches"] = list(self.kernel_launches)
        d["kernel_units"] = list(self.kernel_units)
        return d


_lib = None
_vp = C.c_void_p
_i64 = C.c_int64


def load() -> C.CDLL:
    """Load libsatfill.so and declare its prototypes.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  satellite_approximation_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.sa_create.restype = C.c_int
    L.sa_create.argtypes = [C.POINTER(_vp), C.c_int, _vp]
    L.sa_destroy.restype = None
    L.sa_destroy.argtypes = [_vp]
    L.sa_last_error.restype = C.c_char_p
    L.sa_last_error.argtypes = [_vp]
    L.sa_abi_version.restype = C.c_int
    L.sa_abi_version.argtypes = []
    L.sa_default_options.restype = None
    L.sa_default_options.argtypes = [C.POINTER(Options), C.c_int]
    L.sa_kernel_launches.restype = _i64
    L.sa_kernel_launches.argtypes = [_vp]
    L.sa_mask_scan.restype = C.c_int
    L.sa_mask_scan.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, C.POINTER(_i64), C.POINTER(_i64)]
    L.sa_unknown_numbering.restype = C.c_int
    L.sa_unknown_numbering.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _vp, C.POINTER(_i64)]
    L.sa_label_components.restype = C.c_int
    L.sa_label_components.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _vp, C.POINTER(C.c_int32)]
    L.sa_laplace_fill.restype = C.c_int
    L.sa_laplace_fill.argtypes = [_vp, _vp, C.c_int, _vp, _i64, _i64, _i64, _i64, C.POINTER(Options), C.POINTER(Stats)]
    L.sa_poisson_blend.restype = C.c_int
    L.sa_poisson_blend.argtypes = [_vp, _vp, _vp, C.c_int, _vp, _i64, _i64, _i64, _i64, C.POINTER(Options),
                                   C.POINTER(Stats)]  # fmt: skip
    L.sa_scene_create.restype = C.c_int
    L.sa_scene_create.argtypes = [_vp, C.c_int, _i64, _i64, C.c_int, C.POINTER(_vp)]
    L.sa_scene_destroy.restype = None
    L.sa_scene_destroy.argtypes = [_vp]
    L.sa_scene_set_mask.restype = C.c_int
    L.sa_scene_set_mask.argtypes = [_vp, _vp, _i64, _i64, C.c_int]
    L.sa_scene_set_band.restype = C.c_int
    L.sa_scene_set_band.argtypes = [_vp, C.c_int, _vp, _i64, _i64, C.c_int]
    L.sa_scene_set_guidance.restype = C.c_int
    L.sa_scene_set_guidance.argtypes = [_vp, C.c_int, _vp, _i64, _i64, C.c_int]
    L.sa_scene_solve.restype = C.c_int
    L.sa_scene_solve.argtypes = [_vp, C.POINTER(Options), C.POINTER(Stats)]
    L.sa_scene_get_band.restype = C.c_int
    L.sa_scene_get_band.argtypes = [_vp, C.c_int, _vp, _i64, _i64, C.c_int]
    L.sa_scene_info.restype = C.c_int
    L.sa_scene_info.argtypes = [_vp, C.POINTER(_i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sa_scene_precondition.restype = C.c_int
    L.sa_scene_precondition.argtypes = [_vp, C.POINTER(Options), _vp, _vp, _i64, _i64]
    L.sa_dist_unique_id.restype = C.c_int
    L.sa_dist_unique_id.argtypes = [_vp]
    L.sa_dist_init.restype = C.c_int
    L.sa_dist_init.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.sa_dist_partition.restype = C.c_int
    L.sa_dist_partition.argtypes = [_i64, C.c_int, C.c_int, C.POINTER(_i64)]
    L.sa_dist_levels.restype = C.c_int
    L.sa_dist_levels.argtypes = [_i64, C.c_int]
    L.sa_scene_set_distributed.restype = C.c_int
    L.sa_scene_set_distributed.argtypes = [_vp, C.c_int]
    L.sa_scene_owned_rows.restype = C.c_int
    L.sa_scene_owned_rows.argtypes = [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(C.c_int)]
    L.sa_scene_allgather_band.restype = C.c_int
    L.sa_scene_allgather_band.argtypes = [_vp, C.c_int]
    L.sa_apply_laplace_u8.restype = C.c_int
    L.sa_apply_laplace_u8.argtypes = [_vp, _vp, _vp, _i64, _i64, C.c_int, C.c_double, _vp, _vp, C.POINTER(Options), _vp]
    L.sa_morph_close_mask.restype = C.c_int
    L.sa_morph_close_mask.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, C.c_int, _vp]
    L.sa_last_fill_direct.restype = C.c_int
    L.sa_last_fill_direct.argtypes = [_vp]
    L.sa_scene_plane_elements.restype = _i64
    L.sa_scene_plane_elements.argtypes = [_i64, _i64]
    L.sa_dist_uses_peer_memory.restype = C.c_int
    L.sa_dist_uses_peer_memory.argtypes = [_vp]
    L.sa_has_legacy_variants.restype = C.c_int
    L.sa_has_legacy_variants.argtypes = []
    L.sa_synchronize.restype = C.c_int
    L.sa_synchronize.argtypes = [_vp]
    if L.sa_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {L.sa_abi_version()} != {ABI_VERSION}")
    _lib = L
    return L


def element_strides(a: np.ndarray) -> tuple[int, int]:
    if a.ndim != 2:
        raise ValueError("expected a 2-D array")
    if a.strides[0] % a.itemsize or a.strides[1] % a.itemsize:
        raise ValueError("array strides are not a multiple of the item size")
    return a.strides[0] // a.itemsize, a.strides[1] // a.itemsize


def is_dense_2d(a: np.ndarray) -> bool:
    """True when the C-ABI can address the array directly (one unit stride, the other non-negative and large enough)."""
    if a.ndim != 2:
        return False
    try:
        rs, cs = element_strides(a)
    except ValueError:
        return False
    rows, cols = a.shape
    row_major = (cs == 1 or cols <= 1) and (rs >= cols or rows <= 1)
    col_major = (rs == 1 or rows <= 1) and (cs >= rows or cols <= 1)
    return row_major or col_major
