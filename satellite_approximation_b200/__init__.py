"""satellite_approximation_b200 -- the Laplace / Poisson fill path of ebiederstadt/satellite-approximation on B200.

Host-side mirror of the reference's Python module ``satellite_approximation`` (``src/main.cpp:49-58``,
``src/satellite_approximation/__init__.py``) for the fill path, plus the declared-but-undefined
``find_connected_components`` (``lib/approx/include/approx/laplace.h:11-20``).  Same function names, argument names,
defaults, dtype rules and error behaviour; the arithmetic runs in ``lib/libsatfill.so`` (hand-written CUDA for
sm_100a) through the C-ABI in ``include/satfill.h``.  There is no CPU fallback: importing works without a GPU, calling
does not.

    from satellite_approximation_b200 import filling_missing_portions_smooth_boundaries, blend_images_poisson
    filled = filling_missing_portions_smooth_boundaries(img_f64, mask_bool)
    bands  = blend_images_poisson([f0, f1], [g0, g1], mask_bool, tolerance=1e-6)

Device-resident use (no PCIe traffic inside the solve; what ``bench.py`` times as ``value``)::

    ctx = Context(device=0)
    scene = ctx.scene(LAPLACE, rows, cols, nbands)
    scene.set_mask(mask); scene.set_band(0, img); stats = scene.solve(tolerance=1e-6); out = scene.get_band(0)
"""
from __future__ import annotations

import ctypes as C
import enum
import logging
import threading
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import (  # noqa: F401  (re-exported)
    SA_BAD_ARGUMENT,
    SA_EMPTY_MASK,
    SA_LAPLACE as LAPLACE,
    SA_NOT_CONVERGED,
    SA_OK,
    SA_POISSON as POISSON,
    SA_PRECOND_JACOBI as JACOBI,
    SA_PRECOND_MULTIGRID as MULTIGRID,
    SA_MG_RB32 as MG_RB32,
    SA_MG_JACOBI64 as MG_JACOBI64,
    SA_MG_RB32_CTA as MG_RB32_CTA,
    SatfillError,
)

__all__ = [
    "LogLevel", "Path", "set_log_level", "filling_missing_portions_smooth_boundaries", "blend_images_poisson",
    "find_connected_components", "ConnectedComponents", "mask_scan", "unknown_numbering", "valid_neighbours",
    "Context", "Scene", "SolveStats", "default_context", "set_solver_defaults", "last_perf_info", "write_perf_info",
    "dist_partition", "dist_levels", "apply_laplace", "preprocess_cloud_band", "blend_images_poisson_offset", "valid_pixel_mask", "image_to_channels", "channels_to_image",
    "highlight_area_replaced", "LAPLACE", "POISSON", "JACOBI", "MULTIGRID", "MG_RB32", "MG_JACOBI64", "MG_RB32_CTA", "SatfillError",
]  # fmt: skip

_log = logging.getLogger("satellite_approximation_b200")


class LogLevel(enum.IntEnum):
    """spdlog levels exported by the reference module (src/main.cpp:24-29)."""

    Debug = 1
    Info = 2
    Warn = 3
    Error = 4
    Critical = 5


_PY_LEVEL = {
    LogLevel.Debug: logging.DEBUG, LogLevel.Info: logging.INFO, LogLevel.Warn: logging.WARNING,
    LogLevel.Error: logging.ERROR, LogLevel.Critical: logging.CRITICAL,
}  # fmt: skip


class Path:
    """`Path` of the reference's module (src/main.cpp:20-22): std::filesystem::path constructible from a str."""

    def __init__(self, path: str):
        if not isinstance(path, str):
            raise TypeError("Path(): expected a str")
        self._p = path

    def __fspath__(self) -> str:
        return self._p

    def __str__(self) -> str:
        return self._p

    def __repr__(self) -> str:
        return f"Path({self._p!r})"


def set_log_level(level: LogLevel) -> None:
    """src/main.cpp:30-34.  Unlike the reference nothing is created on disk at import time."""
    _log.setLevel(_PY_LEVEL[LogLevel(level)])
    _log.info("Logging set to level: %s", LogLevel(level).name)


class SolveStats(dict):
    """Per-band solve record: superset of approx::PerfInfo (poisson.h:12-21)."""

    __getattr__ = dict.__getitem__


# Solver knobs the reference does not expose on its Python surface.  `None` = the reference's own behaviour
# (Laplace: Eigen defaults, epsilon tolerance / 2N iterations, laplace.cpp:113-114).
_defaults = {"laplace_tolerance": None, "laplace_max_iterations": None, "precond": MULTIGRID, "check_every": None}
_last_perf: list[SolveStats] = []


def set_solver_defaults(**kw) -> None:
    """laplace_tolerance, laplace_max_iterations, precond (JACOBI | MULTIGRID), check_every."""
    for k, v in kw.items():
        if k not in _defaults:
            raise TypeError(f"unknown solver default {k!r}")
        _defaults[k] = v


def last_perf_info() -> list[SolveStats]:
    """Records of the most recent fill (the reference appends PerfInfo to a hard-coded CSV, poisson.cpp:287-289)."""
    return list(_last_perf)


def write_perf_info(output, records: Optional[Sequence[SolveStats]] = None) -> None:
    """PerfInfo::write (poisson.cpp:13-18): append `region_size,tolerance,max_iterations,iterations,error,solve_time`
    (seconds) to the CSV `output`, one line per record (default: the records of the most recent fill).  Opt-in: the
    reference appends after every mask-overload blend to a path hard-coded to its author's home (poisson.cpp:285-289)."""
    import os

    rows = list(_last_perf if records is None else records)
    with open(os.fspath(output), "a") as f:
        for r in rows:
            f.write(f"{int(r['unknowns'])},{r['tolerance']:g},{int(r['max_iterations'])},{int(r['iterations'])},"
                    f"{r['error']:g},{r['solve_ms'] * 1e-3:g}\n")  # fmt: skip


def _ptr(a) -> int:
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a.data_ptr())


def _describe(a, dtype_np, what: str):
    """(pointer, rows, cols, row_stride, col_stride, on_device) of a numpy array or a CUDA torch tensor."""
    if isinstance(a, np.ndarray):
        if a.dtype != dtype_np:
            raise TypeError(f"{what}: expected dtype {np.dtype(dtype_np)}, got {a.dtype}")
        rs, cs = _capi.element_strides(a)
        return a.ctypes.data, a.shape[0], a.shape[1], rs, cs, 0
    import torch  # device buffers are torch tensors: torch is the plumbing for device memory and streams

    if not isinstance(a, torch.Tensor) or a.dim() != 2:
        raise TypeError(f"{what}: expected a 2-D numpy array or torch tensor")
    want = (torch.float64,) if np.dtype(dtype_np) == np.float64 else (torch.uint8, torch.bool)
    if a.dtype not in want:
        raise TypeError(f"{what}: expected {want}, got {a.dtype}")
    return a.data_ptr(), a.shape[0], a.shape[1], a.stride(0), a.stride(1), 1 if a.is_cuda else 0


class Context:
    """One sa_ctx: a device, a stream and the scene cache of the host-pointer entry points."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._lib = _capi.load()
        h = C.c_void_p()
        st = self._lib.sa_create(C.byref(h), int(device), C.c_void_p(stream) if stream else None)
        if st != SA_OK:
            raise SatfillError(st, f"sa_create(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self._h = h
        self.device = int(device)
        self._lock = threading.Lock()

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.sa_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int, ok=(SA_OK,)):
        if st not in ok:
            raise SatfillError(st, self._lib.sa_last_error(self._h).decode())
        return st

    @property
    def dist_uses_peer_memory(self) -> bool:
        """True when the exchanges of a row-decomposed solve go over peer memory (CUDA IPC) rather than NCCL."""
        return bool(self._lib.sa_dist_uses_peer_memory(self._h))

    @property
    def has_legacy_variants(self) -> bool:
        """True when the loaded library also holds the first-generation kernels (cg_variant = 1, MG_JACOBI64, MG_RB32_CTA):
        lib/libsatfill_legacy.so, built with SATFILL_LEGACY_VARIANTS.  The product library refuses those variants."""
        return bool(self._lib.sa_has_legacy_variants())

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.sa_kernel_launches(self._h))

    @property
    def last_fill_direct(self) -> bool:
        """True if the last laplace_fill / poisson_blend read and wrote the caller's page-locked arrays in place over PCIe
        (no image copies): sa_last_fill_direct."""
        return bool(self._lib.sa_last_fill_direct(self._h))

    def synchronize(self) -> None:
        self._check(self._lib.sa_synchronize(self._h))

    # ---- one system across several GPUs (sa_dist_*) --------------------------------------------------------------
    def dist_init_torch(self) -> None:
        """Join the ranks of the initialised torch.distributed process group into the library's own NCCL communicator:
        rank 0 creates the id, torch.distributed broadcasts its 128 bytes (the only thing the host language moves)."""
        import torch
        import torch.distributed as dist

        from . import multi

        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            self._check(self._lib.sa_dist_unique_id(buf))
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        self.dist_init(multi.broadcast_bytes(bytes(buf) if rank == 0 else None, 128, 0, dev), rank, world)

    def dist_init(self, id128: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self._check(self._lib.sa_dist_init(self._h, buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def options(self, problem: int, tolerance=None, max_iterations=None, precond=None, check_every=None,
                mg_levels=None, mg_smooth=None, profile=None, mg_unfused=None, mg_variant=None, cg_variant=None) -> _capi.Options:  # fmt: skip
        o = _capi.Options()
        self._lib.sa_default_options(C.byref(o), problem)
        if tolerance is not None:
            o.tolerance = float(tolerance)
        if max_iterations is not None:
            o.max_iterations = int(max_iterations)
        if precond is not None:
            o.precond = int(precond)
        if check_every is not None:
            o.check_every = int(check_every)
        if mg_levels is not None:
            o.mg_levels = int(mg_levels)
        if mg_smooth is not None:
            o.mg_smooth = int(mg_smooth)
        if profile is not None:
            o.profile = int(bool(profile))
        if mg_unfused is not None:
            o.mg_unfused = int(bool(mg_unfused))
        if mg_variant is not None:
            o.mg_variant = int(mg_variant)
        if cg_variant is not None:
            o.cg_variant = int(cg_variant)
        return o

    # ---- integer path -----------------------------------------------------------------------------------------
    def mask_scan(self, mask: np.ndarray):
        """laplace.cpp:33-52: (pixels[n, 2] int64 in row-major raster order, bbox = [min_row, max_row, min_col, max_col])."""
        m = _mask_u8(mask)
        rs, cs = _capi.element_strides(m)
        n = C.c_int64()
        bbox = (C.c_int64 * 4)()
        self._check(self._lib.sa_mask_scan(self._h, m.ctypes.data, m.shape[0], m.shape[1], rs, cs, None, 0,
                                           C.byref(n), bbox))  # fmt: skip
        px = np.empty((n.value, 2), np.int64)
        if n.value:
            self._check(self._lib.sa_mask_scan(self._h, m.ctypes.data, m.shape[0], m.shape[1], rs, cs, px.ctypes.data,
                                               n.value, C.byref(n), bbox))  # fmt: skip
        return px, np.array(list(bbox), np.int64)

    def unknown_numbering(self, mask: np.ndarray):
        """poisson.cpp:162-177: (numbering[rows, cols] int32 with -1 at valid pixels, n)."""
        m = _mask_u8(mask)
        rs, cs = _capi.element_strides(m)
        num = np.empty(m.shape, np.int32)
        n = C.c_int64()
        self._check(self._lib.sa_unknown_numbering(self._h, m.ctypes.data, m.shape[0], m.shape[1], rs, cs,
                                                   num.ctypes.data, C.byref(n)))  # fmt: skip
        return num, int(n.value)

    def label_components(self, mask: np.ndarray):
        """laplace.h:11-20 contract: (labels[rows, cols] int32, K)."""
        m = _mask_u8(mask)
        rs, cs = _capi.element_strides(m)
        lab = np.empty(m.shape, np.int32)
        k = C.c_int32()
        self._check(self._lib.sa_label_components(self._h, m.ctypes.data, m.shape[0], m.shape[1], rs, cs,
                                                  lab.ctypes.data, C.byref(k)))  # fmt: skip
        return lab, int(k.value)

    # ---- float path, host buffers -------------------------------------------------------------------------------
    def laplace_fill(self, images: Sequence[np.ndarray], mask: np.ndarray, **opts):
        """In place on `images` (float64, all one layout, same shape as mask).  Returns per-band SolveStats."""
        m = _mask_u8(mask)
        rows, cols = m.shape
        rs, cs = _capi.element_strides(m)
        for a in images:
            if a.dtype != np.float64 or a.shape != m.shape or _capi.element_strides(a) != (rs, cs):
                raise ValueError("laplace_fill: images must be float64 with the shape and memory layout of the mask")
        nb = len(images)
        ptrs = (C.c_void_p * nb)(*[a.ctypes.data for a in images])
        stats = (_capi.Stats * nb)()
        o = self.options(LAPLACE, **opts)
        with self._lock:
            st = self._lib.sa_laplace_fill(self._h, ptrs, nb, m.ctypes.data, rows, cols, rs, cs, C.byref(o), stats)
        self._check(st, ok=(SA_OK, SA_EMPTY_MASK, SA_NOT_CONVERGED))
        return [SolveStats(s.as_dict()) for s in stats]

    def poisson_blend(self, inputs: Sequence[np.ndarray], replacements: Sequence[np.ndarray], mask: np.ndarray,
                      **opts):  # fmt: skip
        m = _mask_u8(mask)
        rows, cols = m.shape
        rs, cs = _capi.element_strides(m)
        for a in list(inputs) + list(replacements):
            if a.dtype != np.float64 or a.shape != m.shape or _capi.element_strides(a) != (rs, cs):
                raise ValueError("poisson_blend: images must be float64 with the shape and memory layout of the mask")
        nb = len(inputs)
        pin = (C.c_void_p * nb)(*[a.ctypes.data for a in inputs])
        prp = (C.c_void_p * nb)(*[a.ctypes.data for a in replacements])
        stats = (_capi.Stats * nb)()
        o = self.options(POISSON, **opts)
        with self._lock:
            st = self._lib.sa_poisson_blend(self._h, pin, prp, nb, m.ctypes.data, rows, cols, rs, cs, C.byref(o), stats)
        self._check(st, ok=(SA_OK, SA_EMPTY_MASK, SA_NOT_CONVERGED))
        return [SolveStats(s.as_dict()) for s in stats]

    def apply_laplace(self, image: np.ndarray, invalid_image: np.ndarray, red_threshold: float = 220.0, **opts):
        """approx::apply_laplace (laplace.cpp:134-168): `image`, `invalid_image` uint8 H x W x 3 in cv::imread order
        (B, G, R).  Returns (filled float64 H x W x 3, bool mask H x W, per-channel SolveStats)."""
        if image.dtype != np.uint8 or invalid_image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise TypeError("apply_laplace: uint8 H x W x 3 images (cv::imread(IMREAD_COLOR))")
        if image.shape != invalid_image.shape:
            raise RuntimeError("Input image and mask are not the same size")  # laplace.cpp:124-127
        img = np.ascontiguousarray(image)
        inv = np.ascontiguousarray(invalid_image)
        rows, cols, ch = img.shape
        out = np.empty((rows, cols, ch), np.float64)
        mask = np.empty((rows, cols), np.uint8)
        stats = (_capi.Stats * ch)()
        o = self.options(LAPLACE, **opts)
        with self._lock:
            st = self._lib.sa_apply_laplace_u8(self._h, img.ctypes.data, inv.ctypes.data, rows, cols, ch, float(red_threshold),
                                               out.ctypes.data, mask.ctypes.data, C.byref(o), stats)  # fmt: skip
        self._check(st, ok=(SA_OK, SA_EMPTY_MASK, SA_NOT_CONVERGED))
        if st == SA_EMPTY_MASK:
            out[...] = img
            mask[...] = 0
        return out, mask.astype(bool), [SolveStats(s.as_dict()) for s in stats]

    def morph_close_mask(self, band: np.ndarray, radius: int = 5) -> np.ndarray:
        """preprocess_cloud_band (poisson-main.cpp:10-21): MORPH_CLOSE with a (2 radius + 1)^2 rectangle, cast to bool."""
        if band.dtype != np.float64 or band.ndim != 2:
            raise TypeError("morph_close_mask: a 2-D float64 band")
        # The entry point writes the mask with the band's own pitch (satfill.h): a strided view (band[:, :k] of a wider
        # array, negative strides ...) is densified first, so that the dense mask allocated here is what the library
        # addresses -- a view's pitch would run past the end of it.
        if not (band.flags.c_contiguous or band.flags.f_contiguous):
            band = np.ascontiguousarray(band)
        rs, cs = _capi.element_strides(band)
        rows, cols = band.shape
        mask = np.empty(band.shape, np.uint8, order="F" if (band.flags.f_contiguous and not band.flags.c_contiguous) else "C")
        mrs, mcs = _capi.element_strides(mask)
        if rows * cols and (rows > 1 and cols > 1) and (mrs, mcs) != (rs, cs):
            raise AssertionError("morph_close_mask: mask and band layouts differ")
        with self._lock:
            st = self._lib.sa_morph_close_mask(self._h, band.ctypes.data, rows, cols, rs, cs, int(radius), mask.ctypes.data)
        self._check(st)
        return mask.astype(bool)

    def scene(self, problem: int, rows: int, cols: int, nbands: int = 1) -> "Scene":
        return Scene(self, problem, rows, cols, nbands)


class Scene:
    """A mask + nbands images (+ guidance) resident in HBM (sa_scene_*)."""

    def __init__(self, ctx: Context, problem: int, rows: int, cols: int, nbands: int):
        self.ctx, self.problem, self.rows, self.cols, self.nbands = ctx, problem, rows, cols, nbands
        h = C.c_void_p()
        ctx._check(ctx._lib.sa_scene_create(ctx._h, problem, rows, cols, nbands, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.sa_scene_destroy(self._h)
        self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _shape_ok(self, rows, cols):
        if (rows, cols) != (self.rows, self.cols):
            raise ValueError(f"expected shape {(self.rows, self.cols)}, got {(rows, cols)}")

    def set_mask(self, mask) -> None:
        if isinstance(mask, np.ndarray):
            mask = _mask_u8(mask)
        p, r, c, rs, cs, dev = _describe(mask, np.uint8, "mask")
        self._shape_ok(r, c)
        self.ctx._check(self.ctx._lib.sa_scene_set_mask(self._h, p, rs, cs, dev))

    def set_band(self, band: int, image) -> None:
        p, r, c, rs, cs, dev = _describe(image, np.float64, "image")
        self._shape_ok(r, c)
        self.ctx._check(self.ctx._lib.sa_scene_set_band(self._h, band, p, rs, cs, dev))

    def set_guidance(self, band: int, image) -> None:
        p, r, c, rs, cs, dev = _describe(image, np.float64, "guidance")
        self._shape_ok(r, c)
        self.ctx._check(self.ctx._lib.sa_scene_set_guidance(self._h, band, p, rs, cs, dev))

    def solve(self, raise_on_failure: bool = False, **opts) -> list[SolveStats]:
        stats = (_capi.Stats * self.nbands)()
        o = self.ctx.options(self.problem, **opts)
        st = self.ctx._lib.sa_scene_solve(self._h, C.byref(o), stats)
        self.ctx._check(st, ok=(SA_OK,) if raise_on_failure else (SA_OK, SA_EMPTY_MASK, SA_NOT_CONVERGED))
        return [SolveStats(s.as_dict()) for s in stats]

    def get_band(self, band: int, out=None, order: str = "C"):
        if out is None:
            out = np.empty((self.rows, self.cols), np.float64, order=order)
        p, r, c, rs, cs, dev = _describe(out, np.float64, "out")
        self._shape_ok(r, c)
        self.ctx._check(self.ctx._lib.sa_scene_get_band(self._h, band, p, rs, cs, dev))
        return out

    def precondition(self, r: np.ndarray, **opts) -> np.ndarray:
        """z = M^-1 r: one application of the multigrid preconditioner (diagnostic hook of the parity tests)."""
        r = np.ascontiguousarray(r, np.float64)
        self._shape_ok(*r.shape)
        z = np.empty_like(r)
        o = self.ctx.options(self.problem, **opts)
        self.ctx._check(self.ctx._lib.sa_scene_precondition(self._h, C.byref(o), r.ctypes.data, z.ctypes.data, r.shape[1], 1))
        return z

    def set_distributed(self, on: bool = True) -> None:
        """One system shared by all ranks of the context's communicator, split by rows (sa_scene_set_distributed)."""
        self.ctx._check(self.ctx._lib.sa_scene_set_distributed(self._h, int(bool(on))))

    def owned_rows(self):
        """(lo, hi, axis): the slice of the caller's array this rank holds the solution for after a distributed solve."""
        lo, hi, axis = C.c_int64(), C.c_int64(), C.c_int()
        self.ctx._check(self.ctx._lib.sa_scene_owned_rows(self._h, C.byref(lo), C.byref(hi), C.byref(axis)))
        return lo.value, hi.value, axis.value

    def allgather_band(self, band: int) -> None:
        self.ctx._check(self.ctx._lib.sa_scene_allgather_band(self._h, int(band)))

    def info(self) -> dict:
        n, a, t = C.c_int64(), C.c_int32(), C.c_int32()
        self.ctx._lib.sa_scene_info(self._h, C.byref(n), C.byref(a), C.byref(t))
        return {"unknowns": n.value, "active_tiles": a.value, "total_tiles": t.value}


def dist_partition(rows: int, world: int, levels: int) -> list[int]:
    """Row boundaries (world + 1) of the row decomposition the distributed solver uses (sa_dist_partition)."""
    lib = _capi.load()
    out = (C.c_int64 * (world + 1))()
    if lib.sa_dist_partition(int(rows), int(world), int(levels), out) != SA_OK:
        raise ValueError("dist_partition: bad arguments")
    return list(out)


def dist_levels(rows: int, world: int) -> int:
    """Number of multigrid levels the distributed solver splits by rows for a scene of `rows` rows (sa_dist_levels)."""
    return int(_capi.load().sa_dist_levels(int(rows), int(world)))


def apply_laplace(image: np.ndarray, invalid_image: np.ndarray, red_threshold: float = 220.0) -> np.ndarray:
    """approx::apply_laplace (lib/approx/include/approx/laplace.h:31, laplace.cpp:134-168): the body of `laplace_main`.
    uint8 H x W x 3 images in cv::imread order; returns the float64 H x W x 3 matrix the reference returns."""
    opts = {}
    if _defaults.get("laplace_tolerance") is not None:
        opts["tolerance"] = _defaults["laplace_tolerance"]
    if _defaults.get("precond") is not None:
        opts["precond"] = _defaults["precond"]
    return default_context().apply_laplace(image, invalid_image, red_threshold, **opts)[0]


def preprocess_cloud_band(cloud_band: np.ndarray, dilation_size: int = 5) -> np.ndarray:
    """preprocess_cloud_band of poisson_main (executables/poisson-main.cpp:10-21): 11 x 11 morphological close of the
    cloud band, cast to bool -- the mask poisson_main hands to blend_images_poisson."""
    return default_context().morph_close_mask(np.asarray(cloud_band), dilation_size)


_default_ctx: Optional[Context] = None
_ctx_lock = threading.Lock()


def default_context() -> Context:
    """The process-wide context the module-level functions use (device 0 unless LOCAL_RANK says otherwise)."""
    global _default_ctx
    with _ctx_lock:
        if _default_ctx is None:
            import os

            _default_ctx = Context(int(os.environ.get("SATFILL_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
        return _default_ctx


def _mask_u8(mask: np.ndarray) -> np.ndarray:
    if not isinstance(mask, np.ndarray) or mask.ndim != 2:
        raise TypeError("mask: expected a 2-D numpy array")
    if mask.dtype == np.bool_:
        mask = mask.view(np.uint8)
    elif mask.dtype != np.uint8:
        raise TypeError(f"mask: expected bool or uint8, got {mask.dtype}")
    if not _capi.is_dense_2d(mask):
        mask = np.ascontiguousarray(mask)
    return mask


# ---- the reference's Python surface ---------------------------------------------------------------------------------


def filling_missing_portions_smooth_boundaries(input_image: np.ndarray, invalid_pixels: np.ndarray) -> np.ndarray:
    """Laplace (harmonic) fill of the invalid pixels -- src/main.cpp:49-54 -> laplace.cpp:122-132.

    Both arguments are ``noconvert`` in the reference: ``input_image`` must be a 2-D float64 array and
    ``invalid_pixels`` a 2-D bool array (any strides), otherwise TypeError.  Returns a NEW Fortran-ordered float64
    array; the argument is not modified.  Raises RuntimeError when the element counts differ (laplace.cpp:124-127).
    """
    if not (isinstance(input_image, np.ndarray) and input_image.ndim == 2 and input_image.dtype == np.float64):
        raise TypeError("filling_missing_portions_smooth_boundaries(): input_image must be a 2-D float64 numpy array")
    if not (isinstance(invalid_pixels, np.ndarray) and invalid_pixels.ndim == 2 and invalid_pixels.dtype == np.bool_):
        raise TypeError("filling_missing_portions_smooth_boundaries(): invalid_pixels must be a 2-D bool numpy array")
    if input_image.size != invalid_pixels.size:
        raise RuntimeError("Input image and mask need to be the same size")  # laplace.cpp:124-127
    out = np.array(input_image, dtype=np.float64, order="F", copy=True)  # the pybind11 caster copies (MatX is col-major)
    if invalid_pixels.shape != out.shape:
        # same element count, different shape: the reference indexes the mask with the image's (row, col)
        raise RuntimeError("Input image and mask need to be the same shape")
    mask = np.asfortranarray(invalid_pixels)
    global _last_perf
    _last_perf = default_context().laplace_fill(
        [out], mask, tolerance=_defaults["laplace_tolerance"], max_iterations=_defaults["laplace_max_iterations"],
        precond=_defaults["precond"], check_every=_defaults["check_every"],
    )  # fmt: skip
    if _last_perf and _last_perf[0]["status"] == SA_EMPTY_MASK:
        _log.info("No invalid pixels: nothing to do")  # laplace.cpp:41-44
    return out


def blend_images_poisson(input_image: Sequence[np.ndarray], replacement_image: Sequence[np.ndarray],
                         invalid_mask: np.ndarray, tolerance: float = 1e-6,
                         max_iterations: Optional[int] = None) -> list[np.ndarray]:  # fmt: skip
    """Poisson (seamless-cloning) blend, mask overload -- src/main.cpp:55-58 -> poisson.cpp:292-303, 145-290.

    Returns a list of new Fortran-ordered float64 arrays.  Like the reference, a size mismatch or a band that does not
    converge within ``max_iterations`` (default: unknowns / 2) is logged and the inputs come back unchanged
    (poisson.cpp:154-157, 263-269).
    """
    outs = [np.array(a, dtype=np.float64, order="F", copy=True) for a in input_image]
    reps = [np.asfortranarray(np.asarray(a, dtype=np.float64)) for a in replacement_image]
    mask = np.asfortranarray(np.asarray(invalid_mask).astype(np.bool_, copy=False))
    if any(a.ndim != 2 for a in outs + reps) or mask.ndim != 2:
        raise TypeError("blend_images_poisson(): expected lists of 2-D arrays and a 2-D mask")
    if not outs:
        return outs
    if len(outs) != len(reps) or any(a.shape != outs[0].shape for a in outs + reps):
        _log.error("Input and replacement images must have the same dimensions")  # poisson.cpp:154-157
        return outs
    if mask.shape != outs[0].shape:
        _log.error("Invalid mask must match the image dimensions")  # poisson.cpp:158-160 logs; continuing would read
        return outs  # out of bounds in the reference (App. B4): return the inputs unchanged instead
    global _last_perf
    work = [a.copy(order="F") for a in outs]
    _last_perf = default_context().poisson_blend(
        work, reps, mask, tolerance=tolerance, max_iterations=max_iterations, precond=_defaults["precond"],
        check_every=_defaults["check_every"],
    )  # fmt: skip
    if any(s["status"] == SA_NOT_CONVERGED for s in _last_perf):
        _log.error("Failed to solve the linear system (no convergence)")  # poisson.cpp:263-269
        return outs
    return work


def valid_pixel_mask(replacement_image: Sequence[np.ndarray]) -> np.ndarray:
    """MultiChannelImage::valid_pixel over a whole image (approx/utils.h:101-105): a pixel whose first three channels
    all truncate to 1 is the white key (invalid); everything else is part of the pasted region."""
    a = [np.asarray(c) for c in replacement_image[:3]]
    key = (a[0].astype(np.int64) == 1) & (a[1].astype(np.int64) == 1) & (a[2].astype(np.int64) == 1)
    return ~key


def blend_images_poisson_offset(input_image: Sequence[np.ndarray], replacement_image: Sequence[np.ndarray], start_row: int,
                                start_column: int, tolerance: float = 1e-12) -> None:  # fmt: skip
    """Offset / white-key overload of approx::blend_images_poisson (poisson.h:30-33, poisson.cpp:21-143; the README's
    beach / chair demo): `replacement_image` is pasted into `input_image` at (start_row, start_column); its unknowns are
    the pixels that are not the white key; the arrays of `input_image` are modified IN PLACE like the reference's
    `MultiChannelImage&`.  The three bounds checks log and return (poisson.cpp:25-39).

    The system lives in the replacement's own rectangle (neighbours outside it are dropped, poisson.cpp:76,108) and its
    boundary values are the input pixels at the offset, so it is the mask overload on the crop.  The reference factorises
    the matrix; here it goes through the same CG to `tolerance`."""
    ins = list(input_image)
    reps = [np.asarray(a, dtype=np.float64) for a in replacement_image]
    if not ins or len(reps) < 3 or len(reps) < len(ins):
        raise TypeError("blend_images_poisson_offset(): the replacement needs >= 3 channels and one per input channel")
    rows, cols = ins[0].shape
    R, C_ = reps[0].shape
    if R * C_ > rows * cols:
        _log.error("Cannot solve problem: replacement image is larger than the input image")
        return
    if start_row < 0 or start_column < 0 or start_row >= rows or start_column >= cols:
        _log.error("Cannot solve problem: row/column is out of bounds")
        return
    if start_row + R > rows or start_column + C_ > cols:
        _log.error("Cannot solve problem: replacement image goes beyond the bounds of the input image")
        return
    unknown = np.asfortranarray(valid_pixel_mask(reps))
    sl = (slice(start_row, start_row + R), slice(start_column, start_column + C_))
    work = [np.array(a[sl], dtype=np.float64, order="F", copy=True) for a in ins]
    g = [np.asfortranarray(reps[b]) for b in range(len(ins))]
    global _last_perf
    _last_perf = default_context().poisson_blend(work, g, unknown, tolerance=tolerance, max_iterations=2**31 - 2,
                                                 precond=_defaults["precond"])  # fmt: skip
    if any(s["status"] == SA_NOT_CONVERGED for s in _last_perf):
        _log.error("Failed to solve the linear system (no convergence)")
        return
    for a, w in zip(ins, work):  # poisson.cpp:126-139: only the unknowns are written
        a[sl][unknown] = w[unknown]


GAMMA = 2.2  # approx/source/utils.cpp:8


def image_to_channels(image_bgr_u8: np.ndarray) -> list[np.ndarray]:
    """The arithmetic of approx::read_image (approx/source/utils.cpp:16-34) on an image that is already in memory
    (uint8 H x W x 3 in cv::imread's B, G, R order): three float64 channels R, G, B, gamma-decoded pow(v / 255, 1 / 2.2).
    File I/O itself (cv::imread) is outside the path."""
    a = np.asarray(image_bgr_u8)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise TypeError("image_to_channels(): uint8 H x W x 3")
    return [np.power(a[..., k] / 255.0, 1.0 / GAMMA) for k in (2, 1, 0)]


def channels_to_image(channels: Sequence[np.ndarray]) -> Optional[np.ndarray]:
    """approx::image_list_to_cv (approx/source/utils.cpp:36-60): float64 channels R, G, B -> uint8 H x W x 3 in B, G, R
    order, static_cast<uchar>(pow(v, 2.2) * 255) (truncation towards zero, wrap-around of out-of-range values like the
    C++ cast).  Anything but three channels is logged and refused."""
    if len(channels) != 3:
        _log.warning("Image with less than 3 channels is not supported. (%d channels provided)", len(channels))
        return None
    out = np.empty(channels[0].shape + (3,), np.uint8)
    for k, c in zip((2, 1, 0), channels):
        out[..., k] = (np.power(np.asarray(c, np.float64), GAMMA) * 255.0).astype(np.int64).astype(np.uint8)
    return out


def highlight_area_replaced(input_image: Sequence[np.ndarray], replacement_image: Sequence[np.ndarray], start_row: int,
                            start_column: int, color: Sequence[float]) -> None:  # fmt: skip
    """approx::highlight_area_replaced (poisson.cpp:305-321): paints the pasted (non white-key) pixels of the replacement
    at the offset with `color` in the first three channels of `input_image`, in place."""
    m = valid_pixel_mask(replacement_image)
    R, C_ = m.shape
    for k in range(3):
        input_image[k][start_row : start_row + R, start_column : start_column + C_][m] = color[k]


class ConnectedComponents:
    """approx::ConnectedComponents (laplace.h:11-14): `matrix` (labels, 0 = valid) and `region_map` label -> pixels."""

    def __init__(self, matrix: np.ndarray, num_labels: int):
        self.matrix = matrix
        self.num_labels = num_labels
        self._region_map = None

    @property
    def region_map(self) -> dict[int, np.ndarray]:
        """label -> int64[n, 2] (row, col) in row-major raster order.  Built lazily from `matrix` (host bookkeeping)."""
        if self._region_map is None:
            flat = self.matrix.ravel(order="C")
            idx = np.flatnonzero(flat)
            order = np.argsort(flat[idx], kind="stable")
            idx = idx[order]
            labels = flat[idx]
            cols = self.matrix.shape[1]
            rc = np.stack([idx // cols, idx % cols], axis=1).astype(np.int64)
            cuts = np.flatnonzero(np.diff(labels)) + 1
            parts = np.split(rc, cuts)
            self._region_map = {int(labels[s]): p for s, p in zip(np.concatenate([[0], cuts]), parts)} if len(idx) else {}
        return self._region_map


def find_connected_components(invalid: np.ndarray) -> ConnectedComponents:
    """approx::find_connected_components (laplace.h:20; contract tests/approximation.h:55-75)."""
    if not (isinstance(invalid, np.ndarray) and invalid.ndim == 2 and invalid.dtype in (np.bool_, np.uint8)):
        raise TypeError("find_connected_components(): expected a 2-D bool array")
    lab, k = default_context().label_components(invalid)
    return ConnectedComponents(lab, k)


def mask_scan(mask: np.ndarray):
    return default_context().mask_scan(mask)


def unknown_numbering(mask: np.ndarray):
    return default_context().unknown_numbering(mask)


def valid_neighbours(rows: int, cols: int, row: int, col: int) -> list[tuple[int, int]]:
    """approx::valid_neighbours (utils.h:35-50): in-image 4-neighbours in the order (-1,0) (+1,0) (0,-1) (0,+1).

    Pure index arithmetic on two integers (API value type); tests/approximation.h:9-33 pins the counts.
    """
    cand = [(row - 1, col), (row + 1, col), (row, col - 1), (row, col + 1)]
    return [(r, c) for r, c in cand if 0 <= r < rows and 0 <= c < cols]
