"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the CPU checkers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``satellite_approximation_b200``) never does; it fails loudly when its CUDA
library is missing instead of falling back to anything in here.

Two libraries:

* ``port``  -- ``oracle/_build/liboracle.so`` from ``satfill_oracle.c``: a plain-C restatement of
  ``lib/approx/source/laplace.cpp:31-120`` and ``lib/approx/source/poisson.cpp:145-290`` (function-level citations are
  in the C file).  Builds anywhere with gcc.
* ``ref``   -- ``oracle/_ref/libref_eigen.so`` from ``ref_eigen.cpp``: the same assembly executed by the reference's own
  vendored Eigen CG.  Buildable only where ``/root/reference`` exists (this container); the built ``.so`` travels to
  the GPU box.  ``ref()`` returns ``None`` when it is not there.

All image arguments are numpy arrays; any strides are accepted (the reference's ``MatX`` is column-major).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_eigen.so")
EIGEN_DIR = os.environ.get("EIGEN_DIR", "/root/reference/thirdparty/eigen-master")

OK, EMPTY, NOT_CONVERGED, BAD_ARG = 0, 1, 2, 3


def build(ref: bool = True) -> None:
    """Compile the checkers (called by ``__graft_entry__.build``)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    if ref and os.path.isdir(os.path.join(EIGEN_DIR, "Eigen")):
        subprocess.run(["make", "-C", _HERE, "ref", f"EIGEN_DIR={EIGEN_DIR}"], check=True, capture_output=True)


class _Stats(C.Structure):
    _fields_ = [
        ("unknowns", C.c_int64),
        ("system_size", C.c_int64),
        ("iterations", C.c_int64),
        ("error", C.c_double),
        ("assemble_s", C.c_double),
        ("solve_s", C.c_double),
    ]


@dataclass
class Stats:
    unknowns: int = 0
    system_size: int = 0
    iterations: int = 0
    error: float = 0.0
    assemble_s: float = 0.0
    solve_s: float = 0.0
    status: int = 0


def _strides_in_elements(a: np.ndarray) -> tuple[int, int]:
    assert a.ndim == 2
    assert a.strides[0] % a.itemsize == 0 and a.strides[1] % a.itemsize == 0
    return a.strides[0] // a.itemsize, a.strides[1] // a.itemsize


def _mask_u8(mask: np.ndarray) -> np.ndarray:
    assert mask.ndim == 2
    if mask.dtype == np.bool_:
        return mask.view(np.uint8)
    assert mask.dtype == np.uint8
    return mask


_p = C.c_void_p


class Port:
    """Plain-C restatement (kind = "port")."""

    def __init__(self) -> None:
        if not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(
            os.path.join(_HERE, "satfill_oracle.c")
        ):
            build(ref=False)
        self.lib = C.CDLL(_PORT_SO)
        L = self.lib
        L.so_valid_neighbours.restype = C.c_int
        L.so_valid_neighbours.argtypes = [C.c_int64] * 4 + [_p]
        L.so_mask_scan.restype = C.c_int64
        L.so_mask_scan.argtypes = [_p] + [C.c_int64] * 4 + [_p, _p]
        L.so_unknown_numbering.restype = C.c_int64
        L.so_unknown_numbering.argtypes = [_p] + [C.c_int64] * 4 + [_p]
        L.so_label_components.restype = C.c_int32
        L.so_label_components.argtypes = [_p] + [C.c_int64] * 4 + [_p]
        L.so_laplace_fill.restype = C.c_int
        L.so_laplace_fill.argtypes = [_p, _p] + [C.c_int64] * 4 + [C.c_int, C.c_double, C.c_int64, _p]
        L.so_poisson_blend.restype = C.c_int
        L.so_poisson_blend.argtypes = [_p, _p, C.c_int, _p] + [C.c_int64] * 4 + [C.c_double, C.c_int64, _p]
        L.so_relative_residual.restype = C.c_double
        L.so_relative_residual.argtypes = [_p, _p, _p] + [C.c_int64] * 4 + [C.c_int]

    # -- integer path ------------------------------------------------------------------------------------------
    def valid_neighbours(self, rows: int, cols: int, r: int, c: int) -> list[tuple[int, int]]:
        out = (C.c_int64 * 8)()
        n = self.lib.so_valid_neighbours(rows, cols, r, c, out)
        return [(out[2 * i], out[2 * i + 1]) for i in range(n)]

    def mask_scan(self, mask: np.ndarray):
        m = _mask_u8(mask)
        rs, cs = _strides_in_elements(m)
        bbox = np.zeros(4, np.int64)
        n = self.lib.so_mask_scan(m.ctypes.data, m.shape[0], m.shape[1], rs, cs, None, bbox.ctypes.data)
        px = np.zeros((n, 2), np.int64)
        self.lib.so_mask_scan(m.ctypes.data, m.shape[0], m.shape[1], rs, cs, px.ctypes.data, bbox.ctypes.data)
        return px, bbox

    def unknown_numbering(self, mask: np.ndarray):
        m = _mask_u8(mask)
        rs, cs = _strides_in_elements(m)
        num = np.empty(m.shape, np.int32)
        n = self.lib.so_unknown_numbering(m.ctypes.data, m.shape[0], m.shape[1], rs, cs, num.ctypes.data)
        return num, int(n)

    def label_components(self, mask: np.ndarray):
        m = _mask_u8(mask)
        rs, cs = _strides_in_elements(m)
        lab = np.empty(m.shape, np.int32)
        k = self.lib.so_label_components(m.ctypes.data, m.shape[0], m.shape[1], rs, cs, lab.ctypes.data)
        return lab, int(k)

    # -- float path --------------------------------------------------------------------------------------------
    def laplace_fill(self, img: np.ndarray, mask: np.ndarray, mode: int = 0, tol: float = 0.0, max_it: int = 0):
        """Returns (filled copy, Stats).  mode 0 = faithful bbox system, 1 = reduced SPD system."""
        out = np.array(img, dtype=np.float64, order="F", copy=True)
        m = _mask_u8(mask)
        assert m.shape == out.shape
        mm = np.asfortranarray(m)
        rs, cs = _strides_in_elements(out)
        st = _Stats()
        status = self.lib.so_laplace_fill(
            out.ctypes.data, mm.ctypes.data, out.shape[0], out.shape[1], rs, cs, mode, tol, max_it, C.byref(st)
        )
        return out, Stats(st.unknowns, st.system_size, st.iterations, st.error, st.assemble_s, st.solve_s, status)

    def poisson_blend(self, inputs, replacements, mask: np.ndarray, tol: float = 1e-6, max_it: int = -1):
        outs = [np.array(a, dtype=np.float64, order="F", copy=True) for a in inputs]
        reps = [np.asfortranarray(np.asarray(a, dtype=np.float64)) for a in replacements]
        mm = np.asfortranarray(_mask_u8(mask))
        nb = len(outs)
        rows, cols = mm.shape
        rs, cs = _strides_in_elements(mm)
        ins = (C.c_void_p * nb)(*[a.ctypes.data for a in outs])
        rps = (C.c_void_p * nb)(*[a.ctypes.data for a in reps])
        st = (_Stats * nb)()
        status = self.lib.so_poisson_blend(ins, rps, nb, mm.ctypes.data, rows, cols, rs, cs, tol, max_it, st)
        stats = [
            Stats(s.unknowns, s.system_size, s.iterations, s.error, s.assemble_s, s.solve_s, status) for s in st
        ]
        return outs, stats

    def relative_residual(self, u: np.ndarray, mask: np.ndarray, g: np.ndarray | None = None) -> float:
        """Reduced-system residual |b_U - A_UU x| / |b_U|; Laplace when g is None, Poisson otherwise."""
        uu = np.asfortranarray(np.asarray(u, dtype=np.float64))
        mm = np.asfortranarray(_mask_u8(mask))
        gg = None if g is None else np.asfortranarray(np.asarray(g, dtype=np.float64))
        rs, cs = _strides_in_elements(uu)
        return float(
            self.lib.so_relative_residual(
                uu.ctypes.data,
                None if gg is None else gg.ctypes.data,
                mm.ctypes.data,
                uu.shape[0],
                uu.shape[1],
                rs,
                cs,
                1 if g is None else 0,
            )
        )


class Ref:
    """The reference's arithmetic on the reference's own vendored Eigen (kind = "reference")."""

    def __init__(self) -> None:
        self.lib = C.CDLL(_REF_SO)
        L = self.lib
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_laplace_fill.restype = C.c_int
        L.ref_laplace_fill.argtypes = [_p, _p, C.c_int64, C.c_int64, C.c_double, C.c_int64, _p, _p, _p, _p, _p]
        L.ref_poisson_blend.restype = C.c_int
        L.ref_poisson_blend.argtypes = [_p, _p, C.c_int, _p, C.c_int64, C.c_int64, C.c_double, C.c_int64] + [_p] * 5
        self.set_threads(1)  # as shipped the approx library is single-threaded (lib/approx/CMakeLists.txt:10-14)

    def set_threads(self, n: int) -> None:
        self.lib.ref_set_threads(int(n))

    def laplace_fill(self, img: np.ndarray, mask: np.ndarray, tol: float = 0.0, max_it: int = 0):
        out = np.array(img, dtype=np.float64, order="F", copy=True)
        mm = np.asfortranarray(_mask_u8(mask))
        assert mm.shape == out.shape
        it, n = C.c_int64(0), C.c_int64(0)
        err, ta, ts = C.c_double(0), C.c_double(0), C.c_double(0)
        status = self.lib.ref_laplace_fill(
            out.ctypes.data, mm.ctypes.data, out.shape[0], out.shape[1], tol, max_it,
            C.byref(it), C.byref(err), C.byref(ta), C.byref(ts), C.byref(n),
        )  # fmt: skip
        return out, Stats(int(mm.sum()), n.value, it.value, err.value, ta.value, ts.value, status)

    def poisson_blend(self, inputs, replacements, mask: np.ndarray, tol: float = 1e-6, max_it: int = -1):
        outs = [np.array(a, dtype=np.float64, order="F", copy=True) for a in inputs]
        reps = [np.asfortranarray(np.asarray(a, dtype=np.float64)) for a in replacements]
        mm = np.asfortranarray(_mask_u8(mask))
        nb = len(outs)
        ins = (C.c_void_p * nb)(*[a.ctypes.data for a in outs])
        rps = (C.c_void_p * nb)(*[a.ctypes.data for a in reps])
        its = (C.c_int64 * nb)()
        errs = (C.c_double * nb)()
        secs = (C.c_double * nb)()
        setup, n = C.c_double(0), C.c_int64(0)
        status = self.lib.ref_poisson_blend(
            ins, rps, nb, mm.ctypes.data, mm.shape[0], mm.shape[1], tol, max_it, its, errs, secs,
            C.byref(setup), C.byref(n),
        )  # fmt: skip
        stats = [
            Stats(n.value, n.value, its[i], errs[i], setup.value if i == 0 else 0.0, secs[i], status)
            for i in range(nb)
        ]
        return outs, stats


_port: Port | None = None
_ref: Ref | None = None


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref() -> Ref | None:
    """None when oracle/_ref/libref_eigen.so has not been built (no /root/reference on this machine)."""
    global _ref
    if _ref is None and os.path.exists(_REF_SO):
        _ref = Ref()
    return _ref


# ---- the steps either side of the path ------------------------------------------------------------------
def rgb_invalid_mask(invalid_bgr: np.ndarray, red_threshold: float = 220.0) -> np.ndarray:
    """laplace.cpp:141-146: channels_cv[2] (red of a cv::imread BGR image) >= red_threshold and channels_cv[1]
    (green) <= 150, compared as doubles."""
    red = invalid_bgr[..., 2].astype(np.float64)
    green = invalid_bgr[..., 1].astype(np.float64)
    return (red >= red_threshold) & (green <= 150)


def apply_laplace(image_bgr: np.ndarray, invalid_bgr: np.ndarray, red_threshold: float = 220.0, engine=None, **kw):
    """Restatement of approx::apply_laplace (laplace.cpp:134-168): mask from the invalid image, every channel of the base
    image widened to double (cv2eigen) and filled on its own (fill_missing_portion_smooth_boundary), channels merged
    back (CV_64FC3).  `engine`: a Port (default, reduced system: border pixels are Dirichlet data) or a Ref."""
    mask = rgb_invalid_mask(invalid_bgr, red_threshold)
    eng = engine if engine is not None else port()
    out = np.empty(image_bgr.shape, np.float64)
    for k in range(image_bgr.shape[2]):
        ch = np.ascontiguousarray(image_bgr[..., k].astype(np.float64))
        if isinstance(eng, Port):
            filled, _ = eng.laplace_fill(ch, mask, mode=1, **kw)
        else:
            filled, _ = eng.laplace_fill(ch, mask, **kw)
        out[..., k] = filled
    return out, mask


def morph_close_mask(band: np.ndarray, radius: int = 5) -> np.ndarray:
    """Restatement of preprocess_cloud_band (executables/poisson-main.cpp:10-21): cv::morphologyEx(MORPH_CLOSE) with a
    (2 radius + 1)^2 MORPH_RECT element = dilation then erosion, both with OpenCV's default border for morphology
    (pixels outside the image never win the max / min: the window is clipped), then MatX<f64>::cast<bool>().
    Pinned against cv2 itself by tests/golden/prepost_cases.npz (oracle/make_golden.py)."""
    a = np.asarray(band, np.float64)
    rows, cols = a.shape

    def run(x, axis, fn, fill):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (radius, radius)
        p = np.pad(x, pad, constant_values=fill)
        out = None
        for k in range(2 * radius + 1):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(k, k + x.shape[axis])
            v = p[tuple(sl)]
            out = v.copy() if out is None else fn(out, v)
        return out

    d = run(run(a, 1, np.maximum, -np.inf), 0, np.maximum, -np.inf)
    e = run(run(d, 1, np.minimum, np.inf), 0, np.minimum, np.inf)
    return e != 0.0


def poisson_offset_dense(inputs, replacements, start_row: int, start_column: int):
    """Statement-for-statement restatement of the offset / white-key overload (poisson.cpp:21-143) for SMALL cases:
    raster-order numbering of the valid (non-key) replacement pixels (:54-64), A with the in-replacement neighbour count
    on the diagonal and -1 for valid neighbours (:70-91), b = sum of replacement gradients + input values at key
    neighbours (:104-122), a dense direct solve in place of the reference's sparse factorisation (:93-95,125), write-back
    of the unknowns only (:128-140).  Returns new input arrays."""
    ins = [np.array(a, np.float64, copy=True) for a in inputs]
    rep = [np.asarray(a, np.float64) for a in replacements]
    R, Cc = rep[0].shape

    def valid(r, c):  # MultiChannelImage::valid_pixel, approx/utils.h:101-105
        return not (int(rep[0][r, c]) == 1 and int(rep[1][r, c]) == 1 and int(rep[2][r, c]) == 1)

    def neighbours(r, c):  # valid_neighbours, approx/utils.h:35-50
        return [(r + dr, c + dc) for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)) if 0 <= r + dr < R and 0 <= c + dc < Cc]

    number = {}
    for r in range(R):
        for c in range(Cc):
            if valid(r, c):
                number[(r, c)] = len(number)
    n = len(number)
    if n == 0:
        return ins
    A = np.zeros((n, n))
    for (r, c), k in number.items():
        nb = neighbours(r, c)
        A[k, k] = float(len(nb))
        for q in nb:
            if q in number:
                A[k, number[q]] = -1.0
    for ch in range(len(ins)):
        b = np.zeros(n)
        for (r, c), k in number.items():
            for (nr, nc) in neighbours(r, c):
                b[k] += rep[ch][r, c] - rep[ch][nr, nc]
                if (nr, nc) not in number:
                    b[k] += ins[ch][nr + start_row, nc + start_column]
        x = np.linalg.solve(A, b)
        for (r, c), k in number.items():
            ins[ch][r + start_row, c + start_column] = x[k]
    return ins
