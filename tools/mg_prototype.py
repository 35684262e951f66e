"""Numerical prototype of the masked-grid multigrid preconditioner (numpy, CPU).  Design tool only -- not on the
product path and not the oracle.  Compares smoothers / precisions by PCG iteration count to 1e-6 on a cloud-like mask.

    python tools/mg_prototype.py [edge] [cell]
"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from satellite_approximation_b200 import synth  # noqa: E402


def pad(x):
    return np.pad(x, 1)


def nsum(x):
    p = pad(x)
    return p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]


def A(x, m, d=4.0):
    return np.where(m, d * x - nsum(x), 0).astype(x.dtype)


def shifted(a, di, dj, fill):
    """b[i, j] = a[i + di, j + dj], `fill` outside."""
    p = np.pad(a, 2, constant_values=fill)
    return p[2 + di : 2 + di + a.shape[0], 2 + dj : 2 + dj + a.shape[1]]


def coarse_diagonal(mf, fixed=True):
    """Diagonal of the boundary-corrected coarse operator (mg.cu, k_coarsen_mask): sum over the four arms of a coarse
    cell of 1 (the fine cell half way is an unknown), 2 (it is known: Dirichlet boundary at half the spacing) or 0
    (Poisson only: it, or the arm's end, lies outside the image)."""
    inside = np.ones(mf.shape, bool)
    d = np.zeros(mf.shape, np.float64)
    for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        mid_in, end_in = shifted(inside, di, dj, False), shifted(inside, 2 * di, 2 * dj, False)
        mid_unknown = shifted(mf, di, dj, False)
        arm = np.where(~mid_in, 2.0 if fixed else 0.0, np.where(~mid_unknown, 2.0, np.where(end_in | fixed, 1.0, 0.0)))
        d += arm
    return np.maximum(d, 1.0)[::2, ::2]


def restrict(t, mc):
    p = pad(t)
    w = (0.25 * (p[:-2, :-2] + p[:-2, 2:] + p[2:, :-2] + p[2:, 2:]) + 0.5 * (p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:])
         + p[1:-1, 1:-1])
    return np.where(mc, w[::2, ::2], 0).astype(t.dtype)


def prolong(e, mf):
    R, C = mf.shape
    ep = np.pad(e, ((0, 1), (0, 1)))
    out = np.zeros((2 * e.shape[0], 2 * e.shape[1]), e.dtype)
    out[0::2, 0::2] = ep[:-1, :-1]
    out[0::2, 1::2] = 0.5 * (ep[:-1, :-1] + ep[:-1, 1:])
    out[1::2, 0::2] = 0.5 * (ep[:-1, :-1] + ep[1:, :-1])
    out[1::2, 1::2] = 0.25 * (ep[:-1, :-1] + ep[:-1, 1:] + ep[1:, :-1] + ep[1:, 1:])
    return np.where(mf, out[:R, :C], 0).astype(e.dtype)


def colour(shape):
    i, j = np.indices(shape)
    return (i + j) % 2 == 0


class MG:
    def __init__(self, mask, smoother="jac2", dtype=np.float64, omega=0.8, levels=12, coarse_sweeps=32, post=None,
                 corrected=False):
        """corrected: boundary-corrected coarse diagonals (the red-black cycle of mg_rb.cu); else diagonal 4."""
        self.m = [mask]
        while len(self.m) < levels and min(self.m[-1].shape) >= 5:  # the next level has >= 3 rows and columns (mg.cu)
            self.m.append(self.m[-1][::2, ::2].copy())
        while len(self.m) > 1 and not self.m[-1].any():
            self.m.pop()
        self.red = [colour(m.shape) for m in self.m]
        self.d = [np.float64(4.0)] + [
            (coarse_diagonal(self.m[l]) if corrected else np.float64(4.0)).astype(dtype) for l in range(len(self.m) - 1)
        ]
        self.s, self.dt, self.om, self.cs = smoother, dtype, omega, coarse_sweeps

    def pre(self, l, b):
        m = self.m[l]
        d = self.d[l]
        if self.s.startswith("jac"):
            nu = int(self.s[3:])
            x = self.om * b / d
            for _ in range(nu - 1):
                x = x + self.om * (b - A(x, m, d)) / d
            return x.astype(self.dt)
        nu = int(self.s[2:])
        x = np.zeros_like(b)
        for _ in range(nu):
            for c in (self.red[l], ~self.red[l]):
                x = np.where(m & c, (b + nsum(x)) / d, x).astype(self.dt)
        return x

    def post(self, l, x, b):
        m, d = self.m[l], self.d[l]
        if self.s.startswith("jac"):
            for _ in range(int(self.s[3:])):
                x = (x + self.om * (b - A(x, m, d)) / d).astype(self.dt)
            return x
        for _ in range(int(self.s[2:])):
            for c in (~self.red[l], self.red[l]):
                x = np.where(m & c, (b + nsum(x)) / d, x).astype(self.dt)
        return x

    def cycle(self, l, b):
        m, d = self.m[l], self.d[l]
        if l == len(self.m) - 1:
            x = np.zeros_like(b)
            if self.s.startswith("jac"):
                x = self.om * b / d
                for _ in range(self.cs - 1):
                    x = x + self.om * (b - A(x, m, d)) / d
                return x.astype(self.dt)
            for _ in range(self.cs // 2):
                for c in (self.red[l], ~self.red[l]):
                    x = np.where(m & c, (b + nsum(x)) / d, x).astype(self.dt)
            for _ in range(self.cs // 2):
                for c in (~self.red[l], self.red[l]):
                    x = np.where(m & c, (b + nsum(x)) / d, x).astype(self.dt)
            return x
        x = self.pre(l, b)
        t = (b - A(x, m, d)).astype(self.dt)
        bc = restrict(t, self.m[l + 1])
        ec = self.cycle(l + 1, bc)
        x = (x + prolong(ec, m)).astype(self.dt)
        return self.post(l, x, b)

    def apply(self, r):
        return self.cycle(0, r.astype(self.dt)).astype(np.float64)


def pcg(mask, b, M, tol=1e-6, maxit=200):
    x = np.zeros_like(b)
    r = b.copy()
    bn = (b * b).sum()
    z = M(r)
    p = z.copy()
    rz = (r * z).sum()
    for k in range(1, maxit + 1):
        q = A(p, mask)
        a = rz / (p * q).sum()
        x += a * p
        r -= a * q
        if (r * r).sum() < tol * tol * bn:
            return x, k
        z = M(r)
        rz2 = (r * z).sum()
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    cell = float(sys.argv[2]) if len(sys.argv) > 2 else 48
    mask = synth.blob_mask(n, n, cover=0.3, sigma=cell / 3.0, seed=2)
    img = synth.smooth_band(n, n, seed=100)
    b = np.where(mask, nsum(np.where(mask, 0, img)), 0.0)
    print(f"{n}x{n}, {mask.sum()} unknowns")
    for name, kw in [
        ("jacobi(2,2) w=0.8 f64", dict(smoother="jac2")),
        ("jacobi(2,2) w=0.8 f32", dict(smoother="jac2", dtype=np.float32)),
        ("rb-gs(1,1) f64", dict(smoother="rb1")),
        ("rb-gs(1,1) f32", dict(smoother="rb1", dtype=np.float32)),
        ("rb-gs(1,1) f32 corrected", dict(smoother="rb1", dtype=np.float32, corrected=True)),
        ("rb-gs(2,2) f32", dict(smoother="rb2", dtype=np.float32)),
        ("jacobi(1,1) w=0.8 f64", dict(smoother="jac1")),
        ("jacobi(3,3) w=0.8 f64", dict(smoother="jac3")),
    ]:
        mg = MG(mask, **kw)
        t0 = time.time()
        for tol in (1e-6, 1e-10):
            x, k = pcg(mask, b, mg.apply, tol=tol)
            res = np.sqrt(((b - A(x, mask)) ** 2).sum() / (b * b).sum())
            print(f"  {name:26s} tol {tol:g}: {k:3d} iterations, true rel residual {res:.2e}  ({time.time() - t0:.1f}s)")
