"""Seeded synthetic Sentinel-2-shaped inputs for tests and bench.py.  Data generation only: nothing
here is on the product path.  numpy generators for test-sized inputs, torch generators for the full-size benchmark
scenes (built directly in HBM so that a 12.5 GB scene does not have to be synthesised on the host)."""
from __future__ import annotations

import numpy as np


def blob_mask(rows: int, cols: int, cover: float = 0.3, sigma: float = 8.0, seed: int = 2, clear_border: bool = True):
    """Cloud-like blobs: threshold of Gaussian-filtered white noise at the (1 - cover) quantile."""
    from scipy.ndimage import gaussian_filter

    rng = np.random.default_rng(seed)
    f = gaussian_filter(rng.standard_normal((rows, cols)), sigma, mode="wrap")
    m = f > np.quantile(f, 1.0 - cover)
    if clear_border:
        m[0, :] = m[-1, :] = False
        m[:, 0] = m[:, -1] = False
    return m


def bernoulli_mask(rows: int, cols: int, cover: float = 0.3, seed: int = 2, clear_border: bool = True):
    rng = np.random.default_rng(seed)
    m = rng.random((rows, cols)) < cover
    if clear_border:
        m[0, :] = m[-1, :] = False
        m[:, 0] = m[:, -1] = False
    return m


def smooth_band(rows: int, cols: int, seed: int = 1, lo: float = 0.0, hi: float = 10000.0):
    """Smooth field + noise in [lo, hi] (a stand-in for a reflectance band)."""
    rng = np.random.default_rng(seed)
    r = np.linspace(0, 1, rows)[:, None]
    c = np.linspace(0, 1, cols)[None, :]
    k = rng.uniform(1.0, 6.0, size=4)
    ph = rng.uniform(0, 2 * np.pi, size=4)
    f = np.sin(k[0] * 2 * np.pi * r + ph[0]) * np.cos(k[1] * 2 * np.pi * c + ph[1]) + 0.5 * np.sin(
        k[2] * 2 * np.pi * (r + c) + ph[2]
    )
    f = f + 0.1 * rng.standard_normal((rows, cols))
    f = (f - f.min()) / (f.max() - f.min())
    return lo + (hi - lo) * f


def second_date(f: np.ndarray, seed: int = 0):
    """Synthetic guidance image for the Poisson blend: a radiometrically shifted, slightly noisy copy."""
    rng = np.random.default_rng(seed)
    return 0.9 * f + 37.0 + 5.0 * rng.standard_normal(f.shape)


def harmonic_field(rows: int, cols: int, kind: int = 0):
    """Discrete-harmonic fields: fixed points of the Laplace fill."""
    r = np.arange(rows, dtype=np.float64)[:, None]
    c = np.arange(cols, dtype=np.float64)[None, :]
    if kind == 0:
        return 3.0 * r - 2.0 * c + 7.0
    if kind == 1:
        return (r * r - c * c) * 1e-2 + 100.0
    return r * c * 1e-2 - 5.0


def region_mask(rows: int, cols: int, n_regions: int, area_lo: float = 1e2, area_hi: float = 5e4, seed: int = 3):
    """Non-touching random ellipses with log-uniform areas (config 4: many small holes)."""
    rng = np.random.default_rng(seed)
    m = np.zeros((rows, cols), bool)
    occupied = np.zeros((rows, cols), bool)
    rr, cc = np.mgrid[0:rows, 0:cols]
    placed = 0
    tries = 0
    while placed < n_regions and tries < 50 * n_regions:
        tries += 1
        area = np.exp(rng.uniform(np.log(area_lo), np.log(area_hi)))
        ratio = rng.uniform(0.5, 2.0)
        a = np.sqrt(area / np.pi * ratio)
        b = area / (np.pi * a)
        if 2 * a + 6 >= rows or 2 * b + 6 >= cols:
            continue
        cy = rng.uniform(a + 3, rows - a - 3)
        cx = rng.uniform(b + 3, cols - b - 3)
        y0, y1 = int(max(cy - a - 3, 0)), int(min(cy + a + 4, rows))
        x0, x1 = int(max(cx - b - 3, 0)), int(min(cx + b + 4, cols))
        if y1 - y0 < 3 or x1 - x0 < 3:
            continue
        sub_r, sub_c = rr[y0:y1, x0:x1], cc[y0:y1, x0:x1]
        e = ((sub_r - cy) / a) ** 2 + ((sub_c - cx) / b) ** 2 <= 1.0
        grown = ((sub_r - cy) / (a + 2)) ** 2 + ((sub_c - cx) / (b + 2)) ** 2 <= 1.0
        if (occupied[y0:y1, x0:x1] & grown).any() or not e.any():
            continue
        m[y0:y1, x0:x1] |= e
        occupied[y0:y1, x0:x1] |= grown
        placed += 1
    return m


def scene_mosaic_mask(scene: int = 2048, grid: int = 8, n_regions: int = 10000, seed: int = 3, area_lo: float = 1e2,
                      area_hi: float = 5e4):
    """BASELINE.json configs[3]: `n_regions` independent cloud regions across grid x grid scenes of scene x scene pixels,
    laid out as one mosaic.  Regions keep >= 3 pixels from their scene's border, so the scenes stay independent linear
    systems (4-connectivity) and the mosaic is solved exactly as the scenes would be one by one."""
    per = [n_regions // (grid * grid) + (1 if i < n_regions % (grid * grid) else 0) for i in range(grid * grid)]
    m = np.zeros((scene * grid, scene * grid), bool)
    for i, n in enumerate(per):
        y, x = divmod(i, grid)
        m[y * scene : (y + 1) * scene, x * scene : (x + 1) * scene] = region_mask(scene, scene, n, area_lo, area_hi, seed=seed + 101 * i)
    return m


# ---- torch generators (device) --------------------------------------------------------------------------------------


def torch_blob_mask(rows: int, cols: int, cover: float = 0.3, cell: int = 48, seed: int = 2, device="cuda"):
    """Cloud-like blobs at full tile size: bicubically upsampled coarse noise (correlation length ~ `cell` pixels)
    thresholded at the (1 - cover) quantile; one-pixel border ring cleared.  Returns a uint8 (0/1) tensor."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    cr, cc = rows // cell + 3, cols // cell + 3
    coarse = torch.randn((1, 1, cr, cc), generator=g, device=device, dtype=torch.float32)
    f = torch.nn.functional.interpolate(coarse, size=(cr * cell, cc * cell), mode="bicubic", align_corners=False)
    f = f[0, 0, cell : cell + rows, cell : cell + cols]
    sample = f[:: max(rows // 1024, 1), :: max(cols // 1024, 1)].flatten()
    thr = torch.quantile(sample, 1.0 - cover)
    m = (f > thr).to(torch.uint8)
    m[0, :] = 0
    m[-1, :] = 0
    m[:, 0] = 0
    m[:, -1] = 0
    return m.contiguous()


def torch_band(rows: int, cols: int, seed: int = 1, device="cuda"):
    """Smooth field + noise in [0, 10000], float64, built on the device."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    k = torch.rand(4, generator=g, device=device, dtype=torch.float64) * 5.0 + 1.0
    ph = torch.rand(4, generator=g, device=device, dtype=torch.float64) * 6.283185307179586
    r = torch.linspace(0, 1, rows, device=device, dtype=torch.float64)[:, None]
    c = torch.linspace(0, 1, cols, device=device, dtype=torch.float64)[None, :]
    f = torch.sin(k[0] * 6.283185307179586 * r + ph[0]) * torch.cos(k[1] * 6.283185307179586 * c + ph[1])
    f = f + 0.5 * torch.sin(k[2] * 6.283185307179586 * (r + c) + ph[2])
    f = f + 0.1 * torch.randn((rows, cols), generator=g, device=device, dtype=torch.float64)
    f = (f + 1.9) * (10000.0 / 3.8)
    return f.clamp_(0.0, 10000.0).contiguous()


# ---- the benchmark scene of SURVEY.md 8d (configs[2]), window-reproducible ------------------------------------------------
# "mask = threshold at the 70th percentile of Gaussian-filtered (sigma ~ 40 px) white noise, border ring cleared".  Both
# bench arms must see the SAME scene: the B200 arm builds the whole 10980^2 tile in HBM, the CPU reference arm solves a
# crop of it.  So the white noise is a counter-based hash of the GLOBAL pixel coordinates (any window can be generated on
# its own, identically in numpy and torch), the Gaussian is truncated at 4 sigma (a window needs a halo of that many
# noise pixels, nothing else), and the threshold is the analytic (1 - cover) quantile of the filtered field (a weighted sum
# of ~4 pi sigma^2 = 20000 uniform variates is Gaussian to every digit that matters), not a quantile of the realisation.


def _hash_uniform(r, c, seed: int):
    """Uniform noise in [-0.5, 0.5) at global coordinates (r, c): int64 tensors / arrays (numpy or torch), broadcastable.
    Plain 64-bit integer arithmetic that both libraries wrap identically; every intermediate is masked to 32 bits."""
    m = 0xFFFFFFFF
    h = (r * 0x9E3779B1 + c * 0x85EBCA77 + (int(seed) & m) * 0xC2B2AE3D + 0x27D4EB2F) & m
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & m
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & m
    h = h ^ (h >> 15)
    return h, 1.0 / 4294967296.0


def _gauss_kernel(sigma: float):
    radius = int(np.ceil(4.0 * sigma))
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum(), radius


def cloud_threshold(sigma: float, cover: float) -> float:
    """Analytic (1 - cover) quantile of the filtered field: zero mean, std = sqrt(1 / 12) * sum(k^2) for the separable
    kernel k (x) k applied to uniform noise of variance 1 / 12."""
    from statistics import NormalDist

    k, _ = _gauss_kernel(sigma)
    return NormalDist().inv_cdf(1.0 - cover) * float(np.sqrt(1.0 / 12.0) * np.sum(k * k))


def cloud_mask(rows: int, cols: int, cover: float = 0.3, sigma: float = 40.0, seed: int = 2, row0: int = 0, col0: int = 0,
               clear_border: bool = True):
    """numpy: the window [row0, row0 + rows) x [col0, col0 + cols) of the SURVEY 8d cloud mask with seed `seed` (bool).
    clear_border clears the one-pixel ring of THIS window (a crop is its own image: Laplace unknowns never sit on the
    image border, laplace.cpp:98-100)."""
    from scipy.ndimage import correlate1d

    k, R = _gauss_kernel(sigma)
    r = np.arange(row0 - R, row0 + rows + R, dtype=np.int64)[:, None]
    c = np.arange(col0 - R, col0 + cols + R, dtype=np.int64)[None, :]
    h, scale = _hash_uniform(r, c, seed)
    f = h.astype(np.float32) * np.float32(scale) - np.float32(0.5)
    k32 = k.astype(np.float32)
    f = correlate1d(f, k32, axis=0, mode="constant")[R:-R]
    f = correlate1d(f, k32, axis=1, mode="constant")[:, R:-R]
    m = f > np.float32(cloud_threshold(sigma, cover))
    if clear_border:
        m[0, :] = m[-1, :] = False
        m[:, 0] = m[:, -1] = False
    return m


def torch_cloud_mask(rows: int, cols: int, cover: float = 0.3, sigma: float = 40.0, seed: int = 2, device="cuda",
                     row0: int = 0, col0: int = 0, clear_border: bool = True):
    """torch: the same window of the same mask as `cloud_mask`, built on `device` (uint8 0/1).  Float rounding of the
    filter differs between the two, so a handful of pixels within 1e-6 of the threshold may differ; tests bound it."""
    import torch

    k, R = _gauss_kernel(sigma)
    r = torch.arange(row0 - R, row0 + rows + R, device=device, dtype=torch.int64)[:, None]
    c = torch.arange(col0 - R, col0 + cols + R, device=device, dtype=torch.int64)[None, :]
    h, scale = _hash_uniform(r, c, seed)
    f = h.to(torch.float32).mul_(scale).sub_(0.5)
    del h
    kt = torch.tensor(k, device=device, dtype=torch.float32)
    f = torch.nn.functional.conv2d(f[None, None], kt.view(1, 1, -1, 1))
    f = torch.nn.functional.conv2d(f, kt.view(1, 1, 1, -1))[0, 0]
    m = (f > float(np.float32(cloud_threshold(sigma, cover)))).to(torch.uint8)
    if clear_border:
        m[0, :] = 0
        m[-1, :] = 0
        m[:, 0] = 0
        m[:, -1] = 0
    return m.contiguous()


def _band_params(seed: int):
    rng = np.random.default_rng(1000003 + int(seed))
    return rng.uniform(1.0, 6.0, size=4), rng.uniform(0, 2 * np.pi, size=4)


def scene_band(rows: int, cols: int, seed: int = 1, row0: int = 0, col0: int = 0, total_rows: int | None = None,
               total_cols: int | None = None):
    """numpy: window of the benchmark band `seed` (smooth field + hash noise in [0, 10000], float64) of a
    total_rows x total_cols tile; any window of it comes out identically, and `torch_scene_band` builds the same values."""
    tr, tc = total_rows or rows, total_cols or cols
    k, ph = _band_params(seed)
    ri = np.arange(row0, row0 + rows, dtype=np.int64)[:, None]
    ci = np.arange(col0, col0 + cols, dtype=np.int64)[None, :]
    r = ri.astype(np.float64) / max(tr - 1, 1)
    c = ci.astype(np.float64) / max(tc - 1, 1)
    tau = 2 * np.pi
    f = np.sin(k[0] * tau * r + ph[0]) * np.cos(k[1] * tau * c + ph[1]) + 0.5 * np.sin(k[2] * tau * (r + c) + ph[2])
    h, scale = _hash_uniform(ri, ci, 7919 + seed)
    f = f + 0.35 * (h.astype(np.float64) * scale - 0.5)
    return np.clip((f + 1.9) * (10000.0 / 3.8), 0.0, 10000.0)


def torch_scene_band(rows: int, cols: int, seed: int = 1, device="cuda", row0: int = 0, col0: int = 0,
                     total_rows: int | None = None, total_cols: int | None = None):
    import torch

    tr, tc = total_rows or rows, total_cols or cols
    k, ph = _band_params(seed)
    ri = torch.arange(row0, row0 + rows, device=device, dtype=torch.int64)[:, None]
    ci = torch.arange(col0, col0 + cols, device=device, dtype=torch.int64)[None, :]
    r = ri.to(torch.float64) / max(tr - 1, 1)
    c = ci.to(torch.float64) / max(tc - 1, 1)
    tau = 2 * np.pi
    f = torch.sin(k[0] * tau * r + ph[0]) * torch.cos(k[1] * tau * c + ph[1]) + 0.5 * torch.sin(k[2] * tau * (r + c) + ph[2])
    h, scale = _hash_uniform(ri, ci, 7919 + seed)
    f = f + 0.35 * (h.to(torch.float64) * scale - 0.5)
    return ((f + 1.9) * (10000.0 / 3.8)).clamp_(0.0, 10000.0).contiguous()
