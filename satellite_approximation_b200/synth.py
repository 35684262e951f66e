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
