"""CPU-side statistics of a workload mask under candidate tilings (no GPU): how much of what the tile kernels touch is
algorithmic work.  For the default bench mask (synth.cloud_mask: SURVEY 8d, sigma = 40 px, 10980^2, 30 % cover; --bicubic: round 1's mask) and a tile shape it prints

  * the share of tiles that hold at least one unknown (the tile lists the kernels walk),
  * the fill of those tiles (unknowns / cells): the padding the dense-in-tile kernels pay,
  * the halo factor (cells of tile + one-cell frame) / (cells of tile): what a frame-staging kernel fetches per cell,
  * frame reads per unknown: unknowns inside the H-wide frame windows of the active tiles / unknowns -- the read
    amplification of a frame-staging kernel whose loads are predicated on the unknown bits (H = 1: CG halo; H = 4: the
    level-0 red-black cycle kernels),
  * unknown runs per active tile row and their mean length, and the share of 32-byte DRAM sectors of the active tiles'
    rows that hold at least one unknown double: what a predicated, run-following access pattern actually moves.

    python tools/tile_stats.py [rows cols]          (defaults to the bench's 10980 x 10980)
Feeds DESIGN.md section 10 (items 1 and 2)."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stats(mask: np.ndarray, th: int, tw: int) -> dict:
    rows, cols = mask.shape
    R, C = -(-rows // th) * th, -(-cols // tw) * tw
    m = np.zeros((R, C), bool)
    m[:rows, :cols] = mask
    t = m.reshape(R // th, th, C // tw, tw).transpose(0, 2, 1, 3)  # tile-major view
    per_tile = t.reshape(t.shape[0], t.shape[1], -1).sum(-1)
    active = per_tile > 0
    n_active = int(active.sum())
    unknowns = int(mask.sum())
    at = t[active]  # (n_active, th, tw)
    # runs of unknowns per tile row
    starts = at & ~np.concatenate([np.zeros_like(at[..., :1]), at[..., :-1]], axis=-1)
    runs = int(starts.sum())
    rows_with = int(at.any(-1).sum())
    # 32-byte sectors (4 doubles) of the active tiles' rows that contain an unknown
    sec = at.reshape(n_active, th, tw // 4, 4).any(-1)
    # unknowns inside the H-wide frame windows of the active tiles / unknowns: what a frame-staging kernel with loads
    # predicated on the unknown bits reads per unknown (H = 1: the CG stencil's halo, H = 4: the level-0 cycle kernels)
    sat = np.zeros((R + 1, C + 1), np.int64)
    np.cumsum(m, axis=0, out=sat[1:, 1:])
    np.cumsum(sat[1:, 1:], axis=1, out=sat[1:, 1:])
    ty, tx = np.nonzero(active)
    amp = {}
    for H in (1, 4):
        r0, r1 = np.clip(ty * th - H, 0, R), np.clip(ty * th + th + H, 0, R)
        c0, c1 = np.clip(tx * tw - H, 0, C), np.clip(tx * tw + tw + H, 0, C)
        inside = sat[r1, c1] - sat[r0, c1] - sat[r1, c0] + sat[r0, c0]
        amp[H] = float(inside.sum()) / unknowns
    return {
        "amp1": amp[1], "amp4": amp[4],
        "tile": f"{th}x{tw}", "tiles": int(active.size), "active_share": n_active / active.size,
        "fill_of_active": unknowns / (n_active * th * tw), "halo_factor": (th + 2) * (tw + 2) / (th * tw),
        "runs_per_active_row": runs / max(rows_with, 1), "mean_run": unknowns / max(runs, 1),
        "sector_share": float(sec.sum()) / sec.size, "sector_efficiency": unknowns / (float(sec.sum()) * 4),
        "full_tiles_share": float((per_tile == th * tw).sum()) / max(n_active, 1),
    }  # fmt: skip


def granule_amplification(mask: np.ndarray, elem: int, granule: int) -> float:
    """Bytes of the `granule`-byte DRAM units that hold at least one unknown / bytes of the unknowns, for a row-major plane
    of `elem`-byte cells (rows padded to a multiple of the granule): what a run-following access pattern moves when DRAM
    is filled `granule` bytes at a time."""
    per = granule // elem
    rows, cols = mask.shape
    C = -(-cols // per) * per
    m = np.zeros((rows, C), bool)
    m[:, :cols] = mask
    touched = m.reshape(rows, C // per, per).any(-1).sum()
    return float(touched) * granule / (float(mask.sum()) * elem)


def main() -> None:
    import torch

    from satellite_approximation_b200 import synth

    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rows = int(args[0]) if len(args) > 0 else 10980
    cols = int(args[1]) if len(args) > 1 else rows
    if "--bicubic" in sys.argv:  # round 1's bench mask: bicubically upsampled 48-pixel noise (blobs of about one tile)
        mask = synth.torch_blob_mask(rows, cols, device="cpu").numpy().astype(bool)
        print(f"# synth.torch_blob_mask({rows}, {cols}): {mask.mean() * 100:.2f} % unknown ({int(mask.sum())} pixels)")
    else:  # the bench mask: SURVEY.md 8d, Gaussian-filtered (sigma = 40 px) white noise
        mask = synth.cloud_mask(rows, cols)
        print(f"# synth.cloud_mask({rows}, {cols}, sigma=40): {mask.mean() * 100:.2f} % unknown ({int(mask.sum())} pixels)")
    print("# DRAM bytes moved per byte of unknowns in a row-major plane, by fill granularity:")
    for elem, name in ((8, "double"), (4, "float")):
        print(f"#   {name:6s}  32 B sectors: {granule_amplification(mask, elem, 32):.3f}   64 B (sector pairs): "
              f"{granule_amplification(mask, elem, 64):.3f}   128 B lines: {granule_amplification(mask, elem, 128):.3f}")
    print("# tile   active  fill   full-tiles  halo   runs/row  mean-run  sectors-touched  unknowns/sector-cell  "
          "frame-reads/unknown H=1  H=4")
    for th, tw in ((32, 32), (32, 64), (64, 32), (64, 64), (16, 64), (16, 128), (8, 128)):
        s = stats(mask, th, tw)
        print(f"{s['tile']:>7}  {s['active_share']:.3f}   {s['fill_of_active']:.3f}  {s['full_tiles_share']:.3f}       "
              f"{s['halo_factor']:.3f}  {s['runs_per_active_row']:.2f}      {s['mean_run']:6.1f}    {s['sector_share']:.3f}"
              f"            {s['sector_efficiency']:.3f}                 {s['amp1']:.3f}                 {s['amp4']:.3f}")  # fmt: skip


if __name__ == "__main__":
    main()
