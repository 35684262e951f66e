"""CPU: the analysis helpers under tools/ whose numbers DESIGN.md quotes (tools/tile_stats.py)."""
from __future__ import annotations

import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tile_stats_known_answers():
    ts = _load("tile_stats")
    m = np.zeros((128, 128), bool)
    m[32:64, 32:64] = True  # exactly one full 32 x 32 tile
    s = ts.stats(m, 32, 32)
    assert s["active_share"] == 1 / 16 and s["fill_of_active"] == 1.0 and s["full_tiles_share"] == 1.0
    assert s["runs_per_active_row"] == 1.0 and s["mean_run"] == 32.0 and s["sector_efficiency"] == 1.0
    assert s["amp1"] == 1.0 and s["amp4"] == 1.0  # the frame around an isolated tile holds no further unknowns
    assert abs(s["halo_factor"] - 34 * 34 / 1024) < 1e-12
    # a dense hole: every frame cell is an unknown, so the frame reads are the geometric halo factor (clipped at the border)
    d = np.ones((128, 128), bool)
    s = ts.stats(d, 32, 32)
    inner = (2 * 36 + 2 * 40) ** 2 / 128**2  # per axis: two border tiles see 36 cells, two inner tiles 40
    assert abs(s["amp4"] - inner) < 1e-12 and s["active_share"] == 1.0
    # fill granularity: a run of 5 doubles starting at element 3 touches 2 sectors of 4, 1 pair of 8
    r = np.zeros((1, 64), bool)
    r[0, 3:8] = True
    assert ts.granule_amplification(r, 8, 32) == 2 * 32 / 40 and ts.granule_amplification(r, 8, 64) == 64 / 40
    assert ts.granule_amplification(r, 4, 32) == 32 / 20 and ts.granule_amplification(r, 4, 128) == 128 / 20
