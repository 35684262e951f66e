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


def test_bench_scene_is_window_reproducible():
    """bench.py's scene (Gaussian-filtered sigma = 40 px white noise, analytic threshold; synth.cloud_mask /
    scene_band): any window can be built on its own and equals the same pixels of a larger window, and the numpy and torch
    generators agree -- that is what lets the CPU reference arm solve a crop of the VERY SAME scene the B200 arm fills."""
    import torch

    from satellite_approximation_b200 import synth

    big = synth.cloud_mask(500, 620, seed=2, row0=4000, col0=3900, clear_border=False)
    sub = synth.cloud_mask(300, 256, seed=2, row0=4096, col0=4100, clear_border=False)
    assert np.array_equal(big[96:396, 200:456], sub)
    assert 0.05 < big.mean() < 0.65  # 30 % cover over the tile; a window a few blobs wide fluctuates widely
    t = synth.torch_cloud_mask(300, 256, seed=2, device="cpu", row0=4096, col0=4100, clear_border=False).numpy().astype(bool)
    assert (t != sub).mean() < 1e-4  # float rounding of the two filters: at most a handful of pixels at the threshold
    cleared = synth.cloud_mask(300, 256, seed=2, row0=4096, col0=4100)
    assert not cleared[0].any() and not cleared[-1].any() and not cleared[:, 0].any() and not cleared[:, -1].any()
    assert not np.array_equal(synth.cloud_mask(64, 64, seed=2, row0=4096, col0=4100), synth.cloud_mask(64, 64, seed=19, row0=4096, col0=4100))
    b = synth.scene_band(120, 90, seed=5, row0=100, col0=50, total_rows=10980, total_cols=10980)
    whole = synth.scene_band(400, 300, seed=5, row0=0, col0=0, total_rows=10980, total_cols=10980)
    assert np.array_equal(whole[100:220, 50:140], b)
    bt = synth.torch_scene_band(120, 90, seed=5, device="cpu", row0=100, col0=50, total_rows=10980, total_cols=10980).numpy()
    assert np.max(np.abs(b - bt)) < 1e-9 and b.min() >= 0.0 and b.max() <= 10000.0
    assert abs(synth.cloud_threshold(40.0, 0.5)) < 1e-12 and synth.cloud_threshold(40.0, 0.3) > 0


def test_every_runtime_switch_of_the_library_is_documented():
    """INTEGRATION.md section 6 lists the environment variables libsatfill reads: no switch without a line there."""
    import glob
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    for path in glob.glob(os.path.join(root, "satellite_approximation_b200", "csrc", "*.cu*")):
        read.update(re.findall(r'(?:getenv|env_int)\("(SATFILL_[A-Z0-9_]+)"', open(path).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert len(read) >= 10
    assert not [v for v in sorted(read) if v not in doc]


def test_bench_kernel_classes_follow_the_update_mode(monkeypatch):
    """bench.py accounts the CG update per class: with the deferred x update (default) the passes that leave x alone are
    class 3 at 25 B per unknown and the two-step passes class 1 at 45 B -- 35 on average against the plain update's 41, which
    SATFILL_DEFER_X=0 selects in the library and in the table alike."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.delenv("SATFILL_DEFER_X", raising=False)
    names, bpu = bench.rb_tables(True)
    assert (bpu[1], bpu[3]) == (45.0, 25.0) and "XM=2" in names[1] and "XM=1" in names[3]
    assert (bpu[1] + bpu[3]) / 2 == 35.0 and len(names) == len(bpu) == bench.NK
    monkeypatch.setenv("SATFILL_DEFER_X", "0")
    names, bpu = bench.rb_tables(True)
    assert bpu[1] == 41.0 and names[1] == bench.KERNEL_NAMES_RB[1] and bpu == bench.BYTES_RB
    assert bench.rb_tables("cta")[1] == bench.BYTES_RB_CTA
