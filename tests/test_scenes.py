"""CPU: scene-level orchestration (satellite_approximation_b200/scenes.py) -- the SQLite bookkeeping
(lib/utils/source/db.cpp, lib/approx/source/db.cpp), find_good_close_image (lib/approx/source/poisson.cpp:323-349),
find_directory_contents (lib/utils/source/filesystem.cpp) and the folder driver the reference keeps commented out
(lib/approx/source/laplace.cpp:170-244).  The folder driver is exercised with the ORACLE as the pixel step (the product
default is the GPU; tests/test_zz_gpu_drivers.py runs that)."""
from __future__ import annotations

import datetime as dt
import math
import os
import sqlite3

import numpy as np
import pytest

from satellite_approximation_b200 import geotiff as gt
from satellite_approximation_b200 import scenes as sc
from satellite_approximation_b200 import synth


def test_simple_date_parsing_and_date_type():
    assert sc.parse_simple_date("2019-05-22") == dt.date(2019, 5, 22)
    assert sc.parse_simple_date("2002-1-25") == dt.date(2002, 1, 25)
    assert sc.parse_simple_date("2002-Jan-25") == dt.date(2002, 1, 25)
    assert sc.parse_simple_date("2002/February/3") == dt.date(2002, 2, 3)
    for bad in ("2002-13-01", "2002-02-30", "20020101", "2002-01", "a-b-c", ""):
        with pytest.raises(ValueError):
            sc.parse_simple_date(bad)
    d = sc.Date.parse("2019-5-2")
    assert str(d) == "2019-05-02" and d.sql() == (2019, 5, 2) and d == sc.Date(2019, 5, 2)
    assert sc.Date(2019, 5, 2) < sc.Date(2019, 5, 10) < sc.Date(2020, 1, 1)
    assert len({sc.Date(2019, 5, 2), sc.Date.parse("2019-05-02")}) == 1


def test_day_info_distance():
    i = sc.DayInfo(dt.date(2019, 5, 10), 0.25)
    assert i.distance(dt.date(2019, 5, 22), 0.5) == 0.5 * 12 + 0.5 * 0.25  # db.cpp:12-16
    assert i.distance(dt.date(2019, 5, 1), 1.0) == 9.0 and i.distance(dt.date(2019, 5, 1), 0.0) == 0.25


def _fill_dates(db, rows):
    for date, inv in rows:
        db.write_detection_result(date, True, True, inv / 2, inv / 2, inv)


def test_database_schema_and_statements(tmp_path):
    with sc.DataBase(tmp_path) as db:
        assert os.path.exists(tmp_path / "approximation.db")  # utils/source/db.cpp:10
        assert db.get_status("2019-05-22") == sc.CloudShadowStatus()  # unknown date: nothing computed
        db.write_detection_result("2019-05-22", True, False, 0.1, 0.0, 0.1)
        assert db.get_status("2019-05-22") == sc.CloudShadowStatus(True, False, 0.1)
        db.write_detection_result("2019-05-22", True, True, 0.1, 0.2, 0.3)  # upsert
        assert db.get_status("2019-5-22") == sc.CloudShadowStatus(True, True, 0.3)
        assert db.get_approx_status("2019-05-22", sc.ApproxMethod.Laplace) == {}  # creates the table on demand
        a = db.write_approx_results("2019-05-22", "B04", sc.ApproxMethod.Laplace)
        b = db.write_approx_results("2019-05-22", "B08", sc.ApproxMethod.Laplace)
        c = db.write_approx_results("2019-05-22", "B04", sc.ApproxMethod.Poisson)
        d = db.write_approx_results("2019-05-22", "B04", sc.ApproxMethod.Laplace)  # a second row, not a replacement
        assert (a, b, c, d) == (1, 2, 3, 4)
        assert db.get_approx_status("2019-05-22", sc.ApproxMethod.Laplace) == {"B04": 1, "B08": 2}
        assert db.get_approx_status("2019-05-22", sc.ApproxMethod.Poisson) == {"B04": 3}
        assert db.get_approx_status("2019-05-23", sc.ApproxMethod.Laplace) == {}
    # the file is what the reference's SQLiteCpp code expects: same tables, columns and method strings
    con = sqlite3.connect(tmp_path / "approximation.db")
    cols = [r[1] for r in con.execute("PRAGMA table_info(dates)")]
    assert cols == ["year", "month", "day", "clouds_computed", "shadows_computed", "percent_cloudy", "percent_shadows",
                    "percent_invalid"]  # fmt: skip
    assert [r[1] for r in con.execute("PRAGMA table_info(approximated_data)")] == ["id", "band_name", "method", "year",
                                                                                   "month", "day"]  # fmt: skip
    assert con.execute("SELECT band_name, method, year, month, day FROM approximated_data WHERE id=3").fetchone() == (
        "B04", "Poisson", 2019, 5, 22)  # fmt: skip
    con.close()
    with sc.DataBase(tmp_path) as db:  # reopening keeps everything
        assert db.get_status("2019-05-22").percent_invalid == 0.3


def test_select_close_images_window_and_quirk(tmp_path):
    with sc.DataBase(tmp_path) as db:
        _fill_dates(db, [("2019-03-31", 0.1), ("2019-04-01", 0.2), ("2019-05-22", 0.3), ("2019-06-30", 0.4),
                         ("2019-07-01", 0.5), ("2018-05-01", 0.6), ("2019-12-15", 0.7), ("2020-01-10", 0.8),
                         ("2020-12-20", 0.9), ("2020-02-28", 0.15), ("2019-01-05", 0.25)])  # fmt: skip
        got = [(i.date.isoformat(), i.percent_invalid) for i in db.select_close_images("2019-05-22")]
        assert got == [("2019-04-01", 0.2), ("2019-06-30", 0.4)]  # previous, same (minus the day itself) and next month
        # year wrap: years {2020, 2020, 2019} x months {1, 2, 12}: the reference's SQL tests them independently
        # (db.cpp:105-112), so December 2020 and January 2019 match a January-2020 date as well
        got = [i.date.isoformat() for i in db.select_close_images("2020-01-10")]
        assert got == ["2019-01-05", "2019-12-15", "2020-02-28", "2020-12-20"]
        assert math.isnan(db.select_info_about_date("2021-01-01").percent_invalid)
        assert db.select_info_about_date("2019-05-22").percent_invalid == 0.3


def test_find_good_close_image(tmp_path):
    with sc.DataBase(tmp_path) as db:
        for w in (-0.01, 1.01):
            with pytest.raises(sc.GenericError):  # poisson.cpp:325-327
                sc.find_good_close_image("2019-05-22", w, db)
        assert sc.find_good_close_image("2019-05-22", 0.5, db) == ""  # no neighbours at all (poisson.cpp:331-334)
        _fill_dates(db, [("2019-05-22", 0.30), ("2019-05-20", 0.90), ("2019-05-02", 0.05), ("2019-06-11", 0.10),
                         ("2019-04-30", 0.0)])  # fmt: skip
        # weight 1: days only -> the 20th, but it is cloudier than the date itself -> keep the date (use Laplace)
        assert sc.find_good_close_image("2019-05-22", 1.0, db) == "2019-05-22"
        # weight 0: invalid fraction only -> April 30th (0 % invalid)
        assert sc.find_good_close_image("2019-05-22", 0.0, db) == "2019-04-30"
        # in between: 0.01*days + 0.99*invalid -> 20th: .911, 2nd: .2495, Jun 11: .299, Apr 30: .22
        assert sc.find_good_close_image("2019-05-22", 0.01, db) == "2019-04-30"
        # 0.05: 20th .955, 2nd 1.0475, Jun 11 1.095, Apr 30 1.1 -> 20th is best but cloudier than the date
        assert sc.find_good_close_image("2019-05-22", 0.05, db) == "2019-05-22"
        # a date that is not in the table itself: the neighbour wins (NaN compares false)
        assert sc.find_good_close_image("2019-05-21", 1.0, db) == "2019-05-20"
        # simple-string input, ISO output (to_iso_extended_string, poisson.cpp:346)
        assert sc.find_good_close_image("2019-May-23", 0.0, db) == "2019-04-30"


def test_find_directory_contents(tmp_path):
    (tmp_path / "2019-05-22").mkdir()
    (tmp_path / "2019-05-22" / "B04.tif").write_bytes(b"")
    (tmp_path / "2019-05-23").mkdir()
    (tmp_path / "notes").mkdir()
    (tmp_path / "2019-05-223").mkdir()
    assert sc.find_directory_contents(tmp_path / "2019-05-22") == sc.DirectoryContents.MultiSpectral
    assert sc.find_directory_contents(str(tmp_path / "2019-05-22") + "/") == sc.DirectoryContents.MultiSpectral
    assert sc.find_directory_contents(tmp_path / "2019-05-23") == sc.DirectoryContents.Radar
    assert sc.find_directory_contents(tmp_path / "notes") == sc.DirectoryContents.NoSatelliteData
    assert sc.find_directory_contents(tmp_path / "2019-05-223") == sc.DirectoryContents.NoSatelliteData  # regex_match


GEO = {gt.T_PIXEL_SCALE: (12, [10.0, 10.0, 0.0]), gt.T_TIEPOINT: (12, [0.0, 0.0, 0.0, 5e5, 6e6, 0.0])}


def make_scene_tree(base, rows=40, cols=56):
    """Three date folders with B04/B08 (u16), cloud and shadow masks (u8); a radar folder; a non-date folder."""
    truth = {}
    for k, (name, cover) in enumerate([("2019-05-22", 0.30), ("2019-05-12", 0.10), ("2019-06-01", 0.55)]):
        d = base / name
        d.mkdir()
        clouds = synth.blob_mask(rows, cols, cover=cover * 0.7, sigma=3.0, seed=10 + k)
        shadows = synth.blob_mask(rows, cols, cover=cover * 0.5, sigma=3.0, seed=20 + k)
        bands = {}
        for j, b in enumerate(("B04", "B08")):
            img = np.round(synth.smooth_band(rows, cols, seed=30 + 2 * k + j)).astype(np.uint16)
            gt.write_tiff(d / f"{b}.tif", [img], extra_tags=GEO, compress=True)
            bands[b] = img
        gt.write_tiff(d / "cloud_mask.tif", [clouds.astype(np.uint8) * 255], extra_tags=GEO)
        gt.write_tiff(d / "shadow_mask.tif", [shadows.astype(np.uint8)], extra_tags=GEO)
        truth[name] = (bands, clouds | shadows)
    (base / "2019-05-30").mkdir()  # no B04.tif: radar
    (base / "logs").mkdir()
    return truth


def oracle_fill(port):
    def fill(bands, mask):
        for b in bands:
            b[...] = port.laplace_fill(b, mask, mode=1)[0]

    return fill


def test_fill_missing_data_folder(tmp_path, port):
    truth = make_scene_tree(tmp_path)
    with sc.DataBase(tmp_path) as db:
        for name, (_, mask) in truth.items():
            db.write_detection_result(name, True, name != "2019-05-12", 0.0, 0.0, float(mask.mean()))
    calls = []
    fill = oracle_fill(port)

    def counting_fill(bands, mask):
        calls.append(len(bands))
        fill(bands, mask)

    done = sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], use_cache=True, skip_threshold=0.5, fill=counting_fill)
    # 05-12: shadows missing -> skipped; 06-01: more than 50 % invalid -> skipped; radar / non-date folders ignored
    assert done == {"2019-05-22": {"B04": 1, "B08": 2}} and calls == [2]  # both bands in ONE batched call
    bands, mask = truth["2019-05-22"]
    for b, id_ in done["2019-05-22"].items():
        out = gt.GeoTIFF(tmp_path / "2019-05-22" / "approximated_data" / f"{b}_{id_}.tif", np.float64)
        got = out.read(1)
        assert out.file.dtype == np.uint16 and out.geo_transform == (5e5, 10.0, 0.0, 6e6, 0.0, -10.0)
        assert np.array_equal(got[~mask], bands[b][~mask])  # known pixels untouched
        want = port.laplace_fill(bands[b].astype(np.float64), mask, mode=1)[0]
        assert np.array_equal(got, gt.gdal_convert(want, np.uint16))
    # cached now: nothing to do; without the cache it is filled again under new ids
    assert sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], True, 0.5, fill=counting_fill) == {} and calls == [2]
    again = sc.fill_missing_data_folder(tmp_path, ["B04"], False, 0.5, write_outputs=False, fill=counting_fill)
    assert again == {"2019-05-22": {"B04": 3}} and calls == [2, 1]
    assert not os.path.exists(tmp_path / "2019-05-22" / "approximated_data" / "B04_3.tif")
    assert sc.fill_missing_data_folder(tmp_path / "nope", ["B04"], True, 0.5, fill=counting_fill) == {}


def test_blend_missing_data_folder(tmp_path, port):
    truth = make_scene_tree(tmp_path)
    with sc.DataBase(tmp_path) as db:
        for name, (_, mask) in truth.items():
            db.write_detection_result(name, True, True, 0.0, 0.0, float(mask.mean()))
    log = []

    def blend(bands, guides, mask):
        log.append(("poisson", len(bands)))
        out = port.poisson_blend(bands, guides, mask, tol=1e-10, max_it=100000)[0]
        for b, o in zip(bands, out):
            b[...] = o
        return True

    def fill(bands, mask):
        log.append(("laplace", len(bands)))
        oracle_fill(port)(bands, mask)

    done = sc.blend_missing_data_folder(tmp_path, ["B04", "B08"], True, 0.9, distance_weight=0.0, blend=blend, fill=fill)
    # weight 0 -> cleanest neighbour.  05-12 is the cleanest of all: it keeps itself -> Laplace.  The other two take 05-12.
    assert log == [("laplace", 2), ("poisson", 2), ("poisson", 2)]
    assert set(done) == {"2019-05-12", "2019-05-22", "2019-06-01"}
    with sc.DataBase(tmp_path) as db:
        assert set(db.get_approx_status("2019-05-12", sc.ApproxMethod.Laplace)) == {"B04", "B08"}
        assert set(db.get_approx_status("2019-05-22", sc.ApproxMethod.Poisson)) == {"B04", "B08"}
        assert db.get_approx_status("2019-05-22", sc.ApproxMethod.Laplace) == {}
    bands, mask = truth["2019-05-22"]
    guide = truth["2019-05-12"][0]
    want = port.poisson_blend([bands["B04"].astype(np.float64)], [guide["B04"].astype(np.float64)], mask, tol=1e-10,
                              max_it=100000)[0][0]  # fmt: skip
    id_ = done["2019-05-22"]["B04"]
    got = gt.TiffFile(tmp_path / "2019-05-22" / "approximated_data" / f"B04_{id_}.tif").read_band(1)
    assert np.array_equal(got, gt.gdal_convert(want, np.uint16))
    assert np.array_equal(got[~mask], bands["B04"][~mask])
    # a solver failure leaves the folder unrecorded (poisson.cpp:263-269: log and leave the image alone)
    done = sc.blend_missing_data_folder(tmp_path, ["B04"], False, 0.9, 0.0, blend=lambda *a: False, fill=fill)
    assert set(done) == {"2019-05-12"}


def test_folder_driver_shards_folders_over_ranks(tmp_path, port, monkeypatch):
    """scenes are independent -> every world-th folder per rank, one shared database, no collective."""
    truth = make_scene_tree(tmp_path)
    with sc.DataBase(tmp_path) as db:
        for name, (_, mask) in truth.items():
            db.write_detection_result(name, True, True, 0.0, 0.0, float(mask.mean()))
    fill = oracle_fill(port)
    parts = [sc.fill_missing_data_folder(tmp_path, ["B04"], True, 1.0, fill=fill, shard=(r, 2)) for r in range(2)]
    assert set(parts[0]) == {"2019-05-12", "2019-06-01"} and set(parts[1]) == {"2019-05-22"}  # sorted, every 2nd
    ids = sorted(i for p in parts for d in p.values() for i in d.values())
    assert ids == [1, 2, 3]  # one id sequence: the ranks share approximation.db
    with pytest.raises(ValueError):
        sc.fill_missing_data_folder(tmp_path, ["B04"], True, 1.0, fill=fill, shard=(2, 2))
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    assert sc.shard_from_env() == (3, 8)


def test_folder_driver_prefetch_overlaps_and_matches_serial(tmp_path, port, monkeypatch):
    import threading

    truth = make_scene_tree(tmp_path)
    with sc.DataBase(tmp_path) as db:
        for name, (_, mask) in truth.items():
            db.write_detection_result(name, True, True, 0.0, 0.0, float(mask.mean()))
    fill = oracle_fill(port)
    loading = {}
    real_load = sc._load_job

    def load(job):
        loading.setdefault(job.name, threading.Event()).set()
        return real_load(job), threading.current_thread().name

    monkeypatch.setattr(sc, "_load_job", lambda job: load(job)[0])
    overlapped = []

    def slow_fill(bands, mask):
        # while folder k is being solved, the reader thread must already be decoding folder k + 1
        nxt = {"2019-05-12": "2019-05-22", "2019-05-22": "2019-06-01"}
        cur = [n for n, (_, m) in truth.items() if m.shape == mask.shape and np.array_equal(m, mask)][0]
        if cur in nxt:
            overlapped.append(loading.setdefault(nxt[cur], threading.Event()).wait(timeout=20))
        fill(bands, mask)

    a = sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], False, 1.0, fill=slow_fill, prefetch=True)
    assert overlapped == [True, True]
    b = sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], False, 1.0, fill=fill, prefetch=False)
    assert list(a) == list(b) == ["2019-05-12", "2019-05-22", "2019-06-01"]
    for name in a:
        for band in ("B04", "B08"):
            x = gt.TiffFile(tmp_path / name / "approximated_data" / f"{band}_{a[name][band]}.tif").read_band(1)
            y = gt.TiffFile(tmp_path / name / "approximated_data" / f"{band}_{b[name][band]}.tif").read_band(1)
            assert np.array_equal(x, y)
    # a folder that cannot be decoded: the error surfaces, earlier folders are recorded, helper threads are gone
    (tmp_path / "2019-05-22" / "B08.tif").write_bytes(b"II*\0garbage")
    with pytest.raises(gt.TiffError):
        sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], False, 1.0, fill=fill)
    with sc.DataBase(tmp_path) as db:
        assert len(db.get_approx_status("2019-05-12", sc.ApproxMethod.Laplace)) == 2
    assert not [t for t in threading.enumerate() if t.name.startswith("satfill-")]


def test_two_ranks_share_one_database(tmp_path):
    """Two processes (RANK 0 / 1 of WORLD_SIZE 2, as torchrun would start them) run the folder driver at the same time on
    one base folder: every folder is filled exactly once, and the ids they draw from the shared approximation.db are
    distinct (SQLite's file lock serialises the writers)."""
    import json
    import subprocess
    import sys

    truth = {}
    for k in range(6):
        name = f"2019-05-{10 + k:02d}"
        d = tmp_path / name
        d.mkdir()
        mask = synth.blob_mask(40, 48, cover=0.25, sigma=3.0, seed=70 + k)
        for j, b in enumerate(("B04", "B08")):
            gt.write_tiff(d / f"{b}.tif", [np.round(synth.smooth_band(40, 48, seed=80 + 2 * k + j)).astype(np.uint16)],
                          extra_tags=GEO)  # fmt: skip
        gt.write_tiff(d / "cloud_mask.tif", [mask.astype(np.uint8)], extra_tags=GEO)
        gt.write_tiff(d / "shadow_mask.tif", [np.zeros((40, 48), np.uint8)], extra_tags=GEO)
        truth[name] = mask
    with sc.DataBase(tmp_path) as db:
        for name, mask in truth.items():
            db.write_detection_result(name, True, True, 0.0, 0.0, float(mask.mean()))
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "folder_worker.py")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, worker, str(tmp_path), str(tmp_path / f"done{r}.json")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))  # fmt: skip
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
    done = [json.load(open(tmp_path / f"done{r}.json")) for r in range(2)]
    assert sorted(done[0]) == sorted(truth)[0::2] and sorted(done[1]) == sorted(truth)[1::2]
    ids = sorted(i for d in done for f in d.values() for i in f.values())
    assert ids == list(range(1, 13))
    with sc.DataBase(tmp_path) as db:
        for name in truth:
            assert set(db.get_approx_status(name, sc.ApproxMethod.Laplace)) == {"B04", "B08"}
    for d in done:
        for name, bands in d.items():
            for b, id_ in bands.items():
                assert os.path.exists(tmp_path / name / "approximated_data" / f"{b}_{id_}.tif")
