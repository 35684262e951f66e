"""Scene-level orchestration around the fill path: the SQLite bookkeeping of the reference
(`utils::DataBase`, lib/utils/source/db.cpp:8-45; `approx::DataBase`, lib/approx/source/db.cpp:12-156), the guidance-date
picker `find_good_close_image` (lib/approx/source/poisson.cpp:323-349), `find_directory_contents`
(lib/utils/source/filesystem.cpp:3-15) and the folder driver the reference keeps commented out
(`fill_missing_data_folder`, lib/approx/source/laplace.cpp:170-244).

Pure host logic: the database file, its two tables and every SQL statement are the reference's, so a folder prepared by the
reference's cloud detection (`approximation.db`, table `dates`) is read as is and `approximated_data` rows written here are
read by the reference.  The pixels go through the GPU fill (satellite_approximation_b200.default_context()); there is no
CPU solve in this module."""
from __future__ import annotations

import datetime as _dt
import enum
import logging
import os
import re
import sqlite3
import threading
from dataclasses import dataclass
from typing import Callable, Optional, Sequence

import numpy as np

__all__ = ["Date", "parse_simple_date", "CloudShadowStatus", "DayInfo", "ApproxMethod", "DataBase", "GenericError",
           "find_good_close_image", "DirectoryContents", "find_directory_contents", "fill_missing_data_folder", "shard_from_env",
           "blend_missing_data_folder"]  # fmt: skip

_log = logging.getLogger("approx")

_MONTHS = {m: i + 1 for i, m in enumerate(["jan", "feb", "mar", "apr", "may", "jun", "jul", "aug", "sep", "oct", "nov", "dec"])}
_MONTHS.update({m: i + 1 for i, m in enumerate(["january", "february", "march", "april", "may", "june", "july", "august",
                                                "september", "october", "november", "december"])})  # fmt: skip


class GenericError(RuntimeError):
    """utils::GenericError (lib/utils/include/utils/error.h:11-20)."""


def parse_simple_date(text: str) -> _dt.date:
    """boost::gregorian::from_simple_string: year, month, day separated by '-', '/', ',' or blanks; the month is a number
    or an English month name ("2002-1-25", "2002-Jan-25").  Malformed input raises ValueError (Boost throws
    bad_lexical_cast / bad_month / bad_day_of_month)."""
    parts = [p for p in re.split(r"[-/,\s]+", text.strip()) if p]
    if len(parts) != 3:
        raise ValueError(f"not a year-month-day date: {text!r}")
    y, m, d = parts
    try:
        month = _MONTHS[m.lower()] if m.lower() in _MONTHS else int(m)
        return _dt.date(int(y), month, int(d))
    except (ValueError, KeyError) as e:
        raise ValueError(f"not a year-month-day date: {text!r}") from e


def _add_months(d: _dt.date, n: int) -> tuple[int, int]:
    """(year, month) of `d + boost::gregorian::months(n)`; the day (Boost snaps to the month's end) is never used."""
    k = d.year * 12 + (d.month - 1) + n
    return k // 12, k % 12 + 1


@dataclass(frozen=True, order=False)
class Date:
    """utils::Date (lib/utils/include/utils/date.h:11-27, source/date.cpp)."""

    year: int = 0
    month: int = 0
    day: int = 0

    @staticmethod
    def parse(text: str) -> "Date":
        d = parse_simple_date(text)
        return Date(d.year, d.month, d.day)

    def __lt__(self, other: "Date") -> bool:
        return (self.year, self.month, self.day) < (other.year, other.month, other.day)

    def __str__(self) -> str:  # operator<< (date.cpp:33-36)
        return f"{self.year}-{self.month:02d}-{self.day:02d}"

    def sql(self) -> tuple[int, int, int]:  # bind_sql (date.cpp:38-46)
        return (self.year, self.month, self.day)


@dataclass
class CloudShadowStatus:
    """utils::CloudShadowStatus (lib/utils/include/utils/db.h:13-17)."""

    clouds_exist: bool = False
    shadows_exist: bool = False
    percent_invalid: float = 0.0


@dataclass
class DayInfo:
    """approx::DayInfo (lib/approx/include/approx/db.h:12-17)."""

    date: _dt.date
    percent_invalid: float

    def distance(self, other: _dt.date, weight: float) -> float:  # db.cpp:12-16
        return weight * float(abs((other - self.date).days)) + (1 - weight) * self.percent_invalid


class ApproxMethod(enum.Enum):
    """approx::ApproxMethod (db.h:19-22); stored by name, as magic_enum::enum_name does (db.cpp:42,51)."""

    Laplace = 0
    Poisson = 1


_CREATE_DATES = """
CREATE TABLE IF NOT EXISTS dates(
    year INTEGER NOT NULL,
    month INTEGER NOT NULL,
    day INTEGER NOT NULL,
    clouds_computed INTEGER,
    shadows_computed INTEGER,
    percent_cloudy REAL,
    percent_shadows REAL,
    percent_invalid REAL,
    PRIMARY KEY(year, month, day));
"""

_CREATE_APPROX = """
CREATE TABLE IF NOT EXISTS approximated_data(
    id INTEGER PRIMARY KEY AUTOINCREMENT,
    band_name TEXT,
    method TEXT,
    year INTEGER NOT NULL,
    month INTEGER NOT NULL,
    day INTEGER NOT NULL,
    FOREIGN KEY(year, month, day) REFERENCES dates(year, month, day));
"""


class DataBase:
    """`approx::DataBase` on top of `utils::DataBase`: `<base_path>/approximation.db`, tables `dates` (written by the
    reference's cloud detection) and `approximated_data`.  One connection, serialised by a lock (the reference's
    commented-out driver guards every call with one mutex, laplace.cpp:187-236)."""

    def __init__(self, base_path):
        self.path = os.path.join(os.fspath(base_path), "approximation.db")
        # SQLite::OPEN_CREATE | OPEN_READWRITE (utils/source/db.cpp:10): the directory has to exist.  Several ranks may share
        # the file (one process per GPU, folders dealt out round-robin): SQLite's file lock serialises them, hence the timeout
        self._db = sqlite3.connect(self.path, check_same_thread=False, isolation_level=None, timeout=60.0)
        self._lock = threading.Lock()
        self._db.execute(_CREATE_DATES)

    def close(self) -> None:
        self._db.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- utils::DataBase ------------------------------------------------------------------------------------------------
    def get_status(self, date_string: str) -> CloudShadowStatus:
        """utils/source/db.cpp:16-28.  A date that is not in the table: the reference falls off the end of a non-void
        function (undefined behaviour); here the default-constructed status (nothing computed) comes back."""
        with self._lock:
            row = self._db.execute(
                "SELECT clouds_computed, shadows_computed, percent_invalid FROM dates WHERE year=? AND month=? AND day=?;",
                Date.parse(date_string).sql()).fetchone()  # fmt: skip
        if row is None:
            return CloudShadowStatus()
        return CloudShadowStatus(bool(row[0] or 0), bool(row[1] or 0), float(row[2] or 0.0))

    def write_detection_result(self, date_string: str, clouds_computed: bool, shadows_computed: bool,
                               percent_cloudy: float, percent_shadows: float, percent_invalid: float) -> None:  # fmt: skip
        """The upsert the reference's cloud detection issues (lib/cloud_shadow_detection/source/db.cpp:45-66) -- here so
        that a `dates` table can be produced without that (out-of-scope) subsystem."""
        with self._lock:
            self._db.execute(
                """INSERT INTO dates (year, month, day, clouds_computed, shadows_computed, percent_cloudy, percent_shadows,
                   percent_invalid) VALUES(?, ?, ?, ?, ?, ?, ?, ?)
                   ON CONFLICT(year, month, day) DO UPDATE SET clouds_computed = excluded.clouds_computed,
                   shadows_computed = excluded.shadows_computed, percent_cloudy = excluded.percent_cloudy,
                   percent_shadows = excluded.percent_shadows, percent_invalid = excluded.percent_invalid;""",
                Date.parse(date_string).sql() + (int(clouds_computed), int(shadows_computed), float(percent_cloudy),
                                                 float(percent_shadows), float(percent_invalid)))  # fmt: skip

    # -- approx::DataBase -----------------------------------------------------------------------------------------------
    def write_approx_results(self, date_string: str, band_name: str, method: ApproxMethod) -> int:
        """db.cpp:38-64: one row per call (the table has no uniqueness constraint besides the id, so INSERT OR REPLACE
        always inserts); returns the new id."""
        with self._lock:
            self._db.execute(_CREATE_APPROX)
            cur = self._db.execute(
                "INSERT OR REPLACE INTO approximated_data (band_name, method, year, month, day) VALUES(?, ?, ?, ?, ?)",
                (band_name, method.name) + Date.parse(date_string).sql())  # fmt: skip
            return int(cur.lastrowid)

    def get_approx_status(self, date_string: str, method: ApproxMethod) -> dict[str, int]:
        """db.cpp:66-95: band name -> id for the date and method (unordered_map::emplace keeps the FIRST id of a band)."""
        with self._lock:
            self._db.execute(_CREATE_APPROX)
            rows = self._db.execute(
                "SELECT id, band_name FROM approximated_data WHERE method = ? AND year = ? AND month = ? AND day = ?;",
                (method.name,) + Date.parse(date_string).sql()).fetchall()  # fmt: skip
        out: dict[str, int] = {}
        for id_, name in rows:
            out.setdefault(name, int(id_))
        return out

    def select_close_images(self, date_string: str) -> list[DayInfo]:
        """db.cpp:97-137: every other date whose YEAR is that of the date, of the date + 1 month or of the date - 1 month
        AND whose MONTH is one of those three months (the two tests are independent, as in the reference's SQL: for a
        January date this also matches December of the same year), ordered by date."""
        date = parse_simple_date(date_string)
        ny, nm = _add_months(date, 1)
        py, pm = _add_months(date, -1)
        with self._lock:
            rows = self._db.execute(
                """SELECT year, month, day, percent_invalid FROM dates WHERE
                   (year = ? OR year = ? OR year = ?) AND (month = ? OR month = ? OR month = ?) AND NOT
                   (year = ? AND month = ? AND day = ?) ORDER BY year, month, day""",
                (date.year, ny, py, date.month, nm, pm, date.year, date.month, date.day)).fetchall()  # fmt: skip
        return [DayInfo(_dt.date(y, m, d), float(p or 0.0)) for y, m, d, p in rows]

    def select_info_about_date(self, date_string: str) -> DayInfo:
        """db.cpp:139-156.  For a date that is not in the table the reference returns an uninitialised percent_invalid;
        here it is NaN (every `<` against it is false, so find_good_close_image keeps the neighbour it found)."""
        date = parse_simple_date(date_string)
        with self._lock:
            rows = self._db.execute(
                "SELECT percent_invalid FROM dates WHERE year = ? AND month = ? AND day = ? ORDER BY year, month, day",
                (date.year, date.month, date.day)).fetchall()  # fmt: skip
        info = DayInfo(_dt.date.min, float("nan"))  # boost's default date is not_a_date_time
        for (p,) in rows:
            info.percent_invalid = float(p or 0.0)
        return info


def find_good_close_image(date_string: str, distance_weight: float, db: DataBase) -> str:
    """poisson.cpp:323-349: the date (ISO, YYYY-MM-DD) of the neighbouring scene that minimises
    `w * |days apart| + (1 - w) * percent_invalid`; `date_string` itself when the scene of that date has fewer invalid
    pixels than the best neighbour (fill it with Laplace instead); "" when there is no neighbour.  GenericError when the
    weight is outside [0, 1].  Ties keep date order (the reference's std::sort leaves them unspecified)."""
    if distance_weight < 0 or distance_weight > 1:
        raise GenericError("Could not find close image: distance weight not between 0 and 1")
    date = parse_simple_date(date_string)
    info = db.select_close_images(date_string)
    if not info:
        _log.warning("Could not find any good images close by. Date: %s", date.strftime("%Y-%b-%d"))
        return ""
    info.sort(key=lambda i: i.distance(date, distance_weight))
    current = db.select_info_about_date(date_string)
    if current.percent_invalid < info[0].percent_invalid:
        _log.debug("The current date has fewer invalid pixels than the date we found. Use laplace approximation")
        return date_string
    _log.debug("Found image: %s %.2f%% invalid", info[0].date.isoformat(), 100 * info[0].percent_invalid)
    return info[0].date.isoformat()


class DirectoryContents(enum.Enum):
    """utils::DirectoryContents (lib/utils/include/utils/filesystem.h:7-11)."""

    NoSatelliteData = 0
    MultiSpectral = 1
    Radar = 2


_DATE_DIR = re.compile(r"\d{4}-\d{2}-\d{2}")


def find_directory_contents(path) -> DirectoryContents:
    """filesystem.cpp:3-15: a folder named YYYY-MM-DD holds multispectral data when it has a B04.tif, else radar."""
    path = os.fspath(path)
    if not _DATE_DIR.fullmatch(os.path.basename(os.path.normpath(path))):
        return DirectoryContents.NoSatelliteData
    return DirectoryContents.MultiSpectral if os.path.exists(os.path.join(path, "B04.tif")) else DirectoryContents.Radar


def shard_from_env() -> tuple[int, int]:
    """(RANK, WORLD_SIZE) as torchrun exports them; (0, 1) for a plain process."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _my_folders(base_folder: str, shard: tuple[int, int]) -> list[str]:
    rank, world = shard
    if not (world >= 1 and 0 <= rank < world):
        raise ValueError(f"shard: rank {rank} of {world}")
    folders = sorted(e.path for e in os.scandir(base_folder)
                     if e.is_dir() and find_directory_contents(e.path) == DirectoryContents.MultiSpectral)  # fmt: skip
    return folders[rank::world]  # = multi.round_robin(len(folders), world, rank)


def _read_scene_mask(folder: str, status: CloudShadowStatus) -> np.ndarray:
    from . import geotiff

    clouds = geotiff.GeoTIFF(os.path.join(folder, "cloud_mask.tif"), np.uint8).read(1) != 0
    if status.shadows_exist:
        shadows = geotiff.GeoTIFF(os.path.join(folder, "shadow_mask.tif"), np.uint8).read(1) != 0
    else:
        shadows = np.zeros_like(clouds)
    return clouds | shadows


def _gpu_laplace(bands: list[np.ndarray], mask: np.ndarray) -> None:
    import satellite_approximation_b200 as sab

    sab.default_context().laplace_fill(bands, mask, tolerance=sab._defaults["laplace_tolerance"],
                                       max_iterations=sab._defaults["laplace_max_iterations"],
                                       precond=sab._defaults["precond"], check_every=sab._defaults["check_every"])  # fmt: skip


def _gpu_poisson(bands: list[np.ndarray], guidance: list[np.ndarray], mask: np.ndarray) -> bool:
    import satellite_approximation_b200 as sab

    stats = sab.default_context().poisson_blend(bands, guidance, mask, tolerance=1e-6, precond=sab._defaults["precond"],
                                                check_every=sab._defaults["check_every"])  # fmt: skip
    return all(s["status"] != sab.SA_NOT_CONVERGED for s in stats)


@dataclass
class _Job:
    folder: str
    name: str
    status: CloudShadowStatus
    todo: list
    method: ApproxMethod
    guide_dir: Optional[str] = None


def _load_job(job: _Job):
    """Decode one folder: mask, the bands to fill, and (Poisson) the guidance date's bands.  Runs on the reader thread
    (zlib and numpy release the GIL, so this overlaps the GPU solve of the previous folder)."""
    from . import geotiff

    def band(folder, b):
        return np.ascontiguousarray(geotiff.GeoTIFF(os.path.join(folder, f"{b}.tif"), np.float64).read(1))

    mask = _read_scene_mask(job.folder, job.status)
    bands = [band(job.folder, b) for b in job.todo]
    guides = [band(job.guide_dir, b) for b in job.todo] if job.guide_dir is not None else None
    return mask, bands, guides


def _run_jobs(jobs: list, db: "DataBase", write_outputs: bool, prefetch: bool, blend, fill) -> dict[str, dict[str, int]]:
    """Read -> solve -> record -> write, folder after folder.  With `prefetch` the next folder is decoded and the
    previous folder's result files are encoded on two helper threads while the GPU solves the current one; the database
    is only touched from the calling thread."""
    from concurrent.futures import ThreadPoolExecutor

    from . import geotiff

    done: dict[str, dict[str, int]] = {}
    if not jobs:
        return done
    reader = ThreadPoolExecutor(1, thread_name_prefix="satfill-read") if prefetch else None
    writer = ThreadPoolExecutor(1, thread_name_prefix="satfill-write") if prefetch else None
    writes = []
    try:
        nxt = reader.submit(_load_job, jobs[0]) if reader else None
        for i, job in enumerate(jobs):
            _log.debug("Starting folder: %s", job.folder)
            mask, bands, guides = nxt.result() if reader else _load_job(job)
            if reader and i + 1 < len(jobs):
                nxt = reader.submit(_load_job, jobs[i + 1])
            if any(b.shape != mask.shape for b in bands):
                raise RuntimeError("Input image and mask need to be the same size")  # laplace.cpp:124-127
            if guides is not None:
                if any(g.shape != mask.shape for g in guides):
                    _log.error("Input and replacement images must have the same dimensions")  # poisson.cpp:154-157
                    continue
                if not blend(bands, guides, mask):
                    _log.error("Failed to solve the linear system (no convergence)")  # poisson.cpp:263-269
                    continue
            else:
                fill(bands, mask)
            out_dir = os.path.join(job.folder, "approximated_data")
            if write_outputs and not os.path.exists(out_dir):
                _log.info("Creating directory: %s", out_dir)
                os.makedirs(out_dir, exist_ok=True)
            done[job.name] = {}
            for b, values in zip(job.todo, bands):
                id_ = db.write_approx_results(job.name, b, job.method)
                done[job.name][b] = id_
                if write_outputs:
                    w = geotiff.GeoTiffWriter(values, os.path.join(job.folder, f"{b}.tif"))
                    dest = os.path.join(out_dir, f"{b}_{id_}.tif")
                    if writer:
                        writes.append(writer.submit(w.write, dest))
                    else:
                        w.write(dest)
            _log.info("Finished folder: %s", job.folder)
        for f in writes:
            f.result()  # re-raise what a write raised
    finally:
        for ex in (reader, writer):
            if ex is not None:
                ex.shutdown(wait=True, cancel_futures=True)
    return done


def _eligible(db: "DataBase", folder: str, skip_threshold: float) -> Optional[CloudShadowStatus]:
    status = db.get_status(os.path.basename(folder))
    if not (status.clouds_exist and status.shadows_exist):
        _log.warning("Both clouds and shadows don't exist for folder %s. Skipping", folder)
        return None
    if status.percent_invalid > skip_threshold:
        _log.info("Skipping %s because there is too little valid data (%.1f%% invalid)", folder,
                  status.percent_invalid * 100.0)  # fmt: skip
        return None
    return status


def fill_missing_data_folder(base_folder, band_names: Sequence[str], use_cache: bool, skip_threshold: float,
                             write_outputs: bool = True,
                             fill: Optional[Callable[[list[np.ndarray], np.ndarray], None]] = None,
                             shard: tuple[int, int] = (0, 1), prefetch: bool = True) -> dict[str, dict[str, int]]:  # fmt: skip
    """The folder driver the reference keeps commented out (laplace.cpp:170-244): for every multispectral date folder
    under `base_folder` whose cloud AND shadow masks exist and whose invalid fraction is at most `skip_threshold`, fill
    the invalid pixels (clouds | shadows) of each band `<folder>/<band>.tif` with the Laplace fill, record the result in
    `approximated_data`, and store it as `<folder>/approximated_data/<band>_<id>.tif` (the reference's write is commented
    out inside the commented-out driver; `write_outputs=False` reproduces that).  With `use_cache`, bands that already
    have a Laplace row for the date are skipped.

    Differences by design: all bands of a folder share the mask, so they go to the GPU as ONE batched solve (the
    reference re-assembles per band); bands are read in raster layout (geotiff.py); with `prefetch` the TIFF decode of
    the next folder and the encode of the previous one overlap the solve (the reference's sketch is a serial
    std::for_each under one mutex).  `fill(bands, mask)` fills float64 C-ordered bands in place; the default is the GPU
    path and there is no other implementation in the product (the parameter exists so that the host logic can be tested
    on a machine without a GPU).  `shard=(rank, world)` makes this process take every world-th folder (SURVEY.md §8e:
    scenes are independent, one process per GPU, no collective; the ranks share the database file).  Returns
    {folder name: {band: id}} for what was filled."""
    base_folder = os.fspath(base_folder)
    _log.debug("Processing directory: %s", base_folder)
    if not os.path.isdir(base_folder):
        _log.warning("Could not process: base folder is not a directory (%s)", base_folder)
        return {}
    with DataBase(base_folder) as db:
        jobs = []
        for folder in _my_folders(base_folder, shard):
            status = _eligible(db, folder, skip_threshold)
            if status is None:
                continue
            name = os.path.basename(folder)
            existing = db.get_approx_status(name, ApproxMethod.Laplace)
            todo = [b for b in band_names if not (use_cache and b in existing)]
            if todo:
                jobs.append(_Job(folder, name, status, todo, ApproxMethod.Laplace))
        return _run_jobs(jobs, db, write_outputs, prefetch, None, fill or _gpu_laplace)


def blend_missing_data_folder(base_folder, band_names: Sequence[str], use_cache: bool, skip_threshold: float,
                              distance_weight: float = 0.5, write_outputs: bool = True,
                              blend: Optional[Callable[[list[np.ndarray], list[np.ndarray], np.ndarray], bool]] = None,
                              fill: Optional[Callable[[list[np.ndarray], np.ndarray], None]] = None,
                              shard: tuple[int, int] = (0, 1), prefetch: bool = True) -> dict[str, dict[str, int]]:  # fmt: skip
    """The Poisson counterpart the reference's pieces imply (find_good_close_image + blend_images_poisson +
    ApproxMethod::Poisson, never wired together upstream): per date folder pick the guidance date with
    find_good_close_image; if it is another date, Poisson-blend each band against that date's band; if it is the date
    itself (it has fewer invalid pixels than any neighbour) or there is no neighbour, fall back to the Laplace fill and
    record it as such.  Same skipping, caching, batching, prefetching and output rules as fill_missing_data_folder."""
    base_folder = os.fspath(base_folder)
    if not os.path.isdir(base_folder):
        _log.warning("Could not process: base folder is not a directory (%s)", base_folder)
        return {}
    with DataBase(base_folder) as db:
        jobs = []
        for folder in _my_folders(base_folder, shard):
            status = _eligible(db, folder, skip_threshold)
            if status is None:
                continue
            name = os.path.basename(folder)
            close = find_good_close_image(name, distance_weight, db)
            guide_dir = os.path.join(base_folder, close) if close and close != name else None
            if guide_dir is not None and not all(os.path.exists(os.path.join(guide_dir, f"{b}.tif")) for b in band_names):
                _log.warning("Guidance date %s lacks some of the bands; using laplace approximation for %s", close, name)
                guide_dir = None
            method = ApproxMethod.Poisson if guide_dir is not None else ApproxMethod.Laplace
            existing = db.get_approx_status(name, method)
            todo = [b for b in band_names if not (use_cache and b in existing)]
            if todo:
                jobs.append(_Job(folder, name, status, todo, method, guide_dir))
        return _run_jobs(jobs, db, write_outputs, prefetch, blend or _gpu_poisson, fill or _gpu_laplace)
