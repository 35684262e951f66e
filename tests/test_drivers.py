"""CPU: host logic of the restated executables and image helpers (satellite_approximation_b200/drivers.py; reference
executables/laplace-main.cpp, executables/poisson-main.cpp, lib/approx/source/utils.cpp:16-68): argument and file checks,
image file round trips, and -- with the ORACLE standing in for the two GPU calls -- the band / layout / output-file
plumbing of poisson_main.  The real GPU run is tests/test_zz_gpu_drivers.py."""
from __future__ import annotations

import os

import numpy as np
import pytest

import satellite_approximation_b200 as sab
from satellite_approximation_b200 import drivers, synth
from satellite_approximation_b200 import geotiff as gt

GEO = {gt.T_PIXEL_SCALE: (12, [10.0, 10.0, 0.0]), gt.T_TIEPOINT: (12, [0.0, 0.0, 0.0, 5e5, 6e6, 0.0])}


def test_usage_and_missing_files(tmp_path):
    assert drivers.laplace_main([]) == -1 and drivers.laplace_main(["a", "b"]) == -1  # laplace-main.cpp:14-17
    assert drivers.laplace_main([str(tmp_path / "a.png"), str(tmp_path / "b.png"), str(tmp_path / "o.png")]) == -1
    assert drivers.poisson_main([]) == -1 and drivers.poisson_main(["x"]) == -1  # poisson-main.cpp:28-31
    assert drivers.poisson_main([str(tmp_path / "a.tif"), str(tmp_path / "b.tif")]) == -1
    assert drivers.main(["nothing"]) == -1


def test_saturate_u8_is_cv_convert_to(tmp_path):
    v = np.array([-5.0, -0.5, 0.5, 1.5, 2.5, 254.5, 255.5, 300.0, np.nan])
    assert drivers.saturate_u8(v).tolist() == [0, 0, 0, 2, 2, 254, 255, 255, 0]  # cvRound: half to even, then clamp
    cv2 = pytest.importorskip("cv2")
    # OpenCV itself: imwrite of a CV_64F matrix to PNG falls back to convertTo(CV_8U)
    a = np.random.default_rng(0).uniform(-20, 280, (13, 17, 3))
    a.ravel()[:6] = [0.5, 1.5, 2.5, 3.5, 254.5, 255.5]
    assert cv2.imwrite(str(tmp_path / "f.png"), a)
    assert np.array_equal(cv2.imread(str(tmp_path / "f.png"), cv2.IMREAD_UNCHANGED), drivers.saturate_u8(a))


def test_image_files_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    bgr = rng.integers(0, 256, (21, 34, 3), dtype=np.uint8)
    p = tmp_path / "a.png"
    assert drivers.imwrite(p, bgr)
    back = drivers.imread_color(p)
    assert back.dtype == np.uint8 and np.array_equal(back, bgr)
    ch = drivers.read_image(p)  # utils.cpp:16-34: R, G, B, pow(v / 255, 1 / 2.2)
    assert len(ch) == 3 and np.allclose(ch[0], (bgr[..., 2] / 255.0) ** (1 / 2.2), rtol=0, atol=0)
    q = tmp_path / "b.png"
    drivers.write_image(ch, q)  # utils.cpp:36-68: static_cast<uchar>(pow(v, 2.2) * 255) truncates
    again = drivers.imread_color(q)
    assert np.all(np.abs(again.astype(int) - bgr.astype(int)) <= 1) and np.all(again <= bgr)
    drivers.write_image(ch[:2], tmp_path / "c.png")  # not three channels: logged, nothing written
    assert not os.path.exists(tmp_path / "c.png")
    with pytest.raises(IOError):
        drivers.read_image(tmp_path / "missing.png")  # utils::IOError
    (tmp_path / "junk.png").write_bytes(b"junk")
    with pytest.raises(IOError):
        drivers.read_image(tmp_path / "junk.png")
    # a float matrix (what apply_laplace returns) is saturated on the way out, like cv::imwrite does for PNG
    f = bgr.astype(np.float64) + 0.25
    drivers.imwrite(p, f)
    assert np.array_equal(drivers.imread_color(p), bgr)


def make_pair(tmp_path, rows, cols, dtype=np.uint16):
    bands_in = [np.round(synth.smooth_band(rows, cols, seed=40 + b)).astype(dtype) for b in range(5)]
    bands_rp = [np.round(0.9 * synth.smooth_band(rows, cols, seed=50 + b) + 37).astype(dtype) for b in range(5)]
    cloud = synth.blob_mask(rows, cols, cover=0.25, sigma=4.0, seed=60, clear_border=False).astype(dtype) * 100
    a, b = tmp_path / "in" / "scene.tif", tmp_path / "rp" / "scene.tif"
    gt.write_tiff(a, bands_in + [cloud], extra_tags=GEO, compress=True)
    gt.write_tiff(b, bands_rp + [np.zeros_like(cloud)], extra_tags=GEO, tile=(16, 16))
    return a, b, bands_in, bands_rp, cloud


@pytest.mark.parametrize("layout", ["raster", "reference"])
def test_poisson_main_plumbing_with_oracle_pixels(tmp_path, monkeypatch, port, layout):
    import oracle

    rows, cols = 48, 36
    a, b, bands_in, bands_rp, cloud = make_pair(tmp_path, rows, cols)
    seen = {}

    def close(band):
        seen["band"] = band
        return oracle.morph_close_mask(np.asarray(band), 5)

    def blend(ins, reps, mask):
        seen["shapes"] = [x.shape for x in ins]
        seen["orders"] = {x.flags.f_contiguous and not x.flags.c_contiguous for x in ins + reps + [mask]}
        for a, o in zip(ins, port.poisson_blend(ins, reps, mask, tol=1e-6)[0]):
            a[...] = o
        return True

    monkeypatch.setattr(drivers, "_close_mask", close)
    monkeypatch.setattr(drivers, "_blend_in_place", blend)
    rc = drivers.poisson_main([str(a), str(b)] + (["--reference-layout"] if layout == "reference" else []))
    assert rc == 0
    out = tmp_path / "in" / "poisson_simple_replace" / "scene.tif"  # poisson-main.cpp:69
    t = gt.GeoTIFF(out, np.float64)
    assert t.raster_count == 6 and t.file.dtype == np.uint16 and t.geo_transform == (5e5, 10.0, 0.0, 6e6, 0.0, -10.0)
    got = t.read()
    assert np.array_equal(got[5], cloud)  # band 6 is the template's (CreateCopy)
    # the expected result, computed here on the layout the driver was asked for
    to = (lambda x: x.astype(np.float64)) if layout == "raster" else (
        lambda x: x.astype(np.float64).ravel().reshape((rows, cols), order="F"))  # fmt: skip
    back = (lambda m: m) if layout == "raster" else (lambda m: m.reshape(-1, order="F").reshape(rows, cols))
    mask = oracle.morph_close_mask(to(cloud), 5)
    assert np.array_equal(np.asarray(seen["band"]), to(cloud)) and seen["shapes"] == [(rows, cols)] * 5
    assert seen["orders"] == {layout == "reference"}  # every array reaches the C-ABI in one memory order, uncopied
    want = port.poisson_blend([to(x) for x in bands_in], [to(x) for x in bands_rp], mask, tol=1e-6)[0]
    for k in range(5):
        assert np.array_equal(got[k], gt.gdal_convert(back(want[k]), np.uint16)), k
        keep = ~back(mask)
        assert np.array_equal(got[k][keep], bands_in[k][keep])
    if layout == "reference":  # and the two layouts really are different problems on a non-square scene
        assert not np.array_equal(back(mask), oracle.morph_close_mask(cloud.astype(np.float64), 5))


def test_fill_folder_cli_arguments(tmp_path, monkeypatch):
    from satellite_approximation_b200 import scenes

    seen = {}
    monkeypatch.setattr(scenes, "fill_missing_data_folder", lambda *a, **k: seen.update(laplace=(a, k)) or {})
    monkeypatch.setattr(scenes, "blend_missing_data_folder", lambda *a, **k: seen.update(poisson=(a, k)) or {"d": {"B04": 1}})
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("LOCAL_RANK", "1")
    monkeypatch.delenv("SATFILL_DEVICE", raising=False)
    assert drivers.main(["fill_folder", str(tmp_path), "--bands", "B04,B08", "--no-cache", "--skip-threshold", "0.7"]) == 0
    assert seen["laplace"] == ((str(tmp_path), ["B04", "B08"], False, 0.7), {"shard": (1, 4)})
    assert drivers.main(["fill_folder", str(tmp_path), "--bands", "B04", "--poisson", "--distance-weight", "0.25"]) == 0
    assert seen["poisson"] == ((str(tmp_path), ["B04"], True, 0.5, 0.25), {"shard": (1, 4)})
    assert os.environ.pop("SATFILL_DEVICE") == "1"  # one process per GPU: the default context follows LOCAL_RANK
    assert drivers.main(["fill_folder", str(tmp_path)]) == -1  # --bands is required


def test_poisson_main_failed_solve_writes_the_inputs(tmp_path, monkeypatch):
    """poisson.cpp:263-269 + 292-303: a band that does not converge leaves the images alone, and poisson_main still writes
    its output file -- a copy of the input."""
    import oracle

    a, b, bands_in, _, cloud = make_pair(tmp_path, 40, 32)
    monkeypatch.setattr(drivers, "_close_mask", lambda band: oracle.morph_close_mask(np.asarray(band), 5))
    monkeypatch.setattr(drivers, "_blend_in_place", lambda *args: False)
    assert drivers.poisson_main([str(a), str(b)]) == 0
    got = gt.TiffFile(tmp_path / "in" / "poisson_simple_replace" / "scene.tif").read_all()
    assert all(np.array_equal(g, w) for g, w in zip(got, bands_in + [cloud]))


def test_cpp_poisson_main_binary_without_a_gpu(tmp_path):
    """cpp/src/poisson_main.cpp (the reference's executable on utils/geotiff.h + the `approx` shim): argument and file
    checks, GeoTIFF decoding up to the first device call, and the loud failure -- exit code 2, no CPU fallback -- on a
    machine without a GPU."""
    import subprocess

    import torch

    from satellite_approximation_b200 import _capi

    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "poisson_main")
    if not os.path.exists(exe):
        pytest.skip("poisson_main is not built (make -C cpp needs Eigen headers)")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 255 and "Usage" in r.stderr  # return -1 (poisson-main.cpp:28-31)
    r = subprocess.run([exe, str(tmp_path / "a.tif"), str(tmp_path / "b.tif")], capture_output=True, text=True)
    assert r.returncode == 255 and "does not exist" in r.stderr
    a, b, *_ = make_pair(tmp_path, 40, 32)
    (tmp_path / "junk.tif").write_bytes(b"not a tiff")
    r = subprocess.run([exe, str(tmp_path / "junk.tif"), str(b)], capture_output=True, text=True)
    assert r.returncode == 1 and "not a TIFF" in r.stderr
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the device path of the binary is not part of the CPU suite")
    r = subprocess.run([exe, str(a), str(b)], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr, r.stderr
    assert not os.path.exists(tmp_path / "in" / "poisson_simple_replace")


@pytest.mark.parametrize("layout", ["raster", "reference"])
def test_cpp_poisson_main_binary_host_logic_with_oracle_pixels(tmp_path, port, layout):
    """The whole C++ poisson_main (cpp/src/poisson_main.cpp: utils/geotiff.h decode, approx::preprocess_cloud_band,
    approx::blend_images_poisson, GeoTiffWriter) run end to end on the CPU, with tests/fake_satfill.c -- the eight C-ABI
    entry points the shim imports, answered by the ORACLE -- put in front of the real library through LD_LIBRARY_PATH.
    Checks what surrounds the device calls: layouts, strides, band order, the output file; the expected file is the one
    the Python driver's plumbing test demands."""
    import subprocess

    import oracle
    from satellite_approximation_b200 import _capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_capi.LIB_PATH)
    exe = os.path.join(libdir, "poisson_main")
    if not os.path.exists(exe):
        pytest.skip("poisson_main is not built (make -C cpp needs Eigen headers)")
    oracle.port()  # makes sure oracle/_build/liboracle.so exists
    odir = os.path.join(root, "oracle", "_build")
    fake = tmp_path / "fake"
    fake.mkdir()
    cmd = ["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", os.path.join(root, "include"),
           os.path.join(root, "tests", "fake_satfill.c"), "-o", str(fake / "libsatfill.so"), "-L", odir, "-loracle",
           f"-Wl,-rpath,{odir}"]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows, cols = 48, 36
    a, b, bands_in, bands_rp, cloud = make_pair(tmp_path, rows, cols)
    env = dict(os.environ, LD_LIBRARY_PATH=str(fake) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    args = [exe, str(a), str(b)] + (["--reference-layout"] if layout == "reference" else [])
    r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    got = gt.GeoTIFF(tmp_path / "in" / "poisson_simple_replace" / "scene.tif", np.float64)
    assert got.raster_count == 6 and got.file.dtype == np.uint16 and got.geo_transform == (5e5, 10.0, 0.0, 6e6, 0.0, -10.0)
    got = got.read()
    assert np.array_equal(got[5], cloud)
    to = (lambda x: x.astype(np.float64)) if layout == "raster" else (
        lambda x: x.astype(np.float64).ravel().reshape((rows, cols), order="F"))  # fmt: skip
    back = (lambda m: m) if layout == "raster" else (lambda m: m.reshape(-1, order="F").reshape(rows, cols))
    mask = oracle.morph_close_mask(to(cloud), 5)
    want = port.poisson_blend([to(x) for x in bands_in], [to(x) for x in bands_rp], mask, tol=1e-6)[0]
    for k in range(5):
        assert np.array_equal(got[k], gt.gdal_convert(back(want[k]), np.uint16)), k
        assert np.array_equal(got[k][~back(mask)], bands_in[k][~back(mask)])
