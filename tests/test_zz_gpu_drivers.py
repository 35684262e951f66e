"""GPU: the restated executables and the folder drivers end to end -- files in, GPU fill through the C-ABI, files out --
against the oracle (executables/laplace-main.cpp, executables/poisson-main.cpp, lib/approx/source/laplace.cpp:170-244,
lib/approx/source/poisson.cpp:323-349).  Integer file samples are compared after GDAL's rounding, so the bound is one
digital number; known pixels must be bit-identical."""
from __future__ import annotations

import os

import numpy as np
import pytest
from conftest import GOLDEN

from satellite_approximation_b200 import drivers
from satellite_approximation_b200 import geotiff as gt
from satellite_approximation_b200 import scenes as sc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", ["raster", "reference"])
def test_poisson_main_on_the_gpu(tmp_path, port, layout):
    import oracle
    from test_drivers import make_pair

    rows, cols = 48, 36
    a, b, bands_in, bands_rp, cloud = make_pair(tmp_path, rows, cols)
    assert drivers.poisson_main([str(a), str(b)] + (["--reference-layout"] if layout == "reference" else [])) == 0
    got = gt.GeoTIFF(tmp_path / "in" / "poisson_simple_replace" / "scene.tif", np.float64).read()
    assert len(got) == 6 and np.array_equal(got[5], cloud)
    to = (lambda x: x.astype(np.float64)) if layout == "raster" else (
        lambda x: x.astype(np.float64).ravel().reshape((rows, cols), order="F"))  # fmt: skip
    back = (lambda m: m) if layout == "raster" else (lambda m: m.reshape(-1, order="F").reshape(rows, cols))
    mask = oracle.morph_close_mask(to(cloud), 5)
    assert mask.any()
    want = port.poisson_blend([to(x) for x in bands_in], [to(x) for x in bands_rp], mask, tol=1e-6)[0]
    keep = ~back(mask)
    for k in range(5):
        w = gt.gdal_convert(back(want[k]), np.uint16).astype(np.float64)
        assert np.max(np.abs(got[k] - w)) <= 1.0, k
        assert np.array_equal(got[k][keep], bands_in[k][keep].astype(np.float64)), k
        assert not np.array_equal(got[k], bands_in[k].astype(np.float64))  # something was blended


def test_laplace_main_on_the_gpu(tmp_path):
    case = dict(np.load(os.path.join(GOLDEN, "prepost_cases.npz")))
    img, inv, want, mask = (case[k] for k in ("al_image", "al_invalid", "al_out", "al_mask"))
    base, marked, out = tmp_path / "base.png", tmp_path / "marked.png", tmp_path / "out.png"
    assert drivers.imwrite(base, img) and drivers.imwrite(marked, inv)
    assert drivers.laplace_main([str(base), str(marked), str(out)]) == 0
    got = drivers.imread_color(out)
    assert got.shape == img.shape and got.dtype == np.uint8
    assert np.array_equal(got[~mask], img[~mask])
    diff = np.abs(got.astype(int) - drivers.saturate_u8(want).astype(int))
    assert diff.max() <= 1


def test_folder_drivers_on_the_gpu(tmp_path, port):
    from test_scenes import make_scene_tree

    truth = make_scene_tree(tmp_path)
    with sc.DataBase(tmp_path) as db:
        for name, (_, mask) in truth.items():
            db.write_detection_result(name, True, True, 0.0, 0.0, float(mask.mean()))
    done = sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], use_cache=True, skip_threshold=0.5)
    assert set(done) == {"2019-05-22", "2019-05-12"}  # 06-01 is more than half invalid
    for name, ids in done.items():
        bands, mask = truth[name]
        for b, id_ in ids.items():
            got = gt.TiffFile(tmp_path / name / "approximated_data" / f"{b}_{id_}.tif").read_band(1).astype(np.float64)
            want = gt.gdal_convert(port.laplace_fill(bands[b].astype(np.float64), mask, mode=1)[0], np.uint16)
            assert np.max(np.abs(got - want)) <= 1.0
            assert np.array_equal(got[~mask], bands[b][~mask].astype(np.float64))
    assert sc.fill_missing_data_folder(tmp_path, ["B04", "B08"], True, 0.5) == {}  # cached
    # Poisson against the date find_good_close_image picks (weight 0: the cleanest neighbour, 05-12)
    done = sc.blend_missing_data_folder(tmp_path, ["B04"], True, 0.9, distance_weight=0.0)
    assert set(done) == {"2019-05-22", "2019-06-01"}  # 05-12 keeps itself -> Laplace, which is cached already
    with sc.DataBase(tmp_path) as db:
        assert set(db.get_approx_status("2019-06-01", sc.ApproxMethod.Poisson)) == {"B04"}
    for name in done:
        bands, mask = truth[name]
        guide = truth["2019-05-12"][0]["B04"].astype(np.float64)
        want = port.poisson_blend([bands["B04"].astype(np.float64)], [guide], mask, tol=1e-6)[0][0]
        id_ = done[name]["B04"]
        got = gt.TiffFile(tmp_path / name / "approximated_data" / f"B04_{id_}.tif").read_band(1).astype(np.float64)
        assert np.max(np.abs(got - gt.gdal_convert(want, np.uint16))) <= 1.0
        assert np.array_equal(got[~mask], bands["B04"][~mask].astype(np.float64))


def test_cpp_poisson_main_binary_on_the_gpu(tmp_path, port):
    """cpp/src/poisson_main.cpp end to end: GeoTIFFs in (utils/geotiff.h), approx::preprocess_cloud_band +
    approx::blend_images_poisson on the GPU, GeoTIFF out -- against the oracle, as for the Python driver above."""
    import subprocess

    import oracle
    from test_drivers import make_pair

    from satellite_approximation_b200 import _capi

    exe = os.path.join(os.path.dirname(_capi.LIB_PATH), "poisson_main")
    if not os.path.exists(exe):
        pytest.skip("poisson_main is not built (make -C cpp needs Eigen headers)")
    os.chmod(exe, 0o755)
    rows, cols = 48, 36
    a, b, bands_in, bands_rp, cloud = make_pair(tmp_path, rows, cols)
    r = subprocess.run([exe, str(a), str(b)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = gt.GeoTIFF(tmp_path / "in" / "poisson_simple_replace" / "scene.tif", np.float64).read()
    assert len(got) == 6 and np.array_equal(got[5], cloud)
    mask = oracle.morph_close_mask(cloud.astype(np.float64), 5)
    want = port.poisson_blend([x.astype(np.float64) for x in bands_in], [x.astype(np.float64) for x in bands_rp], mask,
                              tol=1e-6)[0]  # fmt: skip
    for k in range(5):
        w = gt.gdal_convert(want[k], np.uint16).astype(np.float64)
        assert np.max(np.abs(got[k] - w)) <= 1.0, k
        assert np.array_equal(got[k][~mask], bands_in[k][~mask].astype(np.float64)), k
