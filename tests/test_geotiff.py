"""CPU: the GeoTIFF step either side of the Poisson path (satellite_approximation_b200/geotiff.py; reference
lib/utils/include/utils/geotiff.h:98-263, executables/poisson-main.cpp:53-70).  Pinned against crops of the reference's
sample scene re-encoded in the sample files' own flavour and decoded by OpenCV/libtiff (oracle/make_tiff_fixture.py),
against Pillow/libtiff-written files where Pillow is importable, and by round trips."""
from __future__ import annotations

import hashlib
import os
import struct
import zlib

import numpy as np
import pytest

from satellite_approximation_b200 import geotiff as gt

from conftest import GOLDEN

SCENE = "/root/reference/test_data/2019-05-22"


@pytest.fixture(scope="module")
def expected():
    return dict(np.load(os.path.join(GOLDEN, "scene_crop_expected.npz")))


@pytest.mark.parametrize("name,dtype", [("B04", np.uint16), ("CLD", np.uint8), ("sunZenithAngles", np.float32)])
def test_sample_scene_flavour_decodes_bit_exact(expected, name, dtype):
    t = gt.TiffFile(os.path.join(GOLDEN, f"scene_crop_{name}.tif"))
    assert (t.byteorder, t.compression, t.seg_h, t.tiled, t.samples_per_pixel) == (">", 32946, 8, False, 1)
    a = t.read_band(1)
    assert a.dtype == dtype and a.flags.c_contiguous
    assert np.array_equal(a, expected[name])
    assert np.allclose(t.geo_transform, expected["geo_transform"], rtol=0, atol=0)


def test_full_sample_scene_digests(expected):
    if not os.path.isdir(SCENE):
        pytest.skip("reference sample scene not on this machine")
    for line in expected["full_digests"]:
        name, dtype, shape, digest = str(line).split()
        a = gt.TiffFile(os.path.join(SCENE, name + ".tif")).read_band(1)
        assert str(a.dtype) == dtype and f"{a.shape[0]}x{a.shape[1]}" == shape
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == digest


def test_geotiff_mirror_layouts_and_conversion(expected):
    p = os.path.join(GOLDEN, "scene_crop_B04.tif")
    g = gt.GeoTIFF(p, np.float64)
    assert (g.height, g.width, g.raster_count) == (96, 80, 1)
    a = g.read(1)
    assert a.dtype == np.float64 and np.array_equal(a, expected["B04"].astype(np.float64))
    assert [x.shape for x in g.read([1, 1])] == [(96, 80)] * 2 and len(g.read()) == 1
    # the reference's matrix: column-major height x width over the row-major raster buffer (geotiff.h:234-253)
    m = gt.GeoTIFF(p, np.float64, layout="reference").read(1)
    assert m.shape == (96, 80) and m.flags.f_contiguous
    flat = expected["B04"].astype(np.float64).ravel()
    r, c = np.meshgrid(np.arange(96), np.arange(80), indexing="ij")
    assert np.array_equal(m, flat[r + c * 96])
    assert not np.array_equal(m, a)  # a non-square scene is index-scrambled there
    x, y = g.pixel_to_geo(0, 0)
    assert (x, y) == (expected["geo_transform"][0], expected["geo_transform"][3])
    with pytest.raises(gt.TiffError):
        g.read(2)


def test_gdal_convert_rounds_and_clamps():
    v = np.array([-3.7, -0.5, -0.49, 0.49, 0.5, 1.5, 2.5, 65534.5, 65535.4, 1e9, np.nan, np.inf, -np.inf])
    assert gt.gdal_convert(v, np.uint16).tolist() == [0, 0, 0, 0, 1, 2, 3, 65535, 65535, 65535, 0, 65535, 0]
    assert gt.gdal_convert(v[:7], np.int16).tolist() == [-4, -1, 0, 0, 1, 2, 3]
    assert gt.gdal_convert(np.array([-5, 300], np.int32), np.uint8).tolist() == [0, 255]
    assert gt.gdal_convert(np.array([70000], np.uint32), np.int16).tolist() == [32767]
    assert gt.gdal_convert(np.array([1e30, -1e30]), np.int64).tolist() == [2**63 - 1, -(2**63)]
    a = np.arange(5, dtype=np.uint16)
    assert gt.gdal_convert(a, np.uint16) is a and gt.gdal_convert(a, np.float64).dtype == np.float64


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64,
                                   np.uint64, np.int64])  # fmt: skip
@pytest.mark.parametrize("kw", [{}, {"tile": (16, 32)}, {"compress": True}, {"bigtiff": True}, {"rows_per_strip": 5},
                                {"tile": (32, 16), "compress": True, "bigtiff": True}])  # fmt: skip
def test_write_read_round_trip(tmp_path, dtype, kw):
    rng = np.random.default_rng(7)
    a = (rng.random((37, 53)) * 200 - (50 if np.dtype(dtype).kind != "u" else 0)).astype(dtype)
    b = a[::-1].copy()
    p = tmp_path / "x.tif"
    gt.write_tiff(p, [a, b], **kw)
    t = gt.TiffFile(p)
    assert t.bigtiff == bool(kw.get("bigtiff")) and t.tiled == ("tile" in kw) and t.planar == 2
    got = t.read_all()
    assert got[0].dtype == np.dtype(dtype) and np.array_equal(got[0], a) and np.array_equal(got[1], b)
    assert np.array_equal(t.read_band(2), b)


def test_pillow_written_files(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    a = (rng.random((61, 47)) * 60000).astype(np.uint16)
    rgb = (rng.random((61, 47, 3)) * 255).astype(np.uint8)
    p = tmp_path / "p.tif"
    for comp, code in [("raw", 1), ("tiff_lzw", 5), ("tiff_adobe_deflate", 8), ("packbits", 32773)]:
        Image.fromarray(a).save(p, compression=comp)
        t = gt.TiffFile(p)
        assert t.compression == code and np.array_equal(t.read_band(1), a), comp
        Image.fromarray(rgb).save(p, compression=comp)
        t = gt.TiffFile(p)  # chunky RGB
        assert t.planar == 1 and t.samples_per_pixel == 3
        assert all(np.array_equal(t.read_band(i + 1), rgb[:, :, i]) for i in range(3)), comp
        assert all(np.array_equal(x, rgb[:, :, i]) for i, x in enumerate(t.read_all())), comp
    # and the other direction: libtiff reads what write_tiff makes (strips, tiles, deflate)
    for kw in [{}, {"tile": (16, 16)}, {"compress": True}]:
        gt.write_tiff(p, [a], **kw)
        assert np.array_equal(np.array(Image.open(p)), a), kw


def _encode(path, a, compression=1, predictor=1, byteorder="<", tile=None, geo=False):
    """A deliberately separate minimal encoder for the flavours no library here writes on demand (predictors 2 / 3,
    big-endian, tiles): chunky samples, one strip or fixed tiles."""
    bo = byteorder
    h, w = a.shape[:2]
    ns = 1 if a.ndim == 2 else a.shape[2]
    a3 = a.reshape(h, w, ns)

    def enc(block):
        rows = block.shape[0]
        if predictor == 2:
            d = block.copy()
            d[:, 1:] = block[:, 1:] - block[:, :-1]
            raw = d.astype(d.dtype.newbyteorder(bo)).tobytes()
        elif predictor == 3:
            be = block.astype(block.dtype.newbyteorder(">")).view(np.uint8).reshape(rows, block.shape[1] * ns, -1)
            planes = np.ascontiguousarray(be.transpose(0, 2, 1)).reshape(rows, -1)  # most significant bytes first
            d = planes.copy()
            d[:, ns:] = planes[:, ns:] - planes[:, :-ns]
            raw = d.tobytes()
        else:
            raw = block.astype(block.dtype.newbyteorder(bo)).tobytes()
        return zlib.compress(raw) if compression == 8 else raw

    if tile:
        th, tw = tile
        segs = []
        for r0 in range(0, h, th):
            for c0 in range(0, w, tw):
                t = np.zeros((th, tw, ns), a.dtype)
                blk = a3[r0 : r0 + th, c0 : c0 + tw]
                t[: blk.shape[0], : blk.shape[1]] = blk
                segs.append(enc(t))
    else:
        segs = [enc(a3)]
    blob = bytearray((b"II" if bo == "<" else b"MM") + struct.pack(bo + "HI", 42, 0))
    offs = []
    for s in segs:
        offs.append(len(blob))
        blob += s + (b"\0" if len(s) & 1 else b"")
    n = len(segs)
    o_offs, o_cnts = len(blob), len(blob) + 4 * n
    blob += struct.pack(f"{bo}{n}I", *offs) + struct.pack(f"{bo}{n}I", *[len(s) for s in segs])
    o_bits = len(blob)
    blob += struct.pack(f"{bo}{ns}H", *[a.dtype.itemsize * 8] * ns) + struct.pack(f"{bo}{ns}H", *[{"u": 1, "i": 2, "f": 3}[a.dtype.kind]] * ns)
    sh = lambda v: struct.pack(bo + "HH", v, 0)  # noqa: E731
    lg = lambda v: struct.pack(bo + "I", v)  # noqa: E731
    multi = ns > 2
    ent = [(256, 4, 1, lg(w)), (257, 4, 1, lg(h)),
           (258, 3, ns, lg(o_bits) if multi else struct.pack(f"{bo}{ns}H", *[a.dtype.itemsize * 8] * ns).ljust(4, b"\0")),
           (259, 3, 1, sh(compression)), (262, 3, 1, sh(1)), (277, 3, 1, sh(ns)), (284, 3, 1, sh(1)), (317, 3, 1, sh(predictor)),
           (339, 3, ns, lg(o_bits + 2 * ns) if multi else
            struct.pack(f"{bo}{ns}H", *[{"u": 1, "i": 2, "f": 3}[a.dtype.kind]] * ns).ljust(4, b"\0"))]  # fmt: skip
    one = lambda o, lst: lg(o) if n > 1 else lg(lst[0])  # noqa: E731
    if tile:
        ent += [(322, 4, 1, lg(tile[1])), (323, 4, 1, lg(tile[0])), (324, 4, n, one(o_offs, offs)),
                (325, 4, n, one(o_cnts, [len(s) for s in segs]))]  # fmt: skip
    else:
        ent += [(273, 4, n, one(o_offs, offs)), (278, 4, 1, lg(h)), (279, 4, n, one(o_cnts, [len(s) for s in segs]))]
    if geo:  # ModelPixelScale + ModelTiepoint (type 12 = DOUBLE), values behind the segment tables
        o_scale = len(blob)
        blob += struct.pack(bo + "3d", 10.0, 10.0, 0.0)
        o_tie = len(blob)
        blob += struct.pack(bo + "6d", 0.0, 0.0, 0.0, 5e5, 6e6, 0.0)
        ent += [(33550, 12, 3, lg(o_scale)), (33922, 12, 6, lg(o_tie))]
    ent.sort()
    ifd = len(blob)
    blob += struct.pack(bo + "H", len(ent))
    for tag, typ, cnt, val in ent:
        blob += struct.pack(bo + "HHI", tag, typ, cnt) + val
    blob += struct.pack(bo + "I", 0)
    blob[4:8] = struct.pack(bo + "I", ifd)
    with open(path, "wb") as f:
        f.write(bytes(blob))


@pytest.mark.parametrize("bo", ["<", ">"])
def test_predictors_tiles_and_byte_orders(tmp_path, bo):
    rng = np.random.default_rng(11)
    p = tmp_path / "e.tif"
    u = (rng.random((40, 50, 3)) * 65535).astype(np.uint16)
    for tile in (None, (16, 32)):
        for comp in (1, 8):
            _encode(p, u, compression=comp, predictor=2, byteorder=bo, tile=tile)
            t = gt.TiffFile(p)
            assert t.predictor == 2 and t.samples_per_pixel == 3
            assert all(np.array_equal(x, u[:, :, i]) for i, x in enumerate(t.read_all()))
            assert np.array_equal(t.read_band(2), u[:, :, 1])
    for dt in (np.float32, np.float64):
        f = rng.standard_normal((33, 29)).astype(dt)
        f2 = rng.standard_normal((33, 29, 2)).astype(dt)
        for tile in (None, (16, 16)):
            _encode(p, f, compression=8, predictor=3, byteorder=bo, tile=tile)
            assert np.array_equal(gt.TiffFile(p).read_band(1), f)
            _encode(p, f2, compression=1, predictor=3, byteorder=bo, tile=tile)
            assert np.array_equal(gt.TiffFile(p).read_band(2), f2[:, :, 1])


def test_writer_copies_template_and_overwrites_bands(tmp_path, expected):
    # a 3-band u16 template with the sample scene's geo tags
    src = gt.TiffFile(os.path.join(GOLDEN, "scene_crop_B04.tif"))
    geo = {k: v for k, v in src.tags.items() if k in gt.GEO_TAGS}
    b = expected["B04"]
    tpl = tmp_path / "tpl.tif"
    gt.write_tiff(tpl, [b, (b // 2).astype(np.uint16), (b // 3).astype(np.uint16)], extra_tags=geo)
    vals = [b.astype(np.float64) + 0.5, b.astype(np.float64) * 100.0]  # the second saturates u16
    out = tmp_path / "sub" / "out.tif"
    gt.GeoTiffWriter(vals, tpl).write(out, start_index=2)  # poisson-main.cpp:68-69 writes from band 1; 2 tests the offset
    t = gt.GeoTIFF(out, np.float64)
    assert t.raster_count == 3 and t.file.dtype == np.uint16 and t.file.compression == 1
    assert t.geo_transform == src.geo_transform and t.file.tags[gt.T_GEOASCII][1] == "WGS 84|"
    got = t.read()
    assert np.array_equal(got[0], b)  # untouched band keeps the template's pixels (CreateCopy)
    assert np.array_equal(got[1], np.minimum(b.astype(np.float64) + 1.0, 65535.0))  # x.5 rounds away from zero
    assert np.array_equal(got[2], np.minimum(b.astype(np.float64) * 100.0, 65535.0))
    # single-band form always lands in band 1; reference layout round-trips through read -> write
    m = gt.GeoTIFF(tpl, np.float64, layout="reference").read(2)
    gt.GeoTiffWriter(m, tpl, layout="reference").write(out, start_index=3)
    assert np.array_equal(gt.TiffFile(out).read_band(1), b // 2)
    with pytest.raises(RuntimeError):
        gt.GeoTiffWriter(vals, tpl).write(out, start_index=3)  # band 4 of 3
    with pytest.raises(RuntimeError):
        gt.GeoTiffWriter(vals[0][:10], tpl).write(out)


def test_errors(tmp_path):
    p = tmp_path / "bad.tif"
    p.write_bytes(b"not a tiff at all")
    with pytest.raises(gt.TiffError):
        gt.TiffFile(p)
    with pytest.raises(gt.TiffError):
        gt.TiffFile(tmp_path / "missing.tif")
    a = np.zeros((4, 4), np.uint8)
    gt.write_tiff(p, [a])
    with pytest.raises(gt.TiffError):  # no geo tags: the reference throws IOError (geotiff.h:220-222)
        gt.GeoTIFF(p)
    raw = bytearray(p.read_bytes())
    with pytest.raises(ValueError):
        gt.write_tiff(p, [a, np.zeros((4, 5), np.uint8)])
    with pytest.raises(ValueError):
        gt.write_tiff(p, [a], tile=(8, 8))
    p.write_bytes(bytes(raw[:-20]))  # truncated directory
    with pytest.raises(gt.TiffError):
        gt.TiffFile(p)


def test_native_and_interpreter_decoders_agree(tmp_path):
    """csrc/tiffcodec.c (lib/libsattiff.so) against the pure-Python decoders on libtiff-written LZW / PackBits streams,
    including streams long enough to reset the LZW table, and on corrupt input."""
    Image = pytest.importorskip("PIL.Image")
    lib = gt._native_codec()
    assert lib, "libsattiff.so is not built (make -C satellite_approximation_b200/csrc)"
    rng = np.random.default_rng(5)
    noisy = (rng.random((300, 400)) * 65535).astype(np.uint16)  # incompressible: many table resets
    smooth = (np.add.outer(np.arange(300), np.arange(400)) // 7).astype(np.uint16)  # long matches
    flat = np.zeros((300, 400), np.uint8)
    p = tmp_path / "n.tif"
    for a in (noisy, smooth, flat):
        for comp, fn, py in (("tiff_lzw", lib.st_lzw_decode, gt._lzw_decode), ("packbits", lib.st_packbits_decode, gt._packbits_decode)):
            Image.fromarray(a).save(p, compression=comp)
            t = gt.TiffFile(p)
            assert np.array_equal(t.read_band(1), a)  # through the native decoder
            for off, cnt in zip(t._offsets, t._counts):
                raw = t._buf[int(off) : int(off) + int(cnt)]
                want = py(raw)
                cap = t.seg_h * t.seg_w * a.dtype.itemsize
                assert gt._decode_native(fn, raw, cap) == want[:cap]
                assert gt._decode_native(fn, raw, 10) == want[:10]  # clipped output
    with pytest.raises(gt.TiffError):
        gt._decode_native(lib.st_lzw_decode, b"\x00\x00\x00\x00", 16)  # no clear code first
    with pytest.raises(gt.TiffError):
        gt._lzw_decode(b"\x00\x00\x00\x00")
    bad = bytes([0x80, 0x7F, 0xFF, 0xC0])  # clear, then code 0x1FF (>= next) as the first code
    with pytest.raises(gt.TiffError):
        gt._decode_native(lib.st_lzw_decode, bad, 16)
    with pytest.raises(gt.TiffError):
        gt._lzw_decode(bad)


def test_threaded_segment_decode_matches_serial(tmp_path, monkeypatch):
    rng = np.random.default_rng(9)
    a = (np.add.outer(np.arange(1300), np.arange(1700)) % 4096 + rng.integers(0, 40, (1300, 1700))).astype(np.uint16)
    b = a[::-1].copy()
    for kw in ({"rows_per_strip": 8}, {"tile": (256, 256)}):
        p = tmp_path / "big.tif"
        gt.write_tiff(p, [a, b], compress=True, **kw)
        monkeypatch.setattr(gt.TiffFile, "decode_threads", 4)
        par = gt.TiffFile(p).read_all()
        monkeypatch.setattr(gt.TiffFile, "decode_threads", 1)
        ser = gt.TiffFile(p).read_all()
        assert np.array_equal(par[0], a) and np.array_equal(par[1], b)
        assert all(np.array_equal(x, y) for x, y in zip(par, ser))
    # an error inside a worker surfaces as the reader's error
    raw = bytearray((tmp_path / "big.tif").read_bytes())
    t = gt.TiffFile(tmp_path / "big.tif")
    off = int(t._offsets[len(t._offsets) // 2])
    raw[off : off + 8] = b"\xff" * 8  # corrupt one deflate stream
    (tmp_path / "bad.tif").write_bytes(bytes(raw))
    monkeypatch.setattr(gt.TiffFile, "decode_threads", 4)
    with pytest.raises(gt.TiffError):
        gt.TiffFile(tmp_path / "bad.tif").read_all()


def test_geo_referencing_helpers(tmp_path, expected):
    """geotiff.h:322-421 on the sample-scene crop (north-up, lat/long degrees)."""
    g = gt.GeoTIFF(os.path.join(GOLDEN, "scene_crop_B04.tif"), np.float64)
    x0, dx, _, y0, _, dy = expected["geo_transform"]
    assert (g.west(), g.north(), g.east_west_step(), g.north_south_step()) == (x0, y0, dx, dy)
    assert g.east() == x0 + 80 * dx and g.south() == y0 + 96 * dy
    assert g.north_west() == (y0, x0) and g.south_east() == (g.south(), g.east())
    assert g.north_east() == (y0, g.east()) and g.south_west() == (g.south(), x0)
    v = g.read(1)
    pos = (y0 + 10.5 * dy, x0 + 20.5 * dx)  # the middle of pixel row 10, column 20
    assert g.index_at(pos) == (20, 10) and g.value_at(pos, v) == v[10, 20]
    assert g.uv_at(pos) == (20 / 80, 10 / 96)
    assert np.allclose(g.mid_point_of_pixel((10, 20)), pos, rtol=0, atol=1e-12)
    assert g.index_at((y0 + 1.0, x0 - 1.0)) == (0, 0) and g.index_at((y0 + 1e3 * dy, x0 + 1e3 * dx)) == (79, 95)  # clamped
    # bilinear: between four pixel corners = the weighted mean of the four samples
    q = (y0 + 10.25 * dy, x0 + 20.75 * dx)
    want = (0.25 * 0.75) * v[10, 20] + (0.75 * 0.75) * v[10, 21] + (0.25 * 0.25) * v[11, 20] + (0.75 * 0.25) * v[11, 21]
    assert np.isclose(g.bilinear_value_at(q, v), want, rtol=1e-9)
    assert gt.GeoTIFF.value_domain(v) == (v.min(), v.max())
    dem = np.array([[-32767.0, 5.0], [12.0, -32767.0]], np.float32)
    assert gt.GeoTIFF.dem_value_domain(dem) == (5.0, 12.0) and gt.GeoTIFF.value_domain(dem, np.int32) == (-32767, 12)
    # write(matrix, path, band): a copy with that band replaced
    g.write(v[::-1] + 1.0, tmp_path / "w.tif", 1)
    assert np.array_equal(gt.TiffFile(tmp_path / "w.tif").read_band(1), expected["B04"][::-1] + 1)


def test_round_trip_property(tmp_path):
    """Any shape / type / segmenting / band count survives write -> read (hypothesis; ragged last strips and tiles)."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    dtypes = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.float32, np.float64]

    @hyp.settings(max_examples=60, deadline=None, suppress_health_check=list(hyp.HealthCheck))
    @hyp.given(h=st.integers(1, 70), w=st.integers(1, 70), nb=st.integers(1, 4), dt=st.sampled_from(dtypes),
               seg=st.one_of(st.none(), st.integers(1, 80), st.tuples(st.sampled_from([16, 32, 48]), st.sampled_from([16, 32]))),
               compress=st.booleans(), big=st.booleans(), seed=st.integers(0, 2**31 - 1))  # fmt: skip
    def run(h, w, nb, dt, seg, compress, big, seed):
        rng = np.random.default_rng(seed)
        bands = [(rng.random((h, w)) * 250 - (100 if np.dtype(dt).kind != "u" else 0)).astype(dt) for _ in range(nb)]
        kw = {"tile": seg} if isinstance(seg, tuple) else {"rows_per_strip": seg}
        p = tmp_path / "h.tif"
        gt.write_tiff(p, bands, compress=compress, bigtiff=big, **kw)
        t = gt.TiffFile(p)
        assert (t.height, t.width, t.samples_per_pixel, t.bigtiff) == (h, w, nb, big)
        for k, b in enumerate(bands):
            assert np.array_equal(t.read_band(k + 1), b)
        assert all(np.array_equal(x, y) for x, y in zip(t.read_all(), bands))
        # both layouts of the GeoTIFF mirror invert through the writer
        geo = {gt.T_PIXEL_SCALE: (12, [1.0, 1.0, 0.0]), gt.T_TIEPOINT: (12, [0.0] * 6)}
        gt.write_tiff(p, bands, extra_tags=geo, compress=compress, **kw)
        for layout in ("raster", "reference"):
            g = gt.GeoTIFF(p, np.float64, layout=layout)
            vals = g.read()
            gt.GeoTiffWriter(vals, p, layout=layout).write(tmp_path / "h2.tif")
            assert all(np.array_equal(x, y) for x, y in zip(gt.TiffFile(tmp_path / "h2.tif").read_all(), bands))

    run()


def test_writer_may_overwrite_its_own_template(tmp_path, expected):
    """Destination == template: the template stays mapped while the new file is written beside it and renamed into place."""
    geo = {k: v for k, v in gt.TiffFile(os.path.join(GOLDEN, "scene_crop_B04.tif")).tags.items() if k in gt.GEO_TAGS}
    b = expected["B04"]
    p = tmp_path / "scene.tif"
    gt.write_tiff(p, [b, b // 2], extra_tags=geo, compress=True)
    gt.GeoTiffWriter([b.astype(np.float64) + 3], p).write(p, start_index=1)
    got = gt.TiffFile(p).read_all()
    assert np.array_equal(got[0], b + 3) and np.array_equal(got[1], b // 2)
    assert sorted(os.listdir(tmp_path)) == ["scene.tif"]  # no temporary left behind
    with pytest.raises(ValueError):
        gt.write_tiff(p, [b, b[:5]])
    assert sorted(os.listdir(tmp_path)) == ["scene.tif"]


def test_tiff_file_is_a_context_manager(expected):
    with gt.TiffFile(os.path.join(GOLDEN, "scene_crop_CLD.tif")) as t:
        assert np.array_equal(t.read_band(1), expected["CLD"])
    with pytest.raises(gt.TiffError):
        t.read_band(1)  # closed: the segments are gone


# ---- the C++ twin (cpp/include/utils/geotiff.h), through satellite_approximation._core ---------------------------------
def _core_or_skip():
    try:
        from satellite_approximation import _core
    except ImportError:
        pytest.skip("satellite_approximation._core is not built (make -C cpp pybind needs Eigen headers)")
    if not hasattr(_core, "geotiff_read"):
        pytest.skip("stale _core without the GeoTIFF bindings")
    return _core


def test_cpp_geotiff_reads_what_the_python_reader_reads(tmp_path, expected):
    core = _core_or_skip()
    for name in ("B04", "CLD", "sunZenithAngles"):  # the sample scene's flavour: big-endian, deflate 32946, 8-row strips
        p = os.path.join(GOLDEN, f"scene_crop_{name}.tif")
        got = core.geotiff_read(p, 1)
        assert got.dtype == np.float64 and got.flags.f_contiguous and np.array_equal(got, expected[name].astype(np.float64))
        ref = core.geotiff_read(p, 1, True)  # the reference's matrix: column-major over the row-major raster
        assert np.array_equal(ref, gt.GeoTIFF(p, np.float64, layout="reference").read(1))
    h, w, n, g = core.geotiff_info(os.path.join(GOLDEN, "scene_crop_B04.tif"))
    assert (h, w, n) == (96, 80, 1) and np.array_equal(g, expected["geo_transform"])
    assert np.array_equal(core.geotiff_read_u8(os.path.join(GOLDEN, "scene_crop_B04.tif"), 1),
                          np.minimum(expected["B04"], 255).astype(np.uint8))  # GDAL clamps on a narrowing read
    # every flavour the Python writer and the test encoder produce
    rng = np.random.default_rng(21)
    a = (rng.random((45, 70)) * 60000).astype(np.uint16)
    b = (rng.random((45, 70)) * 60000).astype(np.uint16)
    geo = {gt.T_PIXEL_SCALE: (12, [10.0, 10.0, 0.0]), gt.T_TIEPOINT: (12, [0.0, 0.0, 0.0, 5e5, 6e6, 0.0])}
    p = tmp_path / "x.tif"
    for kw in ({}, {"tile": (16, 32)}, {"compress": True}, {"bigtiff": True, "tile": (32, 16), "compress": True},
               {"rows_per_strip": 7}):  # fmt: skip
        gt.write_tiff(p, [a, b], extra_tags=geo, **kw)
        assert np.array_equal(core.geotiff_read(str(p), 1), a) and np.array_equal(core.geotiff_read(str(p), 2), b), kw
    rgb = (rng.random((40, 50, 3)) * 65535).astype(np.uint16)
    for bo in "<>":
        for tile in (None, (16, 32)):
            for comp in (1, 8):
                _encode(p, rgb, compression=comp, predictor=2, byteorder=bo, tile=tile, geo=True)  # chunky + predictor
                assert gt.TiffFile(p).geo_transform == (5e5, 10.0, 0.0, 6e6, 0.0, -10.0)
                for k in range(3):
                    assert np.array_equal(core.geotiff_read(str(p), k + 1), rgb[:, :, k]), (bo, tile, comp, k)
            _encode(p, rgb, compression=8, predictor=2, byteorder=bo, tile=tile)
            with pytest.raises(OSError):
                core.geotiff_read(str(p), 2)  # no geo tags: IOError, like the reference's constructor (geotiff.h:220-222)
    f = rng.standard_normal((33, 29))
    gt.write_tiff(p, [f, -f], extra_tags=geo, compress=True)
    assert np.array_equal(core.geotiff_read(str(p), 2), -f)
    assert np.array_equal(core.geotiff_read_i16(str(p), 1), gt.gdal_convert(f, np.int16))
    with pytest.raises(RuntimeError):
        core.geotiff_read(str(p), 3)
    with pytest.raises(OSError):
        core.geotiff_read(str(tmp_path / "missing.tif"), 1)
    (tmp_path / "junk.tif").write_bytes(b"II*\0\xff\xff\xff\x7f")
    with pytest.raises(OSError):
        core.geotiff_read(str(tmp_path / "junk.tif"), 1)
    Image = pytest.importorskip("PIL.Image")
    info = {33550: (10.0, 10.0, 0.0), 33922: (0.0, 0.0, 0.0, 5e5, 6e6, 0.0)}
    smooth = (np.add.outer(np.arange(300), np.arange(400)) // 7).astype(np.uint16)  # long LZW matches, table resets
    for img in (a, smooth):
        for comp in ("tiff_lzw", "packbits", "tiff_adobe_deflate", "raw"):  # libtiff-written: csrc/tiffcodec.c and zlib
            Image.fromarray(img).save(p, compression=comp, tiffinfo=info)
            assert np.array_equal(core.geotiff_read(str(p), 1), img), comp


def test_cpp_geotiff_writer_matches_the_python_writer(tmp_path, expected):
    core = _core_or_skip()
    src = gt.TiffFile(os.path.join(GOLDEN, "scene_crop_B04.tif"))
    geo = {k: v for k, v in src.tags.items() if k in gt.GEO_TAGS}
    b = expected["B04"]
    tpl = tmp_path / "tpl.tif"
    gt.write_tiff(tpl, [b, (b // 2).astype(np.uint16), (b // 3).astype(np.uint16)], extra_tags=geo, compress=True)
    vals = [b.astype(np.float64) + 0.5, b.astype(np.float64) * 100.0 - 7e4]  # rounds half up; saturates both ways
    for ref_layout in (False, True):
        layout = "reference" if ref_layout else "raster"
        py_out, cc_out = tmp_path / f"py_{layout}.tif", tmp_path / "sub" / f"cc_{layout}.tif"
        send = [gt._to_layout(np.ascontiguousarray(v), layout) for v in vals]
        gt.GeoTiffWriter(send, tpl, layout=layout).write(py_out, start_index=2)
        core.geotiff_write([np.asfortranarray(v) for v in send], str(tpl), str(cc_out), 2, ref_layout)
        a, c = gt.GeoTIFF(py_out, np.float64), gt.GeoTIFF(cc_out, np.float64)
        assert c.raster_count == 3 and c.file.dtype == np.uint16 and c.geo_transform == a.geo_transform == src.geo_transform
        assert c.file.tags[gt.T_GEOASCII][1] == "WGS 84|"
        for x, y in zip(a.read(), c.read()):
            assert np.array_equal(x, y)
        assert np.array_equal(c.read(1), b)  # band 1 is the template's
    # single-band form lands in band 1 whatever start_index says (geotiff.h:160-163)
    core.geotiff_write([np.asfortranarray(vals[0])], str(tpl), str(tmp_path / "one.tif"), 3, False, True)
    got = gt.TiffFile(tmp_path / "one.tif").read_all()
    assert np.array_equal(got[0], gt.gdal_convert(vals[0], np.uint16)) and np.array_equal(got[2], b // 3)
    with pytest.raises(RuntimeError):
        core.geotiff_write([np.asfortranarray(v) for v in vals], str(tpl), str(tmp_path / "bad.tif"), 3)  # band 4 of 3
    with pytest.raises(RuntimeError):
        core.geotiff_write([np.asfortranarray(vals[0][:10])], str(tpl), str(tmp_path / "bad.tif"), 1)
    assert not os.path.exists(tmp_path / "bad.tif")
