"""Generator of tests/golden/scene_crop_*.tif + scene_crop_expected.npz (TEST INFRASTRUCTURE, run once in the container
that has /root/reference and cv2; the outputs are committed).

The reference's sample scene test_data/2019-05-22 stores its bands as big-endian ("MM") classic TIFFs, compression 32946
(old-style deflate), 8 rows per strip, with ModelPixelScale / ModelTiepoint / GeoKeyDirectory / GeoAsciiParams tags.  This
script decodes B04.tif (u16), CLD.tif (u8) and sunZenithAngles.tif (f32) with OpenCV (libtiff -- independent of the
repo's codec), crops 96 x 80 pixels, and re-encodes the crops in exactly that flavour with a few lines of struct + zlib
that share nothing with satellite_approximation_b200/geotiff.py.  The expected pixel values (from OpenCV) and the GDAL
geo transform of the crop go to the .npz.  It also records SHA-256 digests of the full decoded bands, which
tests/test_geotiff.py checks when /root/reference is present."""
from __future__ import annotations

import hashlib
import os
import struct
import zlib

import numpy as np

SCENE = "/root/reference/test_data/2019-05-22"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
R0, C0, H, W = 700, 500, 96, 80
SCALE = (0.00019446717011477552, 0.0001087675532796607, 0.0)
TIE = (0.0, 0.0, 0.0, -111.93141764318219, 57.105787570770836, 0.0)
# GeographicTypeGeoKey = 4326 (WGS 84), as in the sample files
GEOKEYS = (1, 1, 0, 3, 1024, 0, 1, 2, 1025, 0, 1, 1, 2048, 0, 1, 4326)
GEOASCII = "WGS 84|"


def encode_mm_deflate(a: np.ndarray, tie, path: str) -> None:
    be = a.astype(a.dtype.newbyteorder(">"))
    fmt = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    strips = [zlib.compress(be[r : r + 8].tobytes()) for r in range(0, a.shape[0], 8)]
    blob = bytearray(b"MM" + struct.pack(">HI", 42, 0))
    offs = []
    for s in strips:
        offs.append(len(blob))
        blob += s + (b"\0" if len(s) & 1 else b"")

    def put(data: bytes) -> int:
        o = len(blob)
        blob.extend(data + (b"\0" if len(data) & 1 else b""))
        return o

    n = len(strips)
    o_offs = put(struct.pack(f">{n}I", *offs))
    o_cnts = put(struct.pack(f">{n}I", *[len(s) for s in strips]))
    o_scale = put(struct.pack(">3d", *SCALE))
    o_tie = put(struct.pack(">6d", *tie))
    o_keys = put(struct.pack(f">{len(GEOKEYS)}H", *GEOKEYS))
    o_ascii = put(GEOASCII.encode() + b"\0")
    short = lambda v: struct.pack(">HH", v, 0)  # noqa: E731
    long_ = lambda v: struct.pack(">I", v)  # noqa: E731
    entries = [
        (256, 3, 1, short(a.shape[1])), (257, 3, 1, short(a.shape[0])), (258, 3, 1, short(a.dtype.itemsize * 8)),
        (259, 3, 1, short(32946)), (262, 3, 1, short(1)), (273, 4, n, long_(o_offs)), (277, 3, 1, short(1)),
        (278, 3, 1, short(8)), (279, 4, n, long_(o_cnts)), (339, 3, 1, short(fmt)), (33550, 12, 3, long_(o_scale)),
        (33922, 12, 6, long_(o_tie)), (34735, 3, len(GEOKEYS), long_(o_keys)), (34737, 2, len(GEOASCII) + 1, long_(o_ascii)),
    ]  # fmt: skip
    ifd = len(blob)
    blob += struct.pack(">H", len(entries))
    for tag, typ, cnt, val in entries:
        blob += struct.pack(">HHI", tag, typ, cnt) + val
    blob += struct.pack(">I", 0)
    blob[4:8] = struct.pack(">I", ifd)
    with open(path, "wb") as f:
        f.write(bytes(blob))


def main() -> None:
    import cv2

    expected = {}
    digests = []
    # the crop's tie point: raster (0,0,0) -> the geographic position of pixel (R0, C0) of the full scene
    tie = (0.0, 0.0, 0.0, TIE[3] + C0 * SCALE[0], TIE[4] - R0 * SCALE[1], 0.0)
    for name in ("B04", "CLD", "sunZenithAngles"):
        full = cv2.imread(os.path.join(SCENE, name + ".tif"), cv2.IMREAD_UNCHANGED)
        digests.append(f"{name} {full.dtype} {full.shape[0]}x{full.shape[1]} "
                       + hashlib.sha256(np.ascontiguousarray(full).tobytes()).hexdigest())  # fmt: skip
        crop = np.ascontiguousarray(full[R0 : R0 + H, C0 : C0 + W])
        encode_mm_deflate(crop, tie, os.path.join(OUT, f"scene_crop_{name}.tif"))
        expected[name] = crop
    expected["geo_transform"] = np.array([tie[3], SCALE[0], 0.0, tie[4], 0.0, -SCALE[1]])
    expected["full_digests"] = np.array(digests)
    np.savez_compressed(os.path.join(OUT, "scene_crop_expected.npz"), **expected)
    print("\n".join(digests))


if __name__ == "__main__":
    main()
