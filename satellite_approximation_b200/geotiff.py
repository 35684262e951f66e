"""GeoTIFF band read / write either side of the Poisson path -- the host-side mirror of the reference's `utils::GeoTIFF<T>`
and `utils::GeoTiffWriter<T>` (lib/utils/include/utils/geotiff.h:98-195 writer, :204-263 reader), which `poisson_main`
uses to fetch bands 1-5 + the cloud band and to store the blended bands (executables/poisson-main.cpp:53-70).

The reference sits on GDAL, which this image does not have; SURVEY.md §8f-2 lists the GeoTIFF step as the data format next
to the path.  This module is a self-contained TIFF codec for what Sentinel-2 exports use (numpy + zlib; the LZW and
PackBits byte decoders are C, csrc/tiffcodec.c -> lib/libsattiff.so, with interpreter versions of the same as a stand-in):

  read   classic TIFF and BigTIFF, either byte order, strips or tiles, chunky or planar samples, 8/16/32/64-bit unsigned /
         signed / IEEE samples, compression none (1), deflate (8, 32946), LZW (5), PackBits (32773), predictor 1 / 2 / 3;
         geo tags (ModelPixelScale, ModelTiepoint, ModelTransformation, GeoKeyDirectory, GeoDoubleParams, GeoAsciiParams,
         GDAL_METADATA, GDAL_NODATA) are kept and give the GDAL-style affine geo transform
  write  what `GDALDriver::CreateCopy(template)` + `RasterIO(GF_Write)` produce in the reference: the template's size, band
         count, sample type and geo tags, template pixels for the bands that are not overwritten, uncompressed strips (or
         tiles), BigTIFF when the file would pass 4 GB

Layout (geotiff.h:234-253): the reference hands GDAL the data pointer of a *column-major* height x width
Eigen matrix and asks for width x height row-major samples, so `M(r, c) = raster_flat[r + c * height]` -- for a non-square
scene the matrix the reference blends is an index-scrambled image (a transpose for a square one), unscrambled again by the
writer, which passes the same pointer back.  Both behaviours are here and the caller chooses:

  layout="raster"     (default) the band as the image it is: array[r, c] = pixel (r, c).  Deliberate fix.
  layout="reference"  the reference's matrix: a Fortran-ordered height x width array over the row-major raster buffer.

Pure host code: nothing here touches the GPU, and nothing here is a fallback for it."""
from __future__ import annotations

import mmap
import os
import struct
import zlib
from typing import Iterable, Optional, Sequence, Union

import numpy as np

__all__ = ["TiffError", "TiffFile", "GeoTIFF", "GeoTiffWriter", "write_tiff", "gdal_convert"]

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 13: 4, 16: 8, 17: 8, 18: 8}

T_WIDTH, T_LENGTH, T_BITS, T_COMPRESSION, T_PHOTOMETRIC = 256, 257, 258, 259, 262
T_STRIP_OFFSETS, T_SPP, T_ROWS_PER_STRIP, T_STRIP_COUNTS = 273, 277, 278, 279
T_XRES, T_YRES, T_PLANAR, T_RESUNIT, T_PREDICTOR = 282, 283, 284, 296, 317
T_TILE_W, T_TILE_L, T_TILE_OFFSETS, T_TILE_COUNTS, T_EXTRA, T_SAMPLE_FORMAT = 322, 323, 324, 325, 338, 339
T_PIXEL_SCALE, T_TIEPOINT, T_TRANSFORM = 33550, 33922, 34264
T_GEOKEYS, T_GEODOUBLES, T_GEOASCII, T_GDAL_METADATA, T_GDAL_NODATA = 34735, 34736, 34737, 42112, 42113
GEO_TAGS = (T_PIXEL_SCALE, T_TIEPOINT, T_TRANSFORM, T_GEOKEYS, T_GEODOUBLES, T_GEOASCII, T_GDAL_METADATA, T_GDAL_NODATA)


class TiffError(IOError):
    """Unreadable / unsupported TIFF (the reference throws utils::IOError / std::runtime_error here, geotiff.h:221,248)."""


def _sample_dtype(bits: int, fmt: int) -> np.dtype:
    kind = {1: "u", 2: "i", 3: "f", 4: "u"}.get(fmt)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise TiffError(f"unsupported sample type: {bits} bits, SampleFormat {fmt}")
    return np.dtype(f"{kind}{bits // 8}")


def gdal_convert(a: np.ndarray, dtype) -> np.ndarray:
    """Sample conversion as GDAL's RasterIO does it between a file type and a buffer type (GDALCopyWords): to an integer
    type values are rounded half away from zero and clamped to the target range, NaN becomes 0; to a float type a plain
    cast.  This is what `GeoTIFF<f64>` on a u16 file (exact) and `GeoTiffWriter<f64>` into a u16 file (round + clamp) do."""
    dtype = np.dtype(dtype)
    a = np.asarray(a)
    if a.dtype == dtype:
        return a
    if dtype.kind == "f" or dtype.kind == "b":
        return a.astype(dtype)
    info = np.iinfo(dtype)
    if a.dtype.kind == "f" and dtype.itemsize <= 4:
        # the common case (a filled f64 band into a u16 / i16 / u8 file), done in cache-sized row blocks with in-place
        # passes: a whole 10980 x 10980 band through temporaries costs several seconds
        flat = a.reshape(-1) if a.flags.c_contiguous else np.ascontiguousarray(a).reshape(-1)
        out = np.empty(flat.shape, dtype)
        lo, hi = float(info.min), float(info.max)
        step = 1 << 18
        buf = np.empty(step, np.float64)
        half = np.empty(step, np.float64)
        for i in range(0, flat.size, step):
            src = flat[i : i + step]
            v, h_ = buf[: src.size], half[: src.size]
            np.copyto(v, src, casting="same_kind")
            nan = np.isnan(v)
            if nan.any():
                v[nan] = 0.0
            if info.min == 0:
                v += 0.5  # negatives clamp to 0 either way
                np.floor(v, out=v)
            else:
                np.copysign(0.5, v, out=h_)
                v += h_
                np.trunc(v, out=v)  # half away from zero
            np.clip(v, lo, hi, out=v)
            out[i : i + step] = v
        return out.reshape(a.shape)
    if a.dtype.kind == "f":
        v = np.where(np.isnan(a), 0.0, a).astype(np.float64)
        v = np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))
        # clamp in float, then fix the top end (2**64-1 and 2**63-1 are not representable in f64)
        out = np.clip(v, float(info.min), float(info.max))
        hi = out >= float(info.max)
        res = np.where(hi, 0, out).astype(dtype)
        res[hi] = info.max
        return res
    if a.dtype.kind == "b":
        return a.astype(dtype)
    lo = max(int(info.min), int(np.iinfo(a.dtype).min))
    hi = min(int(info.max), int(np.iinfo(a.dtype).max))
    return np.clip(a, lo, hi).astype(dtype)


# ---------------------------------------------------------------------------------------------------------------------
# decompressors


def _lzw_decode(data: bytes) -> bytes:
    """TIFF LZW (compression 5): MSB-first codes of 9..12 bits, ClearCode 256, EOI 257, the code width grows one code
    early ("early change"), as libtiff writes it."""
    out = bytearray()
    table: list[bytes] = []
    base = [bytes([i]) for i in range(256)] + [b"", b""]
    bitbuf = 0
    nbits = 0
    width = 9
    prev: Optional[bytes] = None
    pos = 0
    n = len(data)
    while True:
        while nbits < width and pos < n:
            bitbuf = (bitbuf << 8) | data[pos]
            pos += 1
            nbits += 8
        if nbits < width:
            break
        code = (bitbuf >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        bitbuf &= (1 << nbits) - 1
        if code == 256:
            table = list(base)
            width = 9
            prev = None
            continue
        if code == 257:
            break
        if not table:
            raise TiffError("LZW stream does not start with a clear code")
        if prev is None:
            if code >= 256:
                raise TiffError("corrupt LZW stream")
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(prev + entry[:1])
        elif code == len(table):
            entry = prev + prev[:1]
            table.append(entry)
        else:
            raise TiffError("corrupt LZW stream")
        out += entry
        prev = entry
        if len(table) + 1 >= (1 << width) and width < 12:
            width += 1
    return bytes(out)


def _packbits_decode(data: bytes) -> bytes:
    out = bytearray()
    i, n = 0, len(data)
    while i < n:
        h = data[i]
        i += 1
        if h < 128:
            out += data[i : i + h + 1]
            i += h + 1
        elif h > 128:
            out += data[i : i + 1] * (257 - h)
            i += 1
    return bytes(out)


_native = None


def _native_codec():
    """lib/libsattiff.so (csrc/tiffcodec.c): LZW / PackBits decoders in C.  Host-side I/O helper, optional: without it the
    interpreter versions above decode the same bytes, only slowly.  False once loading has failed."""
    global _native
    if _native is None:
        import ctypes as C

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libsattiff.so")
        try:
            lib = C.CDLL(path)
            for fn in (lib.st_lzw_decode, lib.st_packbits_decode):
                fn.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
                fn.restype = C.c_int
            _native = lib
        except (OSError, AttributeError):
            _native = False
    return _native


def _decode_native(fn, data: bytes, expected: int) -> bytes:
    import ctypes as C

    out = C.create_string_buffer(max(expected, 1))
    produced = C.c_size_t(0)
    rc = fn(data, len(data), out, expected, C.byref(produced))
    if rc != 0:
        raise TiffError("corrupt LZW stream" if rc == -1 else "LZW stream does not start with a clear code")
    return out.raw[: produced.value]


def _decompress(data: bytes, compression: int, expected: int) -> bytes:
    """`expected` = the decoded size of a full segment (an upper bound: the last strip may be shorter)."""
    if compression == 1:
        return data
    if compression in (8, 32946):
        try:
            return zlib.decompress(data)
        except zlib.error as e:
            raise TiffError(f"corrupt deflate stream: {e}") from e
    if compression in (5, 32773):
        lib = _native_codec()
        if lib:
            return _decode_native(lib.st_lzw_decode if compression == 5 else lib.st_packbits_decode, data, expected)
        return _lzw_decode(data) if compression == 5 else _packbits_decode(data)
    raise TiffError(f"unsupported TIFF compression {compression}")


# ---------------------------------------------------------------------------------------------------------------------
# reader


class TiffFile:
    """First image file directory of a TIFF / BigTIFF: tags, geometry, and band decoding."""

    decode_threads = min(8, os.cpu_count() or 1)  # segments decoded concurrently (class-wide knob)

    def __init__(self, path):
        self.path = os.fspath(path)
        try:
            with open(self.path, "rb") as f:
                # mapped, not read: a 13-band Sentinel-2 tile is gigabytes, and only the segments of the bands asked for
                # are ever touched
                self._buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if os.fstat(f.fileno()).st_size else b""
        except (OSError, ValueError) as e:
            raise TiffError(f"Failed to open {self.path}: {e}") from e
        b = self._buf
        if len(b) < 8 or b[:2] not in (b"II", b"MM"):
            raise TiffError(f"{self.path}: not a TIFF file")
        self.byteorder = "<" if b[:2] == b"II" else ">"
        magic = struct.unpack(self.byteorder + "H", b[2:4])[0]
        if magic == 42:
            self.bigtiff = False
            ifd = struct.unpack(self.byteorder + "I", b[4:8])[0]
        elif magic == 43:
            self.bigtiff = True
            osz, zero, ifd = struct.unpack(self.byteorder + "HHQ", b[4:16])
            if osz != 8 or zero != 0:
                raise TiffError(f"{self.path}: malformed BigTIFF header")
        else:
            raise TiffError(f"{self.path}: not a TIFF file (magic {magic})")
        self.tags: dict[int, tuple[int, object]] = {}
        try:
            self._read_ifd(ifd)
        except struct.error as e:
            raise TiffError(f"{self.path}: truncated image file directory") from e
        g = self._scalar
        self.width = int(g(T_WIDTH))
        self.height = int(g(T_LENGTH))
        self.samples_per_pixel = int(g(T_SPP, 1))
        bits = np.atleast_1d(self._value(T_BITS, [1]))
        fmts = np.atleast_1d(self._value(T_SAMPLE_FORMAT, [1]))
        if len(set(int(x) for x in bits)) != 1 or len(set(int(x) for x in fmts)) != 1:
            raise TiffError(f"{self.path}: samples of mixed type are not supported")
        self.dtype = _sample_dtype(int(bits[0]), int(fmts[0]))
        self.compression = int(g(T_COMPRESSION, 1))
        self.predictor = int(g(T_PREDICTOR, 1))
        self.planar = int(g(T_PLANAR, 1))
        self.photometric = int(g(T_PHOTOMETRIC, 1))
        if T_TILE_W in self.tags:
            self.tiled = True
            self.seg_w = int(g(T_TILE_W))
            self.seg_h = int(g(T_TILE_L))
            self._offsets = np.atleast_1d(self._value(T_TILE_OFFSETS)).astype(np.int64)
            self._counts = np.atleast_1d(self._value(T_TILE_COUNTS)).astype(np.int64)
        else:
            self.tiled = False
            self.seg_w = self.width
            rps = int(g(T_ROWS_PER_STRIP, self.height))
            self.seg_h = min(rps, self.height) if rps > 0 else self.height
            if T_STRIP_OFFSETS not in self.tags:
                raise TiffError(f"{self.path}: no strip or tile offsets")
            self._offsets = np.atleast_1d(self._value(T_STRIP_OFFSETS)).astype(np.int64)
            if T_STRIP_COUNTS in self.tags:
                self._counts = np.atleast_1d(self._value(T_STRIP_COUNTS)).astype(np.int64)
            else:  # allowed for one uncompressed strip
                self._counts = np.full(len(self._offsets), len(b), dtype=np.int64) - self._offsets
        if self.width <= 0 or self.height <= 0 or self.seg_w <= 0 or self.seg_h <= 0:
            raise TiffError(f"{self.path}: empty image")
        self._segs_x = -(-self.width // self.seg_w)
        self._segs_y = -(-self.height // self.seg_h)
        per_plane = self._segs_x * self._segs_y
        need = per_plane * (self.samples_per_pixel if self.planar == 2 else 1)
        if len(self._offsets) < need or len(self._counts) < need:
            raise TiffError(f"{self.path}: {len(self._offsets)} segments listed, {need} needed")

    def close(self) -> None:
        """Drop the file mapping (it also goes with the object)."""
        buf, self._buf = self._buf, b""
        if isinstance(buf, mmap.mmap):
            buf.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- tags ----------------------------------------------------------------------------------------------------------
    def _read_ifd(self, off: int) -> None:
        b, bo = self._buf, self.byteorder
        if self.bigtiff:
            (n,) = struct.unpack_from(bo + "Q", b, off)
            off += 8
            esz, inl, hdr = 20, 8, "HHQ"
        else:
            (n,) = struct.unpack_from(bo + "H", b, off)
            off += 2
            esz, inl, hdr = 12, 4, "HHI"
        for i in range(n):
            e = off + i * esz
            tag, typ, cnt = struct.unpack_from(bo + hdr, b, e)
            if typ not in _TYPE_SIZE:
                continue
            nbytes = _TYPE_SIZE[typ] * cnt
            voff = e + esz - inl
            if nbytes > inl:
                (voff,) = struct.unpack_from(bo + ("Q" if self.bigtiff else "I"), b, voff)
            raw = b[voff : voff + nbytes]
            if len(raw) != nbytes:
                raise TiffError(f"{self.path}: tag {tag} points outside the file")
            self.tags[tag] = (typ, self._decode_tag(typ, cnt, raw))

    def _decode_tag(self, typ: int, cnt: int, raw: bytes):
        bo = self.byteorder
        if typ == 2:
            return raw.rstrip(b"\0").decode("latin-1")
        if typ == 7:
            return raw
        if typ in (5, 10):
            a = np.frombuffer(raw, dtype=np.dtype(bo + ("u4" if typ == 5 else "i4"))).reshape(-1, 2)
            return a.astype(np.int64)
        code = {1: "u1", 3: "u2", 4: "u4", 6: "i1", 8: "i2", 9: "i4", 11: "f4", 12: "f8", 13: "u4", 16: "u8", 17: "i8",
                18: "u8"}[typ]  # fmt: skip
        return np.frombuffer(raw, dtype=np.dtype(bo + code)).astype(np.dtype(code).newbyteorder("="))

    def _value(self, tag: int, default=None):
        if tag not in self.tags:
            if default is None:
                raise TiffError(f"{self.path}: required tag {tag} is missing")
            return default
        return self.tags[tag][1]

    def _scalar(self, tag: int, default=None):
        v = self._value(tag, default)
        return v if np.isscalar(v) else np.atleast_1d(v)[0]

    @property
    def geo_transform(self) -> Optional[tuple[float, ...]]:
        """GDAL's affine transform (x0, dx, rx, y0, ry, dy) from the GeoTIFF tags, None when the file carries none (the
        reference's constructor throws IOError then, geotiff.h:220-222)."""
        if T_TRANSFORM in self.tags:
            m = np.asarray(self._value(T_TRANSFORM), dtype=np.float64)
            if m.size >= 8:
                return (float(m[3]), float(m[0]), float(m[1]), float(m[7]), float(m[4]), float(m[5]))
        if T_PIXEL_SCALE in self.tags and T_TIEPOINT in self.tags:
            s = np.asarray(self._value(T_PIXEL_SCALE), dtype=np.float64)
            t = np.asarray(self._value(T_TIEPOINT), dtype=np.float64)
            if s.size >= 2 and t.size >= 6:
                return (float(t[3] - t[0] * s[0]), float(s[0]), 0.0, float(t[4] + t[1] * s[1]), 0.0, float(-s[1]))
        return None

    # -- pixels --------------------------------------------------------------------------------------------------------
    def _segment(self, index: int, nsamp: int) -> np.ndarray:
        """Decoded segment `index` as (seg_h, seg_w, nsamp) in native byte order (strips may be short at the bottom)."""
        off, cnt = int(self._offsets[index]), int(self._counts[index])
        raw = self._buf[off : off + cnt]
        if len(raw) != cnt:
            raise TiffError(f"{self.path}: segment {index} points outside the file")
        isz = self.dtype.itemsize
        row_bytes = self.seg_w * nsamp * isz
        data = _decompress(raw, self.compression, self.seg_h * row_bytes)
        rows = min(self.seg_h, len(data) // row_bytes) if row_bytes else 0
        if rows <= 0:
            raise TiffError(f"{self.path}: segment {index} is truncated")
        data = data[: rows * row_bytes]
        if self.predictor == 3:
            if self.dtype.kind != "f":
                raise TiffError("floating-point predictor on integer samples")
            # bytes of each row are differenced, then stored most significant byte plane first
            u = np.frombuffer(data, dtype=np.uint8).reshape(rows, row_bytes // nsamp, nsamp)
            acc = np.cumsum(u, axis=1, dtype=np.uint8)  # byte differences with a stride of one pixel, modulo 256
            planes = acc.reshape(rows, isz, self.seg_w * nsamp)  # plane 0 = most significant byte
            be = np.ascontiguousarray(planes.transpose(0, 2, 1))
            a = be.view(np.dtype(">" + self.dtype.str[1:])).reshape(rows, self.seg_w, nsamp)
            return a.astype(self.dtype)
        a = np.frombuffer(data, dtype=self.dtype.newbyteorder(self.byteorder)).reshape(rows, self.seg_w, nsamp)
        a = a.astype(self.dtype.newbyteorder("="))
        if self.predictor == 2:
            if self.dtype.kind == "f":
                raise TiffError("horizontal predictor on floating-point samples")
            a = np.cumsum(a, axis=1, dtype=a.dtype)
        elif self.predictor != 1:
            raise TiffError(f"unsupported TIFF predictor {self.predictor}")
        return a

    def _fill(self, out: np.ndarray, base: int, nsamp: int, pick: Optional[int]) -> None:
        """Decode the segments of one plane into `out` ((H, W) with `pick`, (H, W, nsamp) without).  Segments are
        independent, and zlib, the C decoders and numpy all release the GIL, so they are decoded on a small thread pool
        when the file is big enough for that to pay."""

        def one(k: int) -> None:
            sy, sx = divmod(k, self._segs_x)
            r0, c0 = sy * self.seg_h, sx * self.seg_w
            seg = self._segment(base + k, nsamp)
            h = min(self.seg_h, self.height - r0)
            w = min(self.seg_w, self.width - c0)
            if seg.shape[0] < h:
                raise TiffError(f"{self.path}: segment {sy},{sx} is truncated")
            out[r0 : r0 + h, c0 : c0 + w] = seg[:h, :w, pick] if pick is not None else seg[:h, :w]

        n = self._segs_x * self._segs_y
        workers = min(self.decode_threads, n)
        if workers <= 1 or out.nbytes < (4 << 20):
            for k in range(n):
                one(k)
            return
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(workers, thread_name_prefix="satfill-tiff") as pool:
            for _ in pool.map(one, range(n), chunksize=max(1, n // (8 * workers))):
                pass

    def read_band(self, band: int) -> np.ndarray:
        """Band `band` (1-based, GDAL numbering) as a C-ordered (height, width) array of the file's sample type."""
        spp = self.samples_per_pixel
        if not 1 <= band <= spp:
            raise TiffError(f"{self.path}: band {band} out of range 1..{spp}")
        out = np.empty((self.height, self.width), dtype=self.dtype.newbyteorder("="))
        per_plane = self._segs_x * self._segs_y
        if self.planar == 2:
            self._fill(out, (band - 1) * per_plane, 1, 0)
        else:
            self._fill(out, 0, spp, band - 1)
        return out

    def read_all(self) -> list[np.ndarray]:
        if self.planar == 2 or self.samples_per_pixel == 1:
            return [self.read_band(b + 1) for b in range(self.samples_per_pixel)]
        # chunky: decode every segment once
        spp = self.samples_per_pixel
        out = np.empty((self.height, self.width, spp), dtype=self.dtype.newbyteorder("="))
        self._fill(out, 0, spp, None)
        return [np.ascontiguousarray(out[:, :, b]) for b in range(spp)]


def _to_layout(raster: np.ndarray, layout: str) -> np.ndarray:
    if layout == "raster":
        return raster
    if layout == "reference":
        h, w = raster.shape
        return np.ascontiguousarray(raster).reshape(-1).reshape((h, w), order="F")
    raise ValueError("layout must be 'raster' or 'reference'")


def _from_layout(values: np.ndarray, layout: str) -> np.ndarray:
    values = np.asarray(values)
    if values.ndim != 2:
        raise ValueError("a band is a 2-D array")
    if layout == "raster":
        return values
    if layout == "reference":
        h, w = values.shape
        return values.reshape(-1, order="F").reshape(h, w)
    raise ValueError("layout must be 'raster' or 'reference'")


class GeoTIFF:
    """Mirror of `utils::GeoTIFF<ScalarT>` (geotiff.h:204-263): `GeoTIFF(path, dtype)` opens the file, `read(band)` /
    `read([bands])` / `read()` return bands converted to `dtype` the way GDAL's RasterIO converts (gdal_convert).

    `read()` with no argument returns every band; the reference loops band numbers 0..count-1 there (geotiff.h:267-274),
    which GDAL rejects for band 0 -- deliberately not mirrored."""

    def __init__(self, path, dtype=np.float64, layout: str = "raster"):
        if layout not in ("raster", "reference"):
            raise ValueError("layout must be 'raster' or 'reference'")
        self.path = os.fspath(path)
        self.dtype = np.dtype(dtype)
        self.layout = layout
        self.file = TiffFile(self.path)
        self.width = self.file.width
        self.height = self.file.height
        gt = self.file.geo_transform
        if gt is None:
            raise TiffError(f"Unable to load the geo transformation information: {self.path}")
        self.geo_transform = gt

    @property
    def raster_count(self) -> int:
        return self.file.samples_per_pixel

    def read(self, bands: Union[None, int, Iterable[int]] = None):
        if bands is None:
            return [_to_layout(gdal_convert(b, self.dtype), self.layout) for b in self.file.read_all()]
        if isinstance(bands, (int, np.integer)):
            return _to_layout(gdal_convert(self.file.read_band(int(bands)), self.dtype), self.layout)
        return [self.read(int(b)) for b in bands]

    def pixel_to_geo(self, row: float, col: float) -> tuple[float, float]:
        g = self.geo_transform
        return (g[0] + col * g[1] + row * g[2], g[3] + col * g[4] + row * g[5])

    # -- geo-referencing helpers (geotiff.h:322-421; positions are LatLng = (north-south, east-west), north-up assumed) -----
    def east_west_step(self) -> float:
        return self.geo_transform[1]

    def north_south_step(self) -> float:
        return self.geo_transform[5]

    def north(self) -> float:
        return self.geo_transform[3]

    def west(self) -> float:
        return self.geo_transform[0]

    def south(self) -> float:
        return self.geo_transform[3] + self.height * self.north_south_step()

    def east(self) -> float:
        return self.geo_transform[0] + self.width * self.east_west_step()

    def north_west(self) -> tuple[float, float]:
        return (self.north(), self.west())

    def north_east(self) -> tuple[float, float]:
        return (self.north(), self.east())

    def south_east(self) -> tuple[float, float]:
        return (self.south(), self.east())

    def south_west(self) -> tuple[float, float]:
        return (self.south(), self.west())

    def index_at(self, pos: Sequence[float]) -> tuple[int, int]:
        """(x, y) = (column, row) of the pixel under `pos`, clamped to the image (geotiff.h:383-391; the C++ cast
        truncates towards zero)."""
        x = int((pos[1] - self.west()) / self.east_west_step())
        y = int((pos[0] - self.north()) / self.north_south_step())
        return (min(max(x, 0), self.width - 1), min(max(y, 0), self.height - 1))

    def value_at(self, pos: Sequence[float], values: np.ndarray):
        x, y = self.index_at(pos)
        return values[y, x]

    def uv_at(self, pos: Sequence[float]) -> tuple[float, float]:
        x, y = self.index_at(pos)
        return (x / self.width, y / self.height)

    def mid_point_of_pixel(self, index: Sequence[int]) -> tuple[float, float]:
        """geotiff.h:393-398, as written there: index[0] steps north-south, index[1] east-west."""
        return (self.north() + self.north_south_step() * float(index[0]) + self.north_south_step() * 0.5,
                self.west() + self.east_west_step() * float(index[1]) + self.east_west_step() * 0.5)  # fmt: skip

    def bilinear_value_at(self, pos: Sequence[float], values: np.ndarray):
        """geotiff.h:343-372.  On a pixel centre line (x or y integral) the reference divides by zero (inf * 0 = NaN for
        floating samples); mirrored as is."""
        x = (pos[1] - self.west()) / self.east_west_step()
        y = (pos[0] - self.north()) / self.north_south_step()
        x1, x2, y1, y2 = np.floor(x), np.ceil(x), np.floor(y), np.ceil(y)

        def value(fx, fy):
            xi = min(max(int(fx), 0), self.width - 1)
            yi = min(max(int(fy), 0), self.height - 1)
            return values[yi, xi]

        m = np.array([[value(x1, y1), value(x1, y2)], [value(x2, y1), value(x2, y2)]], dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.float64(1.0) / np.float64((x2 - x1) * (y2 - y1))
            r = s * (np.array([x2 - x, x - x1]) @ (m @ np.array([y2 - y, y - y1])))
        return np.asarray(values).dtype.type(r) if np.asarray(values).dtype.kind == "f" else r

    @staticmethod
    def value_domain(values: np.ndarray, dtype=np.float64) -> tuple:
        """geotiff.h:400-407: (min, max) of a band cast to `dtype`."""
        t = np.dtype(dtype).type
        return (t(np.min(values)), t(np.max(values)))

    @staticmethod
    def dem_value_domain(values: np.ndarray, dtype=np.float64) -> tuple:
        """geotiff.h:409-421: as value_domain, ignoring the DEM no-data sentinel (<= -32767): those samples are lifted to
        the maximum before the minimum is taken."""
        v = np.asarray(values)
        hi = np.max(v)
        sel = (v <= -32767.0).astype(v.dtype)
        tmp = v + 32767 * sel + hi * sel
        t = np.dtype(dtype).type
        return (t(np.min(tmp)), t(hi))

    def write(self, matrix: np.ndarray, destination, band_index: int = 1) -> None:
        """GeoTIFF::write(cv::Mat, path, bandIndex) (geotiff.h:276-320): a copy of this file with band `band_index`
        replaced by the row-major `matrix`."""
        GeoTiffWriter([np.asarray(matrix)], self.path).write(destination, start_index=band_index)


# ---------------------------------------------------------------------------------------------------------------------
# writer


def _pack_tag_value(typ: int, value) -> tuple[int, bytes]:
    if typ == 2:
        raw = value.encode("latin-1") + b"\0"
        return len(raw), raw
    if typ == 7:
        return len(value), bytes(value)
    if typ in (5, 10):
        a = np.asarray(value, dtype=np.int64).reshape(-1, 2)
        return len(a), a.astype("<u4" if typ == 5 else "<i4").tobytes()
    code = {1: "u1", 3: "u2", 4: "u4", 6: "i1", 8: "i2", 9: "i4", 11: "f4", 12: "f8", 13: "u4", 16: "u8", 17: "i8",
            18: "u8"}[typ]  # fmt: skip
    a = np.atleast_1d(np.asarray(value)).astype("<" + code)
    return a.size, a.tobytes()


def write_tiff(path, bands: Sequence[np.ndarray], extra_tags: Optional[dict[int, tuple[int, object]]] = None,
               tile: Optional[tuple[int, int]] = None, rows_per_strip: Optional[int] = None, compress: bool = False,
               bigtiff: Optional[bool] = None) -> None:  # fmt: skip
    """Write C-ordered (height, width) bands of one dtype as a little-endian TIFF with planar sample layout for more than
    one band: uncompressed (what CreateCopy with no creation options produces) or deflate; strips, or tiles of
    `tile=(rows, cols)` (multiples of 16).  BigTIFF is chosen automatically above 4 GB.  `extra_tags` maps tag ->
    (TIFF type, value), e.g. the geo tags of a template."""
    bands = [np.ascontiguousarray(b) for b in bands]
    if not bands:
        raise ValueError("no bands to write")
    h, w = bands[0].shape
    dt = bands[0].dtype
    if any(b.shape != (h, w) or b.dtype != dt for b in bands):
        raise ValueError("bands differ in shape or type")
    if h == 0 or w == 0:
        raise ValueError("empty image")
    if dt.kind not in "uif" or dt.itemsize not in (1, 2, 4, 8) or (dt.kind == "f" and dt.itemsize < 4):
        raise ValueError(f"unsupported sample type {dt}")
    le = dt.newbyteorder("<")
    if tile is not None:
        th, tw = tile
        if th % 16 or tw % 16 or th <= 0 or tw <= 0:
            raise ValueError("tile sides must be positive multiples of 16")
        raw_sizes = [th * tw * dt.itemsize] * (len(bands) * (-(-h // th)) * (-(-w // tw)))
    else:
        if rows_per_strip is None:
            rows_per_strip = max(1, min(h, (1 << 20) // max(1, w * dt.itemsize)))
        raw_sizes = [min(rows_per_strip, h - r0) * w * dt.itemsize for _ in bands for r0 in range(0, h, rows_per_strip)]

    def raw_segments():
        """The uncompressed segments in file order, produced one at a time (a 13-band tile is never held twice)."""
        for b in bands:
            if tile is not None:
                for r0 in range(0, h, th):
                    for c0 in range(0, w, tw):
                        t = np.zeros((th, tw), dtype=le)
                        blk = b[r0 : r0 + th, c0 : c0 + tw]
                        t[: blk.shape[0], : blk.shape[1]] = blk
                        yield t.tobytes()
            else:
                for r0 in range(0, h, rows_per_strip):  # a view of the band's own memory where the byte order allows
                    yield b[r0 : r0 + rows_per_strip].astype(le, copy=False).reshape(-1).view(np.uint8).data

    if compress:
        held = [zlib.compress(seg, 6) for seg in raw_segments()]  # sizes are only known afterwards: hold the compressed form
        counts = [len(c) for c in held]
        segments = lambda: iter(held)  # noqa: E731
    else:
        counts = raw_sizes
        segments = raw_segments
    payload = sum(c + (c & 1) for c in counts)
    if bigtiff is None:
        bigtiff = payload + (1 << 20) + 16 * len(counts) >= (1 << 32)
    nb = len(bands)
    fmt = {"u": 1, "i": 2, "f": 3}[dt.kind]
    off_t = 16 if bigtiff else 4
    tags: dict[int, tuple[int, object]] = {
        T_WIDTH: (4, w), T_LENGTH: (4, h), T_BITS: (3, [dt.itemsize * 8] * nb), T_COMPRESSION: (3, 8 if compress else 1),
        T_PHOTOMETRIC: (3, 1), T_SPP: (3, nb), T_PLANAR: (3, 2 if nb > 1 else 1), T_SAMPLE_FORMAT: (3, [fmt] * nb),
    }  # fmt: skip
    if nb > 1:
        tags[T_EXTRA] = (3, [0] * (nb - 1))
    for t, tv in (extra_tags or {}).items():
        if t not in tags and t not in (T_STRIP_OFFSETS, T_STRIP_COUNTS, T_ROWS_PER_STRIP, T_TILE_W, T_TILE_L, T_TILE_OFFSETS,
                                       T_TILE_COUNTS, T_PREDICTOR, T_EXTRA):  # fmt: skip
            tags[t] = tv
    header = 16 if bigtiff else 8
    offsets, pos = [], header
    for c in counts:
        offsets.append(pos)
        pos += c + (c & 1)
    if tile is not None:
        tags[T_TILE_W], tags[T_TILE_L] = (4, tile[1]), (4, tile[0])
        tags[T_TILE_OFFSETS], tags[T_TILE_COUNTS] = (off_t, offsets), (off_t, counts)
    else:
        tags[T_ROWS_PER_STRIP] = (4, rows_per_strip)
        tags[T_STRIP_OFFSETS], tags[T_STRIP_COUNTS] = (off_t, offsets), (off_t, counts)
    ifd_off = pos
    esz, inl = (20, 8) if bigtiff else (12, 4)
    n = len(tags)
    ifd_size = (8 if bigtiff else 2) + n * esz + (8 if bigtiff else 4)
    extra = bytearray()
    entries = bytearray()
    for t in sorted(tags):
        typ, val = tags[t]
        cnt, raw = _pack_tag_value(typ, val)
        if len(raw) <= inl:
            field = raw.ljust(inl, b"\0")
        else:
            o = ifd_off + ifd_size + len(extra)
            field = struct.pack("<Q" if bigtiff else "<I", o)
            extra += raw
            if len(extra) & 1:
                extra += b"\0"
        entries += struct.pack("<HHQ" if bigtiff else "<HHI", t, typ, cnt) + field
    if not bigtiff and ifd_off + ifd_size + len(extra) >= (1 << 32):
        raise ValueError("file too large for classic TIFF; pass bigtiff=True")
    path = os.path.abspath(os.fspath(path))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    # written next to the destination and renamed into place: a reader that has the old file mapped (the writer's own
    # template, when source and destination are one path) keeps a valid mapping, and no half-written file is ever visible
    tmp = f"{path}.part{os.getpid()}"
    try:
        _write_file(tmp, bigtiff, ifd_off, segments, counts, n, entries, extra)
        os.replace(tmp, path)
    except BaseException:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise


def _write_file(path, bigtiff, ifd_off, segments, counts, n, entries, extra) -> None:
    with open(path, "wb") as f:
        if bigtiff:
            f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd_off))
        else:
            f.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        for seg, c in zip(segments(), counts):
            if len(seg) != c:
                raise RuntimeError("write_tiff: segment size does not match the directory")  # internal invariant
            f.write(seg)
            if c & 1:
                f.write(b"\0")
        f.write(struct.pack("<Q" if bigtiff else "<H", n))
        f.write(entries)
        f.write(struct.pack("<Q" if bigtiff else "<I", 0))
        f.write(extra)


class GeoTiffWriter:
    """Mirror of `utils::GeoTiffWriter<ScalarT>` (geotiff.h:98-195): built from the values (one band or a list of bands)
    and the path of a template file; `write(destination, start_index=1)` makes a copy of the template (size, band count,
    sample type, geo tags, pixels) and overwrites bands start_index, start_index+1, ... with the values, converted to the
    file's sample type as GDAL's RasterIO does."""

    def __init__(self, values, template_path, layout: str = "raster"):
        if layout not in ("raster", "reference"):
            raise ValueError("layout must be 'raster' or 'reference'")
        self.single = isinstance(values, np.ndarray) and values.ndim == 2
        self.values = [values] if self.single else list(values)
        self.layout = layout
        self.template = TiffFile(template_path)
        self.width = self.template.width
        self.height = self.template.height

    def write(self, destination, start_index: int = 1) -> None:
        t = self.template
        if self.single:
            start_index = 1  # the single-band form always writes band 1 (geotiff.h:160-163)
        if start_index < 1 or start_index - 1 + len(self.values) > t.samples_per_pixel:
            raise RuntimeError("Unable to write raster image")  # GDAL: null band -> the reference crashes / throws
        file_dtype = t.dtype.newbyteorder("=")
        new = {}
        for i, v in enumerate(self.values):
            r = _from_layout(v, self.layout)
            if r.shape != (self.height, self.width):
                raise RuntimeError("Unable to write raster image")
            new[start_index - 1 + i] = np.ascontiguousarray(gdal_convert(r, file_dtype))
        # the template's pixels are only needed for the bands that are not overwritten
        kept = [b for b in range(t.samples_per_pixel) if b not in new]
        old = dict(zip(range(t.samples_per_pixel), t.read_all())) if t.planar == 1 and len(kept) > 1 else {
            b: t.read_band(b + 1) for b in kept}  # chunky samples: decode every segment once, not once per band
        bands = [new[b] if b in new else old[b] for b in range(t.samples_per_pixel)]
        keep = {k: v for k, v in t.tags.items() if k in GEO_TAGS or k in (T_XRES, T_YRES, T_RESUNIT)}
        write_tiff(destination, bands, extra_tags=keep)
