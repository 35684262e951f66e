"""The reference's two fill executables and its image file helpers, restated on the GPU path:

  laplace_main   executables/laplace-main.cpp:12-42    <base_image> <invalid_image> <output_path>
  poisson_main   executables/poisson-main.cpp:23-72    <input.tif> <replacement.tif>
  read_image / write_image   lib/approx/source/utils.cpp:16-68

    python -m satellite_approximation_b200.drivers laplace_main  base.png marked.png out.png
    python -m satellite_approximation_b200.drivers poisson_main  input.tif replacement.tif [--reference-layout]
    python -m satellite_approximation_b200.drivers fill_folder   base_folder --bands B02,B03,B04 [--poisson]
        [--no-cache] [--skip-threshold 0.5] [--distance-weight 0.5]      (the commented-out fill_missing_data_folder,
        lib/approx/source/laplace.cpp:170-244; under torchrun every rank takes every WORLD_SIZE-th date folder)

Pixels go through the C-ABI on the GPU (apply_laplace -> sa_apply_laplace_u8, preprocess_cloud_band ->
sa_morph_close_mask, blend_images_poisson -> sa_poisson_blend); there is no CPU solve here.  GeoTIFFs are read and
written by geotiff.py; PNG / JPEG files by OpenCV (cv2) when it is importable, else Pillow."""
from __future__ import annotations

import logging
import os
import sys
from typing import Optional, Sequence

import numpy as np

from . import geotiff

__all__ = ["imread_color", "imwrite", "read_image", "write_image", "laplace_main", "poisson_main", "fill_folder_main", "main"]

_log = logging.getLogger("approx")


class IOError_(IOError):
    """utils::IOError (lib/utils/include/utils/error.h:22-32)."""


def imread_color(path) -> Optional[np.ndarray]:
    """cv::imread(path, IMREAD_COLOR): uint8 H x W x 3 in B, G, R order, None when the file cannot be decoded."""
    path = os.fspath(path)
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        return cv2.imread(path, cv2.IMREAD_COLOR)
    try:
        from PIL import Image

        with Image.open(path) as im:
            return np.ascontiguousarray(np.asarray(im.convert("RGB"))[:, :, ::-1])
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("neither cv2 nor PIL is importable: cannot decode image files") from e
    except OSError:
        return None


def saturate_u8(image: np.ndarray) -> np.ndarray:
    """What cv::imwrite does to a non-8-bit matrix for PNG / JPEG: convertTo(CV_8U) = saturate_cast<uchar>(cvRound(v))
    (round half to even, clamp to 0..255)."""
    a = np.asarray(image)
    if a.dtype == np.uint8:
        return a
    return np.clip(np.rint(np.nan_to_num(a.astype(np.float64), nan=0.0)), 0, 255).astype(np.uint8)


def imwrite(path, image_bgr: np.ndarray) -> bool:
    """cv::imwrite for the drivers: B, G, R (or single-channel) image, any depth (saturate_u8)."""
    path = os.fspath(path)
    img = np.ascontiguousarray(saturate_u8(image_bgr))
    try:
        import cv2
    except ImportError:
        cv2 = None
    if cv2 is not None:
        return bool(cv2.imwrite(path, img))
    from PIL import Image

    Image.fromarray(img[:, :, ::-1] if img.ndim == 3 else img).save(path)
    return True


def read_image(path) -> list[np.ndarray]:
    """approx::read_image (utils.cpp:16-34): three float64 channels R, G, B in [0, 1], gamma-decoded; IOError when the
    file cannot be opened."""
    import satellite_approximation_b200 as sab

    image = imread_color(path)
    if image is None or image.size == 0:
        raise IOError_(f"Failed to open image: {os.fspath(path)}")
    return sab.image_to_channels(image)


def write_image(channels: Sequence[np.ndarray], output_path) -> None:
    """approx::write_image (utils.cpp:62-68): gamma-encode three channels and store them; anything but three channels is
    logged and nothing is written."""
    import satellite_approximation_b200 as sab

    image = sab.channels_to_image(channels)
    if image is None:
        return
    imwrite(output_path, image)


def laplace_main(argv: Sequence[str]) -> int:
    """laplace-main.cpp:12-42: fill the red-marked area of <base_image> (marks in <invalid_image>: R >= 220 and
    G <= 150, laplace.cpp:140-150) channel by channel and store the result."""
    import satellite_approximation_b200 as sab

    if len(argv) != 3:
        _log.error("Usage: laplace_main <base_image> <invalid_image> <output_path>")
        return -1
    file, replacement_file, output_path = argv
    for p in (file, replacement_file):
        if not os.path.exists(p):
            _log.error("%s does not exist", p)
            return -1
    image = imread_color(file)
    invalid_areas = imread_color(replacement_file)
    if image is None or invalid_areas is None:
        _log.error("could not decode %s", file if image is None else replacement_file)
        return -1
    _log.info("Starting laplace")
    res = sab.apply_laplace(image, invalid_areas, 220)  # RuntimeError on a size mismatch, like the reference
    _log.info("Finished. Writing file")
    imwrite(output_path, res)
    return 0


def poisson_main(argv: Sequence[str]) -> int:
    """poisson-main.cpp:23-72: bands 1-5 of <input> are blended against bands 1-5 of <replacement> inside the mask made
    from band 6 of <input> by an 11 x 11 morphological close; the result is a copy of <input> with bands 1-5 replaced,
    stored as <dir of input>/poisson_simple_replace/<name of input>.

    `--reference-layout` reproduces the reference's buffer handling (geotiff.py: a column-major matrix over the row-major
    raster, i.e. an index-scrambled image for a non-square scene); the default treats the raster as the image it is."""
    args = [a for a in argv if a != "--reference-layout"]
    layout = "reference" if len(args) != len(argv) else "raster"
    if len(args) != 2:
        _log.info("Usage: poisson_main input_path replacement_path [--reference-layout]")
        return -1
    input_path, replacement_path = args
    for p in (input_path, replacement_path):
        if not os.path.exists(p):
            _log.error("%s does not exist", p)
            return -1
    bands = [1, 2, 3, 4, 5]
    cloud_band = 6
    tiff = geotiff.GeoTIFF(input_path, np.float64, layout=layout)
    input_bands = tiff.read(bands)
    cloud = tiff.read(cloud_band)
    cloudmask = _close_mask(cloud)
    # one memory order for everything that goes to the C-ABI (the mask is a byte per pixel: cheap to re-lay if it differs)
    cloudmask = np.asfortranarray(cloudmask) if layout == "reference" else np.ascontiguousarray(cloudmask)
    _log.info("Finished close + dilate")
    replacement_bands = geotiff.GeoTIFF(replacement_path, np.float64, layout=layout).read(bands)
    _log.info("Starting solver...")
    # blend_images_poisson's vector overload (poisson.cpp:292-303) returns the inputs unchanged when the sizes differ or a
    # band does not converge; the same happens here, in place (the library writes nothing unless every band converged)
    if any(a.shape != input_bands[0].shape for a in input_bands + replacement_bands):
        _log.error("Input and replacement images must have the same dimensions")  # poisson.cpp:154-157
    elif not _blend_in_place(input_bands, replacement_bands, cloudmask):
        _log.error("Failed to solve the linear system (no convergence)")  # poisson.cpp:263-269
    _log.info("Finished solving. Writing results")
    dest = os.path.join(os.path.dirname(os.path.abspath(input_path)), "poisson_simple_replace", os.path.basename(input_path))
    geotiff.GeoTiffWriter(input_bands, input_path, layout=layout).write(dest)
    return 0


def _close_mask(cloud: np.ndarray) -> np.ndarray:
    """preprocess_cloud_band (poisson-main.cpp:10-21) on the GPU; the mask comes back in the band's memory order."""
    import satellite_approximation_b200 as sab

    return sab.preprocess_cloud_band(cloud)


def _blend_in_place(bands: list, replacements: list, mask: np.ndarray) -> bool:
    """approx::blend_images_poisson at its defaults (tolerance 1e-6, n / 2 iterations) on the GPU, in place on `bands`: the
    decoded bands are handed to the C-ABI as they lie (either memory order), without the Fortran-ordered copies the
    reference-shaped Python function makes.  False when a band did not converge (nothing was written then)."""
    import satellite_approximation_b200 as sab

    stats = sab.default_context().poisson_blend(bands, replacements, mask, tolerance=1e-6, max_iterations=None,
                                                precond=sab._defaults["precond"], check_every=sab._defaults["check_every"])  # fmt: skip
    return all(s["status"] != sab.SA_NOT_CONVERGED for s in stats)


def fill_folder_main(argv: Sequence[str]) -> int:
    import argparse

    from . import scenes

    ap = argparse.ArgumentParser(prog="fill_folder")
    ap.add_argument("base_folder")
    ap.add_argument("--bands", required=True, help="comma-separated band names (files <band>.tif in every date folder)")
    ap.add_argument("--poisson", action="store_true", help="blend against the date find_good_close_image picks")
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--skip-threshold", type=float, default=0.5)
    ap.add_argument("--distance-weight", type=float, default=0.5)
    try:
        a = ap.parse_args(list(argv))
    except SystemExit:
        return -1
    bands = [b for b in a.bands.split(",") if b]
    shard = scenes.shard_from_env()
    if shard[1] > 1:
        os.environ.setdefault("SATFILL_DEVICE", os.environ.get("LOCAL_RANK", "0"))
    if a.poisson:
        done = scenes.blend_missing_data_folder(a.base_folder, bands, not a.no_cache, a.skip_threshold, a.distance_weight,
                                                shard=shard)  # fmt: skip
    else:
        done = scenes.fill_missing_data_folder(a.base_folder, bands, not a.no_cache, a.skip_threshold, shard=shard)
    for name, ids in done.items():
        _log.info("%s: %s", name, ", ".join(f"{b} -> id {i}" for b, i in ids.items()))
    return 0


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=logging.INFO, format="%(levelname)s %(name)s: %(message)s")
    if argv and argv[0] == "laplace_main":
        return laplace_main(argv[1:])
    if argv and argv[0] == "poisson_main":
        return poisson_main(argv[1:])
    if argv and argv[0] == "fill_folder":
        return fill_folder_main(argv[1:])
    print(__doc__)
    return -1


if __name__ == "__main__":
    sys.exit(main())
