"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the reference's own arithmetic.

Run here (the container that has /root/reference):   python oracle/make_golden.py

Every output is produced by oracle/_ref/libref_eigen.so, i.e. the reference's assembly (laplace.cpp:31-120,
poisson.cpp:145-290) executed by the reference's vendored Eigen ConjugateGradient, run to convergence
(tolerance = Eigen's default epsilon for Laplace, 1e-14 for Poisson).  The fixtures travel to the GPU box, where
/root/reference does not exist; tests compare both the plain-C port (oracle/satfill_oracle.c) and the CUDA path
against them.

The C1 inputs come from the reference's sample scene test_data/2019-05-22 (B04.tif, B08.tif, selected_pixels.png; mask
rule (R >= 220) & (G <= 150), laplace.cpp:141-146): the full-resolution mask is stored bit-packed, the bands as a
320 x 320 uint16 crop.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from satellite_approximation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SCENE = "/root/reference/test_data/2019-05-22"


def unknown_values(img, mask):
    return img[mask].astype(np.float64)


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    oracle.build(ref=True)
    ref = oracle.ref()
    assert ref is not None, "oracle/_ref/libref_eigen.so could not be built (no /root/reference?)"

    # ---- 1. small synthetic Laplace cases (border-free masks: the only ones the reference solves)
    cases = {}
    for i, (rows, cols, sigma, cover) in enumerate([(48, 40, 3.0, 0.35), (33, 97, 2.0, 0.5), (64, 64, 6.0, 0.3)]):
        img = synth.smooth_band(rows, cols, seed=10 + i)
        mask = synth.blob_mask(rows, cols, cover=cover, sigma=sigma, seed=20 + i)
        out, st = ref.laplace_fill(img, mask, tol=0.0, max_it=0)
        assert st.status == 0, st
        cases[f"lap{i}_img"] = img
        cases[f"lap{i}_mask"] = mask
        cases[f"lap{i}_out"] = np.ascontiguousarray(out)
        cases[f"lap{i}_iters"] = np.int64(st.iterations)
    # ---- 2. small synthetic Poisson cases, masks may touch the image border (Poisson is SPD either way)
    for i, (rows, cols, sigma, cover) in enumerate([(40, 56, 3.0, 0.4), (61, 35, 4.0, 0.3)]):
        f = [synth.smooth_band(rows, cols, seed=30 + 2 * i + b) for b in range(2)]
        g = [synth.second_date(f[(b + 1) % 2], seed=40 + b) for b in range(2)]
        mask = synth.blob_mask(rows, cols, cover=cover, sigma=sigma, seed=50 + i, clear_border=False)
        assert mask.any() and not mask.all()
        outs, st = ref.poisson_blend(f, g, mask, tol=1e-14, max_it=100000)
        assert st[0].status == 0, st
        cases[f"poi{i}_f"] = np.stack(f)
        cases[f"poi{i}_g"] = np.stack(g)
        cases[f"poi{i}_mask"] = mask
        cases[f"poi{i}_out"] = np.stack([np.ascontiguousarray(o) for o in outs])
    np.savez_compressed(os.path.join(OUT, "small_cases.npz"), **cases)

    # ---- 3. the reference's own sample scene (config 1 / config 2)
    import cv2

    png = cv2.imread(os.path.join(SCENE, "selected_pixels.png"), cv2.IMREAD_UNCHANGED)
    red, green = png[..., 2], png[..., 1]
    full_mask = (red >= 220) & (green <= 150)  # laplace.cpp:141-146 with red_threshold 220 (laplace-main.cpp:38)
    assert full_mask.sum() == 633573, full_mask.sum()
    b04 = cv2.imread(os.path.join(SCENE, "B04.tif"), cv2.IMREAD_UNCHANGED)
    b08 = cv2.imread(os.path.join(SCENE, "B08.tif"), cv2.IMREAD_UNCHANGED)
    r0, c0, n = 520, 480, 320
    crop_mask = full_mask[r0 : r0 + n, c0 : c0 + n].copy()
    crop_mask[0, :] = crop_mask[-1, :] = False
    crop_mask[:, 0] = crop_mask[:, -1] = False
    crop04 = b04[r0 : r0 + n, c0 : c0 + n].copy()
    crop08 = b08[r0 : r0 + n, c0 : c0 + n].copy()
    lap, st = ref.laplace_fill(crop04.astype(np.float64), crop_mask, tol=0.0, max_it=0)
    assert st.status == 0, st
    f = [crop04.astype(np.float64), crop08.astype(np.float64)]
    g = [synth.second_date(f[1], seed=0), synth.second_date(f[0], seed=1)]
    poi, pst = ref.poisson_blend(f, g, crop_mask, tol=1e-14, max_it=1000000)
    assert pst[0].status == 0, pst
    np.savez_compressed(
        os.path.join(OUT, "c1_scene.npz"),
        full_mask_bits=np.packbits(full_mask),
        full_shape=np.array(full_mask.shape, np.int64),
        crop_origin=np.array([r0, c0], np.int64),
        crop_mask_bits=np.packbits(crop_mask),
        crop_b04=crop04,
        crop_b08=crop08,
        laplace_unknowns=unknown_values(np.ascontiguousarray(lap), crop_mask),
        laplace_iters=np.int64(st.iterations),
        poisson_unknowns=np.stack([unknown_values(np.ascontiguousarray(o), crop_mask) for o in poi]),
    )
    # ---- 4. the steps either side of the path: apply_laplace on a crop of the sample scene, and
    #         preprocess_cloud_band's morphological close pinned against OpenCV itself (cv2 exists in this container only)
    rng = np.random.default_rng(7)
    pre = {}
    inv = png[r0 : r0 + 256, c0 : c0 + 256, :3].copy()  # B, G, R as cv::imread(IMREAD_COLOR) returns them
    inv[0, :] = inv[-1, :] = 0  # keep the marked region off the crop's border
    inv[:, 0] = inv[:, -1] = 0
    b02 = cv2.imread(os.path.join(SCENE, "B02.tif"), cv2.IMREAD_UNCHANGED)
    b03 = cv2.imread(os.path.join(SCENE, "B03.tif"), cv2.IMREAD_UNCHANGED)
    base = np.stack([np.clip(x[r0 : r0 + 256, c0 : c0 + 256] / 12.0, 0, 255).astype(np.uint8) for x in (b02, b03, b04)], axis=-1)
    want, m = oracle.apply_laplace(base, inv, 220.0, engine=ref, tol=0.0, max_it=0)  # the reference's Eigen, per channel
    pre["al_image"], pre["al_invalid"], pre["al_out"], pre["al_mask"] = base, inv, want, m
    kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (11, 11))  # poisson-main.cpp:13-16, dilation_size 5
    cld = cv2.imread(os.path.join(SCENE, "CLD.tif"), cv2.IMREAD_UNCHANGED).astype(np.float64)
    bands = [cld[400:700, 300:650].copy(), (rng.random((97, 130)) < 0.08).astype(np.float64) * rng.integers(1, 100, (97, 130)),
             (rng.random((40, 33)) < 0.5).astype(np.float64), np.zeros((20, 20)), rng.standard_normal((64, 5)) * (rng.random((64, 5)) < 0.3)]
    for i, bnd in enumerate(bands):
        closed = cv2.morphologyEx(bnd, cv2.MORPH_CLOSE, kernel)
        pre[f"mc{i}_band"] = bnd
        pre[f"mc{i}_mask"] = closed != 0
        assert np.array_equal(oracle.morph_close_mask(bnd, 5), closed != 0), i  # the restatement against OpenCV
    np.savez_compressed(os.path.join(OUT, "prepost_cases.npz"), **pre)
    print("apply_laplace crop: masked", int(m.sum()), "morph cases", len(bands))
    print("crop unknowns", int(crop_mask.sum()), "laplace iters", st.iterations, "poisson iters", pst[0].iterations)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
