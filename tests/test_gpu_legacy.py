"""GPU: the first-generation kernels -- cg_variant = 1 (cg.cu), MG_JACOBI64 (mg.cu, mg_fused.cu), MG_RB32_CTA (mg_rb.cu) -- as
the tested references of the product kernels.  They are not in the product library (SATFILL_LEGACY_VARIANTS = 0): this test
re-runs the parity tests that exercise them in a child process against lib/libsatfill_legacy.so (SATFILL_LIB), and checks
that the product library refuses them loudly."""
import os
import subprocess
import sys

import numpy as np
import pytest

import satellite_approximation_b200 as sab
from satellite_approximation_b200 import _capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEGACY = os.path.join(os.path.dirname(_capi.LIB_PATH), "libsatfill_legacy.so")


def test_product_library_refuses_the_legacy_variants(ctx):
    if ctx.has_legacy_variants:
        pytest.skip("SATFILL_LIB points at the legacy library")
    img = synth.smooth_band(40, 40, seed=1)
    mask = synth.blob_mask(40, 40, cover=0.3, sigma=4.0, seed=2)
    for opts in (dict(cg_variant=1), dict(mg_variant=sab.MG_JACOBI64), dict(mg_variant=sab.MG_RB32_CTA)):
        work = img.copy()
        with pytest.raises(sab.SatfillError) as e:
            ctx.laplace_fill([work], mask, precond=sab.MULTIGRID, tolerance=1e-8, **opts)
        assert e.value.status == sab.SA_BAD_ARGUMENT and np.array_equal(work, img)


def test_first_generation_kernels_against_the_product_kernels():
    if not os.path.exists(LEGACY):
        pytest.skip("lib/libsatfill_legacy.so is not built (make -C satellite_approximation_b200/csrc legacy)")
    env = dict(os.environ, SATFILL_LIB=LEGACY)
    sel = "mask_changes or fused_multigrid or rb_preconditioner or multigrid_variants or golden or tiny"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)  # fmt: skip
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]  # nothing fell back to "needs legacy"
