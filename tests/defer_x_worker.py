"""Child of tests/test_gpu_parity.py::test_deferred_x_update_equals_the_plain_update: runs a list of fills with the
environment it was given (SATFILL_DEFER_X) and stores the filled bands."""
import sys

import numpy as np

import satellite_approximation_b200 as sab
from satellite_approximation_b200 import synth


def cases():
    rows, cols = 300, 420
    mask = synth.blob_mask(rows, cols, cover=0.4, sigma=7.0, seed=5)
    # bands of very different smoothness: with one tolerance they stop after different numbers of passes
    bands = [synth.smooth_band(rows, cols, seed=3), synth.smooth_band(rows, cols, seed=4) * 1e-3 + 7.0,
             np.random.default_rng(8).normal(size=(rows, cols))]
    guides = [synth.second_date(b, seed=6) for b in bands]
    out = []
    for it in (1, 2, 3, 4):  # stopped by the iteration limit: last pass of either parity
        out.append(("laplace-maxit%d" % it, sab.LAPLACE, mask, bands, None, dict(tolerance=1e-30, max_iterations=it)))
        out.append(("poisson-maxit%d" % it, sab.POISSON, mask, bands, guides, dict(tolerance=1e-30, max_iterations=it)))
    for tol in (1e-2, 1e-3, 1e-4, 1e-5):  # stopped by the tolerance, band by band
        out.append(("laplace-tol%g" % tol, sab.LAPLACE, mask, bands, None, dict(tolerance=tol)))
    return out


def run(path):
    ctx = sab.Context(0)
    res = {}
    for name, problem, mask, bands, guides, opts in cases():
        sc = ctx.scene(problem, mask.shape[0], mask.shape[1], len(bands))  # resident scene: the iterate as the solver left it
        sc.set_mask(mask)
        for b, img in enumerate(bands):
            sc.set_band(b, img)
            if problem == sab.POISSON:
                sc.set_guidance(b, guides[b])
        st = sc.solve(precond=sab.MULTIGRID, **opts)
        for b in range(len(bands)):
            res["%s/%d" % (name, b)] = sc.get_band(b)
        res["%s/iters" % name] = np.array([s["iterations"] for s in st])
        sc.close()
    np.savez(path, **res)
    ctx.close()


if __name__ == "__main__":
    run(sys.argv[1])
