import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import satellite_approximation_b200 as sab
from satellite_approximation_b200 import synth
ctx = sab.Context(0)
rows, cols, nb = 700, 900, 3
mask = synth.blob_mask(rows, cols, cover=0.3, sigma=12.0, seed=2)
bands = [synth.smooth_band(rows, cols, seed=100 + b) for b in range(nb)]
sc = ctx.scene(sab.LAPLACE, rows, cols, nb)
for b in range(nb): sc.set_band(b, bands[b])
for it in range(2):
    sc.set_mask(mask)
    st = sc.solve(tolerance=1e-6, precond=sab.MULTIGRID)
    print(it, [s["iterations"] for s in st], flush=True)
sc.set_mask(mask)
st = sc.solve(tolerance=1e-6, precond=sab.JACOBI)
print("jacobi", [s["iterations"] for s in st], flush=True)
g = [synth.second_date(b, seed=i) for i, b in enumerate(bands)]
m2 = synth.blob_mask(rows, cols, cover=0.3, sigma=12.0, seed=2, clear_border=False)
w = [b.copy() for b in bands]
st = ctx.poisson_blend(w, g, m2, tolerance=1e-6, precond=sab.MULTIGRID)
print("poisson", [s["iterations"] for s in st], flush=True)
