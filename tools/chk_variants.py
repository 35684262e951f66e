import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import satellite_approximation_b200 as sab
from satellite_approximation_b200 import synth
ctx = sab.Context(0)
rows = cols = 4096; nb = 4
dev = torch.device("cuda", 0)
mask = synth.torch_blob_mask(rows, cols, cover=0.3, cell=48, seed=2, device=dev)
bands = [synth.torch_band(rows, cols, seed=100 + b, device=dev) for b in range(nb)]
torch.cuda.synchronize()
sc = ctx.scene(sab.LAPLACE, rows, cols, nb)
for b in range(nb): sc.set_band(b, bands[b])
sc.set_mask(mask)
st = sc.solve(tolerance=1e-6, precond=sab.MULTIGRID, profile=True)
print(os.environ.get("SATFILL_LIB","default")[-20:], [s["iterations"] for s in st], [f'{s["error"]:.2e}' for s in st], st[0]["solve_ms"], [round(x,2) for x in st[0]["kernel_ms"]], st[0]["kernel_units"])
