import sys, os, torch
sys.path.insert(0, os.getcwd())
import satellite_approximation_b200 as sab
from satellite_approximation_b200 import synth
ctx = sab.Context(0, stream=torch.cuda.current_stream().cuda_stream)
rows = cols = 10980; nb = 13
dev = torch.device("cuda", 0)
mask = synth.torch_blob_mask(rows, cols, cover=0.3, cell=48, seed=2, device=dev)
bands = [synth.torch_band(rows, cols, seed=100 + b, device=dev) for b in range(nb)]
print("band sums", [float(b.sum()) for b in bands])
torch.cuda.synchronize()
sc = ctx.scene(sab.LAPLACE, rows, cols, nb)
for b in range(nb): sc.set_band(b, bands[b])
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    sc.set_mask(mask)
    st = sc.solve(tolerance=1e-6, precond=sab.MULTIGRID, profile=True)
    print(it, [s["iterations"] for s in st], flush=True)
out = torch.empty_like(bands[0])
for b in (0, 8, 12):
    sc.get_band(b, out)
    m = mask.bool()
    lap = 4 * out[1:-1, 1:-1] - out[:-2, 1:-1] - out[2:, 1:-1] - out[1:-1, :-2] - out[1:-1, 2:]
    res = lap[m[1:-1, 1:-1]]
    print("band", b, "max |5-point residual| at unknowns", float(res.abs().max()), "known unchanged", bool((out[~m] == bands[b][~m]).all()))
