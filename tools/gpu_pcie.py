"""PCIe facts for the host-pointer entry points: pinned H2D, D2H, and both at once (2D copies like api.cu's)."""
import time, torch
n = 10980
nb = 6
dev = torch.device("cuda", 0)
h_in = [torch.empty((n, n), dtype=torch.float64).pin_memory() for _ in range(nb)]
h_out = [torch.empty((n, n), dtype=torch.float64).pin_memory() for _ in range(nb)]
d = [torch.empty((n, n), dtype=torch.float64, device=dev) for _ in range(nb)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
gb = nb * n * n * 8 / 1e9
def run(do_in, do_out):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if do_in:
        with torch.cuda.stream(s_in):
            for b in range(nb): d[b].copy_(h_in[b], non_blocking=True)
    if do_out:
        with torch.cuda.stream(s_out):
            for b in range(nb): h_out[b].copy_(d[b], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    a = run(True, False); b = run(False, True); c = run(True, True)
    print(f"H2D {gb/a:.1f} GB/s  D2H {gb/b:.1f} GB/s  both: {gb/c:.1f} GB/s each direction ({c*1e3:.0f} ms for {gb:.1f} GB each way)")
