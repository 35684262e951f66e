"""Worker of tests/test_gpu_dist.py: one rank of the multi-GPU checks (launched by torch.distributed.run, one process per
GPU).  The checks themselves live in satellite_approximation_b200/distcheck.py (bench.py --gpus N runs the same ones in
front of the driver); here the CPU oracle is added on top of them."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import satellite_approximation_b200 as sab  # noqa: E402
from satellite_approximation_b200 import distcheck  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sab.Context(local)
    ctx.dist_init_torch()
    port = oracle.port()

    def vs_oracle(name, problem, mask, bands, guides, got):
        if mask.size >= 400000:
            return {}
        if problem == sab.POISSON:
            want, _ = port.poisson_blend(bands, guides, mask, tol=1e-13, max_it=10**6)
        else:
            want = [port.laplace_fill(b_, mask, mode=1, tol=1e-13)[0] for b_ in bands]
        worst = max(distcheck.rel(got[b], want[b], mask) for b in range(len(bands)))
        return {"max_rel_vs_oracle": worst, "ok": worst < 1e-7}

    res = distcheck.run_all(ctx, world, rank, extra=vs_oracle)
    if rank == 0:
        for r in res["cases"]:
            print(("ok " if r["ok"] else "FAILED ") + r["case"] + ": " + str({k: v for k, v in r.items() if k not in ("case", "ok")}), flush=True)
    assert res["ok"], res
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
