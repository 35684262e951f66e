"""GPU, >= 2 devices: one linear system split by rows across ranks (sa_dist_*, dist.cu) gives the same fill as one GPU
and as the oracle.  Launched the way the driver launches bench.py: torch.distributed.run, one process per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_row_decomposed_solve_matches_single_gpu(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert r.stdout.count("ok ") >= 5, r.stdout
