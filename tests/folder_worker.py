"""Worker of tests/test_scenes.py::test_two_ranks_share_one_database: one rank of a torchrun-style launch of the folder
driver (RANK / WORLD_SIZE in the environment), with the oracle as the pixel step so that it runs on a CPU-only machine."""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    import oracle
    from satellite_approximation_b200 import scenes as sc

    base, out = sys.argv[1], sys.argv[2]
    port = oracle.port()

    def fill(bands, mask):
        for b in bands:
            b[...] = port.laplace_fill(b, mask, mode=1)[0]

    done = sc.fill_missing_data_folder(base, ["B04", "B08"], True, 1.0, fill=fill, shard=sc.shard_from_env())
    with open(out, "w") as f:
        json.dump(done, f)


if __name__ == "__main__":
    main()
