#!/bin/bash
# e2e (pinned, direct mode) against the size of the background scatter grid
out=gpurun_out; mkdir -p $out
for sc in 16 48 148 296; do
  SATFILL_SCATTER_CTAS=$sc timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu --no-dropin --no-multi > $out/r2w_e2e_$sc.json 2> $out/r2w_e2e_$sc.err
  python - $out/r2w_e2e_$sc.json $sc <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("scatter ctas", sys.argv[2], "e2e s/step", round(d["e2e"]["seconds_per_step"],4), "device ms", round(d["ms_per_step"],1))
PY
done
