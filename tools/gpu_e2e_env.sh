#!/bin/bash
# e2e (host buffers in, host buffers out through sa_laplace_fill) under different run-time switches of libsatfill
#   gpurun -- 'bash tools/gpu_e2e_env.sh r2q "" "SATFILL_SCATTER_ROWS=8" "SATFILL_SCATTER_ROWS=8 SATFILL_SCATTER_CTAS=32"'
tag=$1; shift
out=gpurun_out; mkdir -p $out
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --steps ${STEPS:-3} --warmup 2 --no-cpu --no-dropin --no-multi $EXTRA_ARGS > $out/${tag}_e2e$i.json 2> $out/${tag}_e2e$i.err
  echo "== '$v' rc=$?"
  python - "$out/${tag}_e2e$i.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.2f   e2e s/step %.4f  (%.2f G px/s)" % (d["ms_per_step"], d["e2e"]["seconds_per_step"], d["e2e"]["value"]*1e-9))
except Exception as e: print(" failed", e)
PY
done
