#!/bin/bash
# A/B of run-time switches of libsatfill: one short bench line per environment setting.
#   gpurun -- 'bash tools/gpu_envab.sh r2p "SATFILL_RBW_PIPE=1" "SATFILL_RBW_PIPE=0"'
tag=$1; shift
out=gpurun_out; mkdir -p $out
B="--steps ${STEPS:-4} --warmup 2 --no-e2e --no-cpu --no-dropin --no-multi"
if [ "${PYTEST:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rb_preconditioner or multigrid or tiny or mask_changes or c1_full or bench_tolerance" > $out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
fi
i=0
for v in "$@"; do
  i=$((i+1))
  env $v timeout 300 python bench.py $B $EXTRA_ARGS > $out/${tag}_bench$i.json 2> $out/${tag}_bench$i.err
  echo "== '$v' rc=$?"
  python - "$out/${tag}_bench$i.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms/step", round(d["ms_per_step"],2), "iters", d["config"]["cg_iterations"][:2], "step_frac", round(d["roofline"]["step_frac"],3))
    for k,v in d["roofline"]["all_kernels"].items(): print("   %-45s %8.1f ms %5d  %6.0f GB/s  %.3f" % (k, v["ms"], v["launches"], v["GBps"] or 0, v["frac"] or 0))
except Exception as e: print(" failed", e)
PY
done
