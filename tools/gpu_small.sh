#!/bin/bash
# scene-sized workloads (launch-bound regime): c1 (the reference's own scene shape) and small
tag=${1:-run}; out=gpurun_out; mkdir -p $out
for w in c1 small; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu --no-dropin --no-multi $EXTRA_ARGS > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err; echo "$w rc=$?"
  python - $out/${tag}_bench_$w.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(" ms/step", round(d["ms_per_step"],3), "setup", round(d["config"]["setup_ms"],3), "solve", round(d["config"]["solve_ms"],3), "iters", d["config"]["cg_iterations"][:3], "launches/step", d["gpu_launches"]/d["steps"], "e2e s", d["e2e"]["seconds_per_step"] if d["e2e"] else None)
for k,v in d["roofline"]["all_kernels"].items(): print("   %-45s %8.2f ms %5d  %6.0f GB/s  %.3f" % (k, v["ms"], v["launches"], v["GBps"] or 0, v["frac"] or 0))
PY
done
