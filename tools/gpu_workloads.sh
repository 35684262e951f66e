#!/bin/bash
# the other configurations of BASELINE.json and the variants DESIGN.md quotes, one short bench line each (1 GPU)
tag=${1:-run}; out=gpurun_out; mkdir -p $out
run() { name=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-dropin --no-multi "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; echo "$name rc=$?";
  python - $out/${tag}_bench_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("   ms/step %.2f value %.3g px/s iters %s e2e %s s step_frac %.3f dominant %s %.3f" % (d["ms_per_step"], d["value"], d["config"]["cg_iterations"][:1], e.get("seconds_per_step"), d["roofline"]["step_frac"], d["roofline"]["kernel"], d["roofline"]["frac"]))
except Exception as ex: print("   failed", ex)
PY
}
run c3_poisson --workload c3-poisson
run c4 --workload c4
run c3_iid --workload c3-iid
run c3_bicubic48 --workload c3-bicubic48
run c1 --workload c1 --steps 20 --warmup 5
run small --workload small --steps 20 --warmup 5
run c3_jacobi --precond jacobi --steps 1 --warmup 1 --no-e2e
run c5 --workload c5
