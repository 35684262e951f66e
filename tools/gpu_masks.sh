#!/bin/bash
# gpurun: throughput of the CG / multigrid kernels against the shape of the unknown set (diagnostic masks of bench.py)
out=gpurun_out; mkdir -p $out
for m in full tilecheck tilecheck64 halfrows halfcols; do
  python bench.py --steps 1 --warmup 1 --bands 4 --no-e2e --no-cpu --mask $m --tol 1e-3 > $out/mask_$m.log 2>&1
  python - $out/mask_$m.log $m <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]["all_kernels"]
        print("mask[%s] unknowns %d ms/step %.1f it %s"%(sys.argv[2], d["config"]["unknowns_per_band"], d["ms_per_step"], d["config"]["cg_iterations"]), {k[:14]:(round(v["ms"]/max(v["launches"],1),2), round(v["GBps"] or 0)) for k,v in r.items() if v["ms"]})
        break
else:
    print("mask[%s] FAILED"%sys.argv[2]); print(open(sys.argv[1]).read()[-1500:])
P
done
