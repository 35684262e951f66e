#!/bin/bash
# gpurun: GPU tests on the product library, then the bench line of every tuning variant given
#   gpurun -- 'bash tools/gpu_variants.sh tag "" _u4 _u6'
tag=$1; shift
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "$@"; do
  lib=$PWD/satellite_approximation_b200/lib/libsatfill$v.so
  SATFILL_LIB=$lib python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $out/${tag}_b$v.log 2>&1
  python - $out/${tag}_b$v.log "$v" <<'P'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]["all_kernels"]
        print("variant[%s] ms/step %.1f it %s"%(sys.argv[2], d["ms_per_step"], d["config"]["cg_iterations"]), {k[:14]:(round(v["ms"]/d["steps"],1), round(v["GBps"] or 0)) for k,v in r.items() if v["ms"]})
        break
else:
    print("variant[%s] FAILED"%sys.argv[2]); print(open(sys.argv[1]).read()[-2000:])
P
done
