#!/bin/bash
# One gpurun call (1 GPU): the whole GPU test suite, smoke(), then the default bench (both arms).
tag=${1:-run}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -3 $out/${tag}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "reference rc=$?"
python - $out/${tag}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],2), "value", d["value"], "iters", d["config"]["cg_iterations"][:3])
r=d["roofline"]; print("dominant", r["kernel"], round(r["frac"],3), "step_frac", round(r["step_frac"],3))
for k,v in r["all_kernels"].items(): print("   %-45s %8.1f ms %5d  %6.0f GB/s  %.3f" % (k, v["ms"], v["launches"], v["GBps"] or 0, v["frac"] or 0))
print("e2e", json.dumps(d["e2e"])[:600]); print("dropin", json.dumps(d["e2e_dropin"])[:900]); print("cpu", json.dumps(d["cpu_baseline"])[:900])
print("row_decomposed", json.dumps(d["config"].get("row_decomposed"))[:1200])
PY
