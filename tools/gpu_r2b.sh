#!/bin/bash
# round 2, call b: correctness of the warp-per-tile cycle kernels, then A/B bench lines against the first generation
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rb_preconditioner or multigrid or tiny or mask_changes" > $out/r2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 $out/r2c_pytest.log
B="--steps 5 --warmup 2 --no-e2e --no-cpu --no-dropin --no-multi"
timeout 300 python bench.py $B > $out/r2c_bench_c3_rbw.json 2> $out/r2c_bench_c3_rbw.err; echo "c3 rbw rc=$?"
timeout 300 python bench.py $B --mg-variant rb32_cta > $out/r2c_bench_c3_cta.json 2> $out/r2c_bench_c3_cta.err; echo "c3 cta rc=$?"


for f in $out/r2c_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms/step", round(d["ms_per_step"],2), "iters", d["config"]["cg_iterations"], "step_frac", round(d["roofline"]["step_frac"],3))
    for k,v in d["roofline"]["all_kernels"].items(): print("   ", k, round(v["ms"],1), v["launches"], round(v["GBps"] or 0), round(v["frac"] or 0,3))
except Exception as e: print(" failed", e)
PY
done
tail -5 $out/r2c_bench_c3_rbw.err
