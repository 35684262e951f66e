#!/bin/bash
# gpurun --gpus N -- 'bash tools/gpu_multi.sh TAG N': the 2-GPU parity test (when N >= 2) and the N-rank bench line the driver runs
tag=${1:-run}; n=${2:-2}; out=gpurun_out; mkdir -p $out
if [ "${PYTEST:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -x -q > $out/${tag}_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 $out/${tag}_pytest_dist.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-3} --warmup 2 $BENCH_ARGS > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
echo "bench rc=$?"; tail -5 $out/${tag}_bench_${n}gpu.err | cut -c1-400
python - $out/${tag}_bench_${n}gpu.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"],2), "value", d["value"], "n_gpus", d["n_gpus"])
print("per_rank", json.dumps(d["config"]["per_rank"])[:800])
print("e2e", json.dumps(d["e2e"])[:500])
print("one_tile_strong", json.dumps(d["config"].get("one_tile_strong"))[:500])
rd=d["config"].get("row_decomposed") or {}
print("row_decomposed", {k:v for k,v in rd.items() if k!="kernels_rank0"})
for k,v in (rd.get("kernels_rank0") or {}).items(): print("   %-45s %8.1f ms %5d  %6.0f GB/s  %.3f" % (k, v["ms"], v["launches"], v["GBps"] or 0, v["frac"] or 0))
p=d.get("dist_parity") or {}
print("dist_parity ok", p.get("ok"), "max_rel", p.get("max_rel"))
for c in p.get("cases", []): print("   ", c)
PY
