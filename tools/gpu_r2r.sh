export BENCH_ARGS="--no-e2e"
export PYTEST=0
STEPS=3 bash tools/gpu_multi.sh r2r $1 2>&1 | grep -v "^    {'case'\|Setting OMP\|^\*\*\*\*\|^$"
