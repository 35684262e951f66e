export BENCH_ARGS="--no-e2e"
STEPS=3 bash tools/gpu_multi.sh r2q_win $1 2>&1 | grep -v "^    {'case'\|Setting OMP\|^\*\*\*\*\|^$\|per_rank\|one_tile"
export PYTEST=0
SATFILL_DIST_NO_WINDOW=1 STEPS=3 bash tools/gpu_multi.sh r2q_nowin $1 2>&1 | grep -E "row_decomposed|dist_parity|bench rc"
