export BENCH_ARGS="--no-e2e"
STEPS=3 bash tools/gpu_multi.sh r2p_peer $1 2>&1 | grep -v "^    {'case'\|Setting OMP\|^\*\*\*\*\|^$"
export PYTEST=0
SATFILL_DIST_NCCL_ONLY=1 STEPS=3 bash tools/gpu_multi.sh r2p_nccl $1 2>&1 | grep -E "row_decomposed|dist_parity|bench rc|cg_|mg_"
