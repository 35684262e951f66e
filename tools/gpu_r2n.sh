for t in 1 2 3; do echo "== tail ctas/SM $t"; SATFILL_TAIL_CTAS_PER_SM=$t bash tools/gpu_small.sh r2n_t$t 2>&1 | grep -E "ms/step|tail"; done
echo "== deep tiles 0 (grid barriers only), 3 ctas"; SATFILL_TAIL_CTAS_PER_SM=3 SATFILL_TAIL_DEEP_TILES=0 bash tools/gpu_small.sh r2n_d0 2>&1 | grep -E "ms/step|tail"
echo "== tail from items<=2048"; SATFILL_TAIL_CTAS_PER_SM=3 SATFILL_TAIL_ITEMS=2048 bash tools/gpu_small.sh r2n_i2k 2>&1 | grep -E "ms/step|tail|coarse"
