PYTEST=1 bash tools/gpu_variants.sh r2g "" _c20 _c24
export PYTEST=0
SATFILL_RBW_BAND_MAJOR_MB=-1 bash tools/gpu_variants.sh r2g_tilemajor "" _c20
SATFILL_RBW_BAND_MAJOR_MB=0 bash tools/gpu_variants.sh r2g_allband ""
