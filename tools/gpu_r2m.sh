PYTEST=1 bash tools/gpu_variants.sh r2m ""
bash tools/gpu_small.sh r2m
