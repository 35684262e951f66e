#!/bin/bash
# Lean capture (1 GPU): `ncu --set full` of the level-0 cycle kernels (second V-cycle of a solve) on the bench scene, 4 bands.
#   gpurun --timeout 1200 -- 'bash tools/gpu_cycle_capture.sh r2d [mask]'
#   mask: "" = the bench's cloud mask, "full" = a dense hole of the same size (every tile full)
tag=${1:-run}; mask=${2:-}
out=gpurun_out
mkdir -p $out
M=""; [ -n "$mask" ] && M="--mask $mask"
B="--steps 1 --warmup 0 --bands 4 --no-e2e --no-cpu --no-dropin --no-multi $M"
python bench.py --steps 1 --warmup 1 --bands 4 --no-e2e --no-cpu --no-dropin --no-multi $M > $out/${tag}_plain.log 2>&1
echo "plain rc=$?"
# level-0 kernels by their template arguments (second launch = second V-cycle of the solve)
for ks in "k_rbw_down<.int.0>:down0" "k_rbw_up<.int.0, .bool.1>:up0"; do
    k=${ks%%:*}; name=${ks##*:}
    timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"${k}" --launch-skip 1 --launch-count 1 \
        -f -o $out/${tag}_full_${name} python bench.py $B > $out/${tag}_ncu_${name}.log 2>&1
    echo "ncu $name rc=$?"
done
ls -la $out | grep ${tag}_ | tail -12
