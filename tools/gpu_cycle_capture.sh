#!/bin/bash
# Lean capture (1 GPU): `ncu --set full` of the two level-0 cycle kernels on the cloud-like bench mask and on a dense hole of
# the same size -- the pair that separates "frame halos miss in L2" from "short DRAM runs" (DESIGN.md section 10, item 1).
#   gpurun --timeout 1200 -- 'bash tools/gpu_cycle_capture.sh r2a'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python bench.py --steps 1 --warmup 1 --bands 4 --no-e2e --no-cpu > $out/${tag}_plain.log 2>&1
echo "plain rc=$?"
for ks in k_rb_down:11 k_rb_up:21; do
    k=${ks%%:*}; skip=${ks##*:}
    for m in blobs full; do
        timeout 500 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip --launch-count 1 \
            -f -o $out/${tag}_${m}_full_${k} python bench.py --steps 1 --warmup 0 --bands 4 --mask $m --no-e2e --no-cpu > $out/${tag}_${m}_ncu_${k}.log 2>&1
        echo "ncu $m $k rc=$?"
    done
done
ls -la $out | tail -12
