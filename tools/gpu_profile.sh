#!/bin/bash
# One gpurun call (1 GPU): GPU parity tests, the default bench (both arms), the ncu launch list of the same command and
# one `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/<tag>_*; profiles/summarize.py turns the
# ncu outputs into the text files committed under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r1b'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.log 2> $out/${tag}_bench.err
echo "bench rc=$?"
python bench.py --impl reference > $out/${tag}_bench_reference.log 2> $out/${tag}_bench_reference.err
echo "bench reference rc=$?"
# launch list: the default workload, one warm-up + one timed step (shares must agree with the bench's own event times)
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full captures, 4 bands (ncu saves and restores device memory around each of its ~40 replays)
# (a V-cycle launches k_rb_down on levels 0..10 and k_rb_up on levels 10..0: skip to a level-0 launch of the 2nd cycle)
for ks in k_update2:2 k_direction2:2 k_rb_down:11 k_rb_up:21; do
    k=${ks%%:*}; skip=${ks##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip --launch-count 1 \
        -f -o $out/${tag}_full_${k} python bench.py --steps 1 --warmup 0 --bands 4 --no-e2e --no-cpu > $out/${tag}_ncu_${k}.log 2>&1
    echo "ncu $k rc=$?"
done
# DENSE=1: the same two cycle kernels on a dense hole of the same size (bench.py --mask full): same halo factor, every DRAM
# granule full -- the pair of captures that separates "halos miss in L2" from "short runs" 
if [ "${DENSE:-0}" = "1" ]; then
    for ks in k_rb_down:11 k_rb_up:21 k_update2:2; do
        k=${ks%%:*}; skip=${ks##*:}
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^${k}\$" --launch-skip $skip --launch-count 1 \
            -f -o $out/${tag}_dense_full_${k} python bench.py --steps 1 --warmup 0 --bands 4 --mask full --no-e2e --no-cpu > $out/${tag}_dense_ncu_${k}.log 2>&1
        echo "ncu dense $k rc=$?"
    done
fi
ls -la $out | tail -20
