#!/bin/bash
# One gpurun call (1 GPU): the default bench (both arms), the ncu launch list of the same workload and one `ncu --set full`
# capture per hot kernel (4 bands: ncu saves and restores device memory around each of its ~40 replays).  Everything lands in
# gpurun_out/<tag>_*; tools/collect_profiles.sh turns the ncu outputs into the text files committed under profiles/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh r2z'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python bench.py --steps ${STEPS:-20} --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "bench reference rc=$?"
# launch list: the default workload, one warm-up + one timed step (shares must agree with the bench's own event times)
A="--steps 1 --warmup 1 --no-e2e --no-cpu --no-dropin --no-multi"
python bench.py $A > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $A > $out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# full captures of the level-0 kernels of the second CG iteration (template arguments select the level-0 instantiations)
B="--steps 1 --warmup 0 --bands 4 --no-e2e --no-cpu --no-dropin --no-multi"
# (k_update2: launch 1 = the second pass, XM = 2, x takes two steps; launch 2 = the third pass, XM = 1, x left alone)
for ks in "k_update2<|update2|1" "k_update2<|update2_x_left_alone|2" "k_direction2<|direction2|1" "k_rbw_down<.int.0>|rbw_down_L0|1" "k_rbw_up<.int.0, .bool.1>|rbw_up_L0|1" "k_rbw_down<.int.1>|rbw_down_L1|1" "k_rbw_up<.int.1, .bool.0>|rbw_up_L1|1"; do
    IFS='|' read -r k name skip <<< "$ks"
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"${k}" --launch-skip $skip --launch-count 1 \
        -f -o $out/${tag}_full_${name} python bench.py $B > $out/${tag}_ncu_${name}.log 2>&1
    echo "ncu $name rc=$?"
done
grep -o '"unknowns_per_band": [0-9]*' $out/${tag}_ncu_update2.log | head -1 > $out/${tag}_units.txt
ls -la $out | grep ${tag}_ | tail -20
