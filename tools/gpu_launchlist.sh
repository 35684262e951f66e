#!/bin/bash
# ncu launch list (gpu__time_duration only) of one timed step of a workload: per-kernel shares of the step
tag=${1:-run}; shift; out=gpurun_out; mkdir -p $out
A="--steps 1 --warmup 1 --no-e2e --no-cpu --no-dropin --no-multi $@"
python bench.py $A > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches.csv python bench.py $A > $out/${tag}_ncu.log 2>&1
echo "launch list rc=$?"
python profiles/summarize.py launches $out/${tag}_launches.csv | head -40
