#!/bin/bash
# ncu evidence for the fused cell kernel (B200_PROFILING.md recipe). Run under gpurun.
set -x
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --nwfc 256"
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cell_matvec -s 200 -c 2 -f -o gpurun_out/prof_cell python bench.py $ARGS > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
