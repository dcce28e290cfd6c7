#!/bin/bash
# ncu capture of the atom-parallel non-local kernels (filter phase of a reduced config 3: 13^3 cells, 64 atoms, B = 200)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ARGS="--config 3 --cells 13 --atoms 64 --nwfc 400 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-scf --lanes 0"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:nl_.*_vec_kernel -s 12 -c 8 -f -o gpurun_out/r02_prof_nl python bench.py $ARGS > gpurun_out/r02_prof_nl.log 2>&1; echo "nl capture rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_nl.ncu-rep gpurun_out/r02_nonlocal_ncu_summary.csv; cat gpurun_out/r02_nonlocal_ncu_summary.csv | cut -c1-260
