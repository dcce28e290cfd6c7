#!/bin/bash
# round-2 ncu evidence for the kernels added this round: single-column cell kernel, atom-parallel non-local kernels
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cell_gemv -s 16 -c 2 -f -o gpurun_out/r02_prof_gemv python tools/run_lanczos_timing.py > gpurun_out/r02_prof_gemv.log 2>&1; echo "gemv capture rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_gemv.ncu-rep gpurun_out/r02_cell_gemv_ncu_summary.csv; cat gpurun_out/r02_cell_gemv_ncu_summary.csv | cut -c1-160
ARGS="--config 3 --cells 13 --atoms 64 --nwfc 400 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-scf --lanes 0"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:nl_ -s 30 -c 6 -f -o gpurun_out/r02_prof_nl python bench.py $ARGS > gpurun_out/r02_prof_nl.log 2>&1; echo "nl capture rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_nl.ncu-rep gpurun_out/r02_nonlocal_ncu_summary.csv; cat gpurun_out/r02_nonlocal_ncu_summary.csv | cut -c1-200
ls -la gpurun_out/*.ncu-rep
