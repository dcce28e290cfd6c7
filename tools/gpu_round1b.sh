#!/bin/bash
# GPU pass: parity tests, full bench line, ncu captures of the projection / rotation kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 3000 gpurun_out/bench_1gpu.json
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --nwfc 512 --degree 4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"xty_partial|xq_kernel" -c 4 -f -o gpurun_out/prof_proj python bench.py $ARGS > gpurun_out/prof_proj.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_solve.csv python bench.py $ARGS > gpurun_out/launches_solve.log 2>&1
ls -la gpurun_out
