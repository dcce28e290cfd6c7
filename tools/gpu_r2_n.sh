#!/bin/bash
# round-2 GPU check N: atom-parallel non-local kernels + segment-major tf32 kernel: tests, mixed timing, config 3 on 1 GPU
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "nonlocal or mixed or adaptive or multirank or solve or spectrum or band" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -8 gpurun_out/r2n_pytest.log
timeout 300 python tools/run_mixed_projection.py > gpurun_out/r02_mixed_projection.json 2> gpurun_out/r02_mixed_projection.err; echo "mixed rc=$?"; tail -c 1500 gpurun_out/r02_mixed_projection.json; tail -3 gpurun_out/r02_mixed_projection.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-e2e --no-scf --no-cpu-baseline --no-parity > gpurun_out/r2n_config3_1gpu.json 2> gpurun_out/r2n_config3_1gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2n_config3_1gpu.json; tail -3 gpurun_out/r2n_config3_1gpu.err
