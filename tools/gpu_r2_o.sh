#!/bin/bash
# round-2 GPU check O (1 GPU): non-local tests, then bench lines of configs 3, 5 and 1
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "nonlocal or adaptive_multirank or complex_nonlocal" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -4 gpurun_out/r2o_pytest.log
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-e2e --no-scf --no-cpu-baseline --no-parity > gpurun_out/r2o_config3_1gpu.json 2> gpurun_out/r2o_config3_1gpu.err; echo "config3 rc=$?"
tail -c 700 gpurun_out/r2o_config3_1gpu.json; tail -2 gpurun_out/r2o_config3_1gpu.err
timeout 900 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2o_config5_1gpu.json 2> gpurun_out/r2o_config5_1gpu.err; echo "config5 rc=$?"
tail -c 2500 gpurun_out/r2o_config5_1gpu.json; tail -3 gpurun_out/r2o_config5_1gpu.err
timeout 900 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r2o_config1_1gpu.json 2> gpurun_out/r2o_config1_1gpu.err; echo "config1 rc=$?"
tail -c 2500 gpurun_out/r2o_config1_1gpu.json; tail -3 gpurun_out/r2o_config1_1gpu.err
