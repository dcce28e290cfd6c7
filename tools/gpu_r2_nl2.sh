#!/bin/bash
# non-local kernels with row-id prefetch and eight rows in flight: parity tests, then the same ncu capture as before
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread -k "nonlocal or adaptive or loopback_multirank_filter or first_order or kpoint" > gpurun_out/r2nl2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2nl2_pytest.log
tail -5 gpurun_out/r2nl2_pytest.log
bash tools/gpu_r2_ncu3.sh 2>&1 | head -8 | cut -c1-260
