#!/bin/bash
# the whole -m gpu suite on the final commit
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 165 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -5 gpurun_out/r2final_pytest.log
