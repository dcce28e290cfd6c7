#!/bin/bash
# round-2 GPU check Y (2 GPUs): reverse push on the copy engines: loopback transport tests, live-transport parity, config 2 weak line
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "multirank or p2p or nccl or lanes or band" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -5 gpurun_out/r2y_pytest.log
bash tools/gpu_r2_w.sh 2
