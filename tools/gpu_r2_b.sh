#!/bin/bash
# round-2 GPU check B: p2p probe (watchdog), then the new test files under per-test time limits
mkdir -p gpurun_out
timeout 90 python tools/p2p_probe.py > gpurun_out/r2b_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2b_probe.log
tail -25 gpurun_out/r2b_probe.log
P2P=0 timeout 90 python tools/p2p_probe.py > gpurun_out/r2b_probe0.log 2>&1; echo "probe0 rc=$?" >> gpurun_out/r2b_probe0.log
tail -3 gpurun_out/r2b_probe0.log
timeout 900 python -m pytest tests/test_gpu_reference_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_parity.py tests/test_gpu_mixed_split.py -q --timeout 240 --timeout-method=thread -k "not p2p" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
