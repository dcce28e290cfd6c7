#!/bin/bash
# round-2 GPU check E: the whole GPU suite under per-test time limits
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -30 gpurun_out/r2e_pytest.log
