#!/bin/bash
# One GPU pass: parity tests, HBM-kernel bandwidths, the 1-GPU bench line.  Run under gpurun:
#   gpurun --timeout 1200 -- 'bash tools/gpu_validate.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.jsonl 2> gpurun_out/hbm_kernels.err
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 1200 gpurun_out/bench_1gpu.json
