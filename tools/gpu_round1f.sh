#!/bin/bash
# 2-GPU pass: NCCL parity, then the bench line at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time python -m pytest tests/test_gpu_nccl.py tests/test_gpu_mixed_split.py -m gpu -x -q ) > gpurun_out/pytest_nccl.log 2>&1
tail -8 gpurun_out/pytest_nccl.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_2gpu.log 2>&1
grep '"metric"' gpurun_out/bench_2gpu.log > gpurun_out/bench_2gpu.json
tail -c 2500 gpurun_out/bench_2gpu.log
