#!/bin/bash
# 8-GPU sanity of the bench line the driver will run (strict timeout)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L | wc -l; free -g | head -2
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_8gpu.log 2>&1
grep '"metric"' gpurun_out/bench_8gpu.log > gpurun_out/bench_8gpu.json
tail -c 1800 gpurun_out/bench_8gpu.log
