#!/bin/bash
# round-2 multi-GPU check: usage gpu_r2_multi.sh <N> [also-config2]
# strong scaling point of config 3 on N GPUs (live-communicator parity included), optionally the default weak line
N=$1
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $TR bench.py --gpus $N --config 3 --steps 3 --warmup 1 --no-e2e --no-scf --no-cpu-baseline > gpurun_out/r2q_config3_${N}gpu.json 2> gpurun_out/r2q_config3_${N}gpu.err; echo "config3 N=$N rc=$?"
tail -c 1800 gpurun_out/r2q_config3_${N}gpu.json; tail -3 gpurun_out/r2q_config3_${N}gpu.err
if [ -n "$2" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-scf --no-cpu-baseline > gpurun_out/r2q_config2_${N}gpu.json 2> gpurun_out/r2q_config2_${N}gpu.err; echo "config2 N=$N rc=$?"
  tail -c 1800 gpurun_out/r2q_config2_${N}gpu.json; tail -3 gpurun_out/r2q_config2_${N}gpu.err
fi
