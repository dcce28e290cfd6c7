#!/bin/bash
# round-2 GPU check W: usage gpu_r2_w.sh <N>: config 2 (weak) on N GPUs, device-resident value only
N=$1
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-scf --no-cpu-baseline --no-e2e > gpurun_out/r2w_config2_${N}gpu.json 2> gpurun_out/r2w_config2_${N}gpu.err; echo "config2 N=$N rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r2w_config2_${N}gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["ms_per_step"], d["roofline"]["kernel_share_of_step"], d["roofline"]["other_kernels_ms_in_that_step"], d["parity_multi_gpu"]["ok"])
PY
tail -3 gpurun_out/r2w_config2_${N}gpu.err
