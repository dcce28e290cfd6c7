#!/bin/bash
# round-2 GPU check Z (2 GPUs): the whole -m gpu suite on the final code on a 2-GPU box (the real-NCCL test runs too),
# then config 3 on 2 GPUs with useMixedPrecCheby (FP32 ghost payloads)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -6 gpurun_out/r2z_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus 2 --config 3 --mixed --steps 3 --warmup 1 --no-e2e --no-scf --no-cpu-baseline > gpurun_out/r2z_config3_2gpu_mixed.json 2> gpurun_out/r2z_config3_2gpu_mixed.err; echo "config3 mixed rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2z_config3_2gpu_mixed.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], d["ms_per_step"], d["config"]["mixed_prec_cheby"], d["parity_multi_gpu"]["ok"], d["roofline"]["other_kernels_ms_in_that_step"])
PY
tail -3 gpurun_out/r2z_config3_2gpu_mixed.err
