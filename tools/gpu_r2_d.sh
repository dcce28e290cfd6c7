#!/bin/bash
# round-2 GPU check D: BASELINE config 3 (Al-FCC-like supercell, FIXED global mesh = strong scaling) on $NG GPUs,
# then (1 GPU only) the new assembly / density-gradient tests
NG=${NG:-1}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ARGS="--config 3 --gpus $NG --steps 3 --warmup 1 --no-e2e --no-scf --no-cpu-baseline --no-parity"
if [ "$NG" = "1" ]; then
  timeout 900 python bench.py $ARGS > gpurun_out/r2d_config3_${NG}gpu.json 2> gpurun_out/r2d_config3_${NG}gpu.err; echo "bench rc=$?"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 bench.py $ARGS > gpurun_out/r2d_config3_${NG}gpu.json 2> gpurun_out/r2d_config3_${NG}gpu.err; echo "bench rc=$?"
fi
tail -c 2500 gpurun_out/r2d_config3_${NG}gpu.json; tail -5 gpurun_out/r2d_config3_${NG}gpu.err
if [ "$NG" = "1" ]; then
  timeout 600 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_reference_kernels.py -q --timeout 240 --timeout-method=thread -k "assembly or density or constraint" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
  tail -30 gpurun_out/r2d_pytest.log
fi
