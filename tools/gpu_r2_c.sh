#!/bin/bash
# round-2 GPU check C (2 GPUs): NCCL-mode p2p transport parity + 2-GPU bench lines (p2p and nccl transports)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/nccl_parity_main.py > gpurun_out/r2c_nccl_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c_nccl_parity.log
grep "NCCL_PARITY\|rc=" gpurun_out/r2c_nccl_parity.log | cut -c1-1500
DFTFE_B200_TRANSPORT=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/nccl_parity_main.py > gpurun_out/r2c_nccl_parity_nccl.log 2>&1; echo "parity(nccl transport) rc=$?" >> gpurun_out/r2c_nccl_parity_nccl.log
grep "NCCL_PARITY\|rc=" gpurun_out/r2c_nccl_parity_nccl.log | cut -c1-1500
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-scf > gpurun_out/r2c_bench2_p2p.json 2> gpurun_out/r2c_bench2_p2p.err; echo "bench p2p rc=$?"
tail -c 1800 gpurun_out/r2c_bench2_p2p.json; tail -3 gpurun_out/r2c_bench2_p2p.err
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-scf --no-e2e --no-parity --transport nccl > gpurun_out/r2c_bench2_nccl.json 2> gpurun_out/r2c_bench2_nccl.err; echo "bench nccl rc=$?"
tail -c 600 gpurun_out/r2c_bench2_nccl.json
timeout 200 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_ref_arm.json 2>&1; echo "ref arm rc=$?"; tail -c 400 gpurun_out/r2c_ref_arm.json
