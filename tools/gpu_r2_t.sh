#!/bin/bash
# round-2 GPU check T (1 GPU): first-order response / onlyHPrime / complex mixed tests, config 1 and config 4 bench lines
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "first_order or mixed or spectrum or nonlocal or multirank" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r2t_config1_1gpu.json 2> gpurun_out/r2t_config1_1gpu.err; echo "config1 rc=$?"
tail -c 1500 gpurun_out/r2t_config1_1gpu.json; tail -3 gpurun_out/r2t_config1_1gpu.err
timeout 1200 python bench.py --config 4 --steps 3 --warmup 3 --no-scf > gpurun_out/r2t_config4_1gpu.json 2> gpurun_out/r2t_config4_1gpu.err; echo "config4 rc=$?"
tail -c 1500 gpurun_out/r2t_config4_1gpu.json; tail -3 gpurun_out/r2t_config4_1gpu.err
# host topology, for the e2e discussion in DESIGN.md
{ ls /sys/devices/system/node | tr '\n' ' '; echo; for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status; nproc; nvidia-smi topo -m; } > gpurun_out/r2t_topology.txt 2>&1
tail -30 gpurun_out/r2t_topology.txt
