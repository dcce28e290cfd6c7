#!/bin/bash
# round-2 GPU check P (1 GPU): the full -m gpu suite on the final code, then the default bench line (both arms)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -12 gpurun_out/r2p_pytest.log
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2p_bench.json; tail -3 gpurun_out/r2p_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; echo "ref rc=$?"
tail -c 1200 gpurun_out/r2p_bench_reference.json
