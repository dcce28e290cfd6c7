#!/bin/bash
# round-2 GPU check F: new cell kernel (parked accumulators + epilogue warps): parity tests, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py tests/test_gpu_adaptive.py tests/test_gpu_mixed_split.py -q -x --timeout 300 --timeout-method=thread > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -15 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-scf --no-e2e > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -c 2600 gpurun_out/r2f_bench.json | cut -c1-2600; tail -3 gpurun_out/r2f_bench.err
for B in 128 200 100; do timeout 300 python bench.py --steps 2 --warmup 1 --no-scf --no-e2e --no-cpu-baseline --no-parity --block $B --nwfc $((B*8)) > gpurun_out/r2f_bench_B$B.json 2>/dev/null; python -c "
import json,sys
for l in open('gpurun_out/r2f_bench_B$B.json'):
    if l.startswith('{'):
        d=json.loads(l); print('B=$B', d['roofline']['achieved'], d['roofline']['frac'], d['tflops_fp64_filter'])
"; done
