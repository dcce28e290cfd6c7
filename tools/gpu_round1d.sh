#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.jsonl 2> gpurun_out/hbm_kernels.err
python -c "
import json
for l in open('gpurun_out/hbm_kernels.jsonl'):
    d=json.loads(l); print(d['kernel'], round(d['ms'],4), round(d['GBps']), round(d['frac'],3))"
bash tools/sweep_libs.sh dftfe_b200/lib/variants/lib_*.so 2>&1 | tee gpurun_out/sweep.txt
