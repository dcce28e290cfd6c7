#!/bin/bash
# experiment: does a 12-warp cell CTA (11 MMA warps + producer; 3 warps per register-file partition) let the other
# lane's HBM-bound kernels (non-local, row kernels) run in its shadow?  config 3 on one GPU, base vs variant
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ARGS="--config 3 --steps 2 --warmup 1 --no-e2e --no-scf --no-cpu-baseline --no-parity"
for v in base w11; do
  if [ $v = base ]; then unset DFTFE_B200_LIB; else export DFTFE_B200_LIB=$PWD/dftfe_b200/lib/variants/lib_$v.so; fi
  timeout 600 python bench.py $ARGS > gpurun_out/r2w11_config3_$v.json 2> gpurun_out/r2w11_config3_$v.err; echo "$v rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/r2w11_config3_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("$v", d["ms_per_step"], r["achieved"], r["avg_launch_ms"], r["other_kernels_ms_in_that_step"])
PY
done
export DFTFE_B200_LIB=$PWD/dftfe_b200/lib/variants/lib_w11.so
timeout 300 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread -k "nonlocal or test_chebyshev_filter or pipeline" 2>&1 | tail -3
