#!/bin/bash
# non-local kernels, final check: config 3 on one GPU with the shipped kernels (4 rows in flight per warp) and with
# the 8-row variant (tools: -DDB_NLV_UNROLL=8), plus the non-local parity tests under the variant
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ARGS="--config 3 --steps 2 --warmup 1 --no-e2e --no-scf --no-cpu-baseline --no-parity"
for v in base nlu8; do
  if [ $v = base ]; then unset DFTFE_B200_LIB; else export DFTFE_B200_LIB=$PWD/dftfe_b200/lib/variants/lib_$v.so; fi
  timeout 400 python bench.py $ARGS > gpurun_out/r2nl4_config3_$v.json 2> gpurun_out/r2nl4_config3_$v.err; echo "$v rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/r2nl4_config3_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("$v", d["ms_per_step"], r["achieved"], r["other_kernels_ms_in_that_step"])
PY
done
timeout 200 python -m pytest tests -m gpu -q --timeout 150 --timeout-method=thread -k "nonlocal or loopback_multirank_filter" 2>&1 | tail -2
