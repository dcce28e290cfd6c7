#!/bin/bash
# round-2 GPU check U (1 GPU): single-column cell kernel (Lanczos applies): parity + solve() pass timing
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "hx_and_hxcheby or lanczos or solve or multirank or adaptive" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -6 gpurun_out/r2u_pytest.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2u_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["scf_iteration"])
PY
tail -3 gpurun_out/r2u_bench.err
