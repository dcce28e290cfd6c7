#!/bin/bash
# round-2 GPU check V (1 GPU): row kernels capped at 48 registers (co-residency with the cell CTAs): tests, step time, Lanczos timing
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "constraints or row_kernels or reference_kernels or multirank or adaptive or lanes" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -5 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-scf > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/r2v_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["roofline"]["other_kernels_ms_in_that_step"])
PY
timeout 300 python tools/run_lanczos_timing.py > gpurun_out/r02_lanczos_timing.json 2> gpurun_out/r02_lanczos_timing.err; echo "lanczos rc=$?"; cat gpurun_out/r02_lanczos_timing.json; tail -3 gpurun_out/r02_lanczos_timing.err
