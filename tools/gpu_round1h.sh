#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cfg in "--nwfc 512" "--nwfc 400 --block 100" "--nwfc 440 --block 110" "--nwfc 480 --block 120"; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-scf --lanes 0 $cfg 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$cfg', 'cellTF', round(d['roofline']['achieved'],2), 'filterTF', round(d['tflops_fp64_filter'],2), 'ms', round(d['ms_per_step'],1))"
done
