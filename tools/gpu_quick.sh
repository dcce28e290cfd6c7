#!/bin/bash
# quick GPU check of a cell-kernel change: parity tests + three short bench configurations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "--nwfc 512 --lanes 0" "--nwfc 400 --block 100 --lanes 0" "--nwfc 1024"; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-scf $cfg 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$cfg', 'cellTF', round(d['roofline']['achieved'],2), 'filterTF', round(d['tflops_fp64_filter'],2), 'ms', round(d['ms_per_step'],1))"
done
