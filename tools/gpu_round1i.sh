#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for lib in base pf; do
for cfg in "--nwfc 512 --block 128" "--nwfc 400 --block 100"; do
  DFTFE_B200_LIB=$PWD/dftfe_b200/lib/variants/lib_$lib.so timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-scf --lanes 0 $cfg 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib $cfg', 'cellTF', round(d['roofline']['achieved'],2), 'filterTF', round(d['tflops_fp64_filter'],2), 'ms', round(d['ms_per_step'],1))"
done; done
