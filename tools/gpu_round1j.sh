#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for lib in base cs5 cs10; do
  DFTFE_B200_LIB=$PWD/dftfe_b200/lib/variants/lib_$lib.so timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-scf --lanes 0 --nwfc 512 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$lib', 'cellTF', round(d['roofline']['achieved'],2), 'ms/launch', round(d['roofline']['avg_launch_ms'],4))"
done
