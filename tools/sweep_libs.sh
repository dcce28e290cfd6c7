#!/bin/bash
# A/B of kernel variants: runs a short bench per shared object and prints the cell-kernel TF/s
for lib in "$@"; do
  out=$(DFTFE_B200_LIB=$PWD/$lib python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-scf --lanes 0 --nwfc 512 2>&1 | tail -1)
  echo "$lib $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("cellTF", round(d["roofline"]["achieved"],2), "ms/launch", round(d["roofline"]["avg_launch_ms"],4), "filterTF", round(d["tflops_fp64_filter"],2))
except Exception as e: print("ERR", e)')" $(echo "$out" | grep -oE 'DftfeB200Error.*|Error.*' | head -1 | cut -c1-160)
done
