#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"density_kernel|ham_assemble_kernel" -c 4 -f -o gpurun_out/prof_next python bench.py --nwfc 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_next.log 2>&1
tail -3 gpurun_out/prof_next.log | cut -c1-300
ls -la gpurun_out/prof_next.ncu-rep
