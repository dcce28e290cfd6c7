cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cell_matvec_persistent -s 40 -c 1 -f -o gpurun_out/prof_cell_src python bench.py --nwfc 256 --steps 1 --warmup 1 --no-scf --no-e2e --no-cpu-baseline --lanes 0 > gpurun_out/prof_cell_src.log 2>&1
ls -la gpurun_out/prof_cell_src.ncu-rep
