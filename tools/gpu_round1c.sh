#!/bin/bash
# GPU pass 2: new tests, HBM-kernel bandwidths, lanes A/B at N=1, ncu of the row kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
python tools/bench_hbm_kernels.py > gpurun_out/hbm_kernels.jsonl 2> gpurun_out/hbm_kernels.err
cat gpurun_out/hbm_kernels.jsonl | cut -c1-200; tail -3 gpurun_out/hbm_kernels.err
for L in 0 1; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-scf --lanes $L > gpurun_out/bench_lanes$L.json 2> gpurun_out/bench_lanes$L.err
  python -c "import json;d=json.load(open('gpurun_out/bench_lanes$L.json'));print('lanes',$L,d['value'],d['ms_per_step'],d['roofline']['achieved'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"distribute_vec|slave_to_master_vec|strided_copy_vec|row_scale_vec|unpack_add_vec|pack_rows_vec" -c 12 -f -o gpurun_out/prof_rows python tools/bench_hbm_kernels.py --reps 1 > gpurun_out/prof_rows.log 2>&1
ls -la gpurun_out
