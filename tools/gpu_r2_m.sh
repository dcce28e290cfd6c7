#!/bin/bash
# round-2 GPU evidence: ncu launch list + full capture of the cell kernel and of the tcgen05 TF32 GEMM, mixed-precision
# projection timings, HBM-bound row kernels at the config-2 size
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-scf --nwfc 512"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_matvec_persistent -s 40 -c 2 -f -o gpurun_out/r02_prof_cell python bench.py $ARGS --lanes 0 > gpurun_out/r02_prof_cell.log 2>&1; echo "cell capture rc=$?"
timeout 300 python tools/run_mixed_projection.py > gpurun_out/r02_mixed_projection.json 2> gpurun_out/r02_mixed_projection.err; echo "mixed rc=$?"; tail -c 1500 gpurun_out/r02_mixed_projection.json; tail -3 gpurun_out/r02_mixed_projection.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 2 -c 2 -f -o gpurun_out/r02_prof_tf32 python tools/run_mixed_projection.py > gpurun_out/r02_prof_tf32.log 2>&1; echo "tf32 capture rc=$?"
timeout 600 python tools/bench_hbm_kernels.py --shape 34,17,17 --rows 60000 > gpurun_out/r02_hbm_kernels.jsonl 2> gpurun_out/r02_hbm_kernels.err; echo "hbm rc=$?"
timeout 600 python tools/bench_hbm_kernels.py --shape 34,17,17 --rows 60000 --p2p 1 > gpurun_out/r02_hbm_kernels_p2p.jsonl 2>> gpurun_out/r02_hbm_kernels.err; echo "hbm p2p rc=$?"
python -c "
import json
for f in ['gpurun_out/r02_hbm_kernels.jsonl','gpurun_out/r02_hbm_kernels_p2p.jsonl']:
    for l in open(f):
        d=json.loads(l); print(f[-12:], d['kernel'][:28], round(d['ms']*1e3,1),'us', round(d['GBps'] or 0), round(d['frac'] or 0,2))
"
ls -la gpurun_out/*.ncu-rep
