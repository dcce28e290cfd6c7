#!/bin/bash
# round-2 GPU check X (1 GPU): final code: full -m gpu suite, default bench (both arms), ncu launch list of a short bench
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -6 gpurun_out/r2x_pytest.log
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/r2x_bench.json; tail -3 gpurun_out/r2x_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; echo "ref rc=$?"
ARGS="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-scf --nwfc 512"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 800 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py $ARGS > gpurun_out/r02_launches_final_bench.log 2>&1; echo "launch list rc=$?"
python tools/ncu_summary.py --launches gpurun_out/r02_launches_final.csv gpurun_out/r02_launch_shares_final.csv; head -12 gpurun_out/r02_launch_shares_final.csv
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
