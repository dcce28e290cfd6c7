#!/bin/bash
# 2-GPU A/B of how the ghost exchange shares the machine with the persistent cell kernel (every run under timeout)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 170 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline --no-scf --no-e2e $EXTRA ) > gpurun_out/ab_$tag.log 2>&1; python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/ab_$tag.log') if l.startswith('{"metric"')][0]
    print('$tag', 'ms/step', round(d['ms_per_step'],1), 'value', '%.4g'%d['value'])
except Exception as e:
    print('$tag', 'ERR', e); print(open('gpurun_out/ab_$tag.log').read()[-800:])
PY
}
EXTRA="--reserved-sms 2" run res2_ch2 NCCL_MAX_NCHANNELS=2 NCCL_MIN_NCHANNELS=2
EXTRA="--reserved-sms 4" run res4_ch4 NCCL_MAX_NCHANNELS=4 NCCL_MIN_NCHANNELS=4
