#!/bin/bash
# Multi-GPU legs of the bench on N GPUs of one box (run under `gpurun --gpus N`): weak scaling, configs[2]
# strong scaling (--total-clips 512), configs[3] training step.  Every leg under its own timeout.
N=${1:-2}
run() { tag=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/bench_${tag}_${N}gpu.json 2> gpurun_out/bench_${tag}_${N}gpu.err; echo "$tag rc=$?"; tail -c 300 gpurun_out/bench_${tag}_${N}gpu.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${tag}_${N}gpu.json').read().strip().splitlines()[-1])
    print('$tag', d['metric'], round(d['value'],1), 'ms', round(d['ms_per_step'],3), d.get('scaling'), 'e2e', d.get('e2e',{}).get('value'), d.get('allreduce'))
except Exception as e: print('no line', e)
PY
}
run weak --total-clips 0 --steps 20 --warmup 3
run total512 --steps 10 --warmup 3
run train --train --steps 3 --warmup 1
