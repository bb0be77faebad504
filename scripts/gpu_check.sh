#!/bin/bash
# Bounded GPU check used during kernel work: every leg under its own timeout, logs into gpurun_out/.
# usage: bash scripts/gpu_check.sh <tag> [pytest -k expression]
tag=$1; kexpr=${2:-"encoder or tensor_core_block"}
timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q -k "$kexpr" > gpurun_out/test_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_$tag.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-inference > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
    print('clips/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
    for k,v in d['kernels'].items(): print(' ', k, v.get('ms_per_step'), v.get('launches_per_step'), v.get('frac'))
except Exception as e: print('no bench line', e)
PY
