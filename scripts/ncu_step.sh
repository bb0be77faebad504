#!/bin/bash
# One eager step under ncu (per-launch gpu time), summarised by scripts/step_share.py.
# usage (on the GPU box): bash scripts/ncu_step.sh <tag> [extra bench.py flags]
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --total-clips 0 --steps 1 --warmup 1 --no-cpu-baseline --no-cuda-graph "$@" > gpurun_out/ncu_step_$tag.log 2>&1
python scripts/step_share.py gpurun_out/launches_$tag.csv > gpurun_out/step_share_$tag.txt
cat gpurun_out/step_share_$tag.txt
