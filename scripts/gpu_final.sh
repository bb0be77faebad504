#!/bin/bash
# Round-end evidence on one B200: full GPU suite, smoke, the bench line, ncu launch list of one eager step and
# one `ncu --set full` capture of the block kernels.  Every leg under its own timeout; logs in gpurun_out/.
tag=${1:-r02}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/${tag}_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*pack)(?!.*fold_bias)(?!.*scale_cols).*$' -c 2400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --total-clips 0 --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph --no-full-inference --roofline-seconds 0 > gpurun_out/${tag}_ncu_step.log 2>&1; echo "ncu list rc=$?"
python scripts/step_share.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_step_share.txt 2>&1; head -30 gpurun_out/${tag}_step_share.txt
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:tc_back|tc_front1|tc_apply|gram_project' --launch-skip 8 -c 4 -f -o gpurun_out/${tag}_prof_blocks \
  python scripts/prof_block.py 32 6912 1 4 > gpurun_out/${tag}_prof_blocks.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flow_encoder -s 3 -c 1 -f -o gpurun_out/${tag}_prof_flow \
  python scripts/prof_flow.py > gpurun_out/${tag}_prof_flow.log 2>&1; echo "ncu flow rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rsb_level -s 270 -c 18 -f -o gpurun_out/${tag}_prof_rsb \
  python scripts/prof_rsb.py > gpurun_out/${tag}_prof_rsb.log 2>&1; echo "ncu rsb rc=$?"
timeout 100 python scripts/prof_flow.py > gpurun_out/${tag}_time_flow.log 2>&1; timeout 100 python scripts/prof_rsb.py > gpurun_out/${tag}_time_rsb.log 2>&1; cat gpurun_out/${tag}_time_flow.log gpurun_out/${tag}_time_rsb.log
