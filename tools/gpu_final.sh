#!/bin/bash
# Final pass of a round: GPU tests, smoke, bench (both arms), ncu launch list + full-set captures of the conv kernels.
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 --op-table gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_conv3x3_tc2$ -c 12 -f -o gpurun_out/prof_${TAG}_pair python tools/profile_step.py 0 > gpurun_out/ncu_pair_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_conv3x3_tc$ -c 10 -f -o gpurun_out/prof_${TAG}_single python tools/profile_step.py 0 > gpurun_out/ncu_single_$TAG.log 2>&1
ls -la gpurun_out/ | grep -E "$TAG"
