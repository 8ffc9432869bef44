#!/bin/bash
# Round-end style GPU pass: tests, bench, ncu launch list, ncu full captures of selected conv launches.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --op-table gpurun_out/ops_$TAG.csv > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1
for spec in "fe_conv1_L0 82" "flow_conv0_L3 128" "flow_conv0_L0 137" "fusion_conv1_L0 162"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s $2 -c 1 -f -o gpurun_out/prof_${TAG}_$1 python tools/profile_step.py 1 > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
done
ls -la gpurun_out/
fi
