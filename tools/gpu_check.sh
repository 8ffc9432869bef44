#!/bin/bash
# usage: tools/gpu_check.sh  (on the GPU box, from the repo root)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for cfg in "simt 128 128 0" "tc 128 128 0" "tc 256 256 1" "tc 192 320 1"; do
  set -- $cfg
  echo "=== $cfg ===" | tee -a gpurun_out/check.log
  timeout 600 python tools/gpu_check.py --impl $1 --h $2 --w $3 --graph $4 >> gpurun_out/check.log 2>&1
  echo "exit $?" | tee -a gpurun_out/check.log
done
tail -n 400 gpurun_out/check.log
