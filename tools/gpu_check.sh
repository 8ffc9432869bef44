#!/bin/bash
# usage: tools/gpu_check.sh  (on the GPU box, from the repo root): staged parity check, each in its own process
mkdir -p gpurun_out; rm -f gpurun_out/check.log
for cfg in "tc 128 128 0" "tc 256 256 1" "tc 192 320 1" "tc 200 136 0"; do
  set -- $cfg
  echo "=== $cfg ===" | tee -a gpurun_out/check.log
  timeout 300 python tools/gpu_check.py --impl $1 --h $2 --w $3 --graph $4 >> gpurun_out/check.log 2>&1
  echo "exit $?" | tee -a gpurun_out/check.log
done
grep -E "===|RESULT|exit|repeat|rror|timeout" gpurun_out/check.log | cut -c1-300
