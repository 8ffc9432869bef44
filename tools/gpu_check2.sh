#!/bin/bash
# CTA-pair kernel validation: every eligible layer on the pair kernel (FILM_2CTA=2), staged parity vs the oracle
mkdir -p gpurun_out; rm -f gpurun_out/check2.log
for cfg in "128 128" "256 256" "192 320"; do
  set -- $cfg
  echo "=== 2cta $cfg ===" | tee -a gpurun_out/check2.log
  FILM_2CTA=2 timeout 200 python tools/gpu_check.py --impl tc --h $1 --w $2 --graph 0 >> gpurun_out/check2.log 2>&1
  echo "exit $?" | tee -a gpurun_out/check2.log
done
grep -E "===|RESULT|exit|repeat|rror|timeout|feat0/0|feat0/3|res_fwd/0|res_fwd/6" gpurun_out/check2.log | cut -c1-200
