#!/bin/bash
# Source-level ncu captures of the CTA-pair conv kernel's small-N launches:
# which barrier does the MMA warp sleep on in k_conv3x3_tc2<64,64> (fusion_conv1@L0) / <32,64> (flow_conv0@L0)?
#   gpurun --timeout 900 -- 'bash tools/ncu_pair_stalls.sh r2a'
# then here:  ncu -i gpurun_out/prof_<tag>_<name>.ncu-rep --page source --csv > ...   (tools/summarize_ncu.py)
mkdir -p gpurun_out
TAG=${1:-r2a}
# launch list first: the -s indices below are positions among kernels matching the regex in ONE eager call
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 1 > gpurun_out/ncu_list_$TAG.log 2>&1
for spec in "tc2_64_64 k_conv3x3_tc2<64 0" "tc2_32_64 k_conv3x3_tc2<32 0"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -f \
    -o gpurun_out/prof_${TAG}_$1 python tools/profile_step.py 1 > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
done
ls -la gpurun_out/
