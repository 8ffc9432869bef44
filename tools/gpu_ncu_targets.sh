#!/bin/bash
# Targeted full-set ncu captures (source-level stall reasons) of the level-0 small-N launches and the first layer.
mkdir -p gpurun_out
T=${1:-r2h}
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_conv3x3_tc<32, 32>" -c 2 -f -o gpurun_out/${T}_tc3232 python tools/profile_step.py 0 > gpurun_out/${T}_ncu_tc3232.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_conv3x3_tc2<32, 64>" -c 1 -f -o gpurun_out/${T}_tc2_3264 python tools/profile_step.py 0 > gpurun_out/${T}_ncu_tc2_3264.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_fe_conv0" -c 1 -f -o gpurun_out/${T}_feconv0 python tools/profile_step.py 0 > gpurun_out/${T}_ncu_feconv0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_conv3x3_tc<64, 64>" -c 1 -f -o gpurun_out/${T}_tc6464 python tools/profile_step.py 0 > gpurun_out/${T}_ncu_tc6464.log 2>&1
ls -la gpurun_out | grep ${T}
