#!/bin/bash
# Targeted full-set ncu captures (source-level stall reasons).  -k matches the function base name; the launch is picked by
# its index among the launches of that kernel in one eager call (schedule order, see profiles/r2g_ops_eventtimed_1080p.csv).
mkdir -p gpurun_out
T=${1:-r2h}
cap() { name=$1; kern=$2; skip=$3; cnt=$4; timeout 600 ncu --set full --clock-control none --import-source on -k $kern -s $skip -c $cnt -f -o gpurun_out/${T}_$name python tools/profile_step.py 0 > gpurun_out/${T}_ncu_$name.log 2>&1; }
cap flow_conv12_L0 k_conv3x3_tc 46 2
cap fe_conv1_L0 k_conv3x3_tc 0 1
cap flow_conv0_L0 k_conv3x3_tc2 9 1
cap fusion_conv1_L0 k_conv3x3_tc2 14 1
ls -la gpurun_out | grep ${T}
