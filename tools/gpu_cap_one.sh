mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k k_conv3x3_tc2 -s 9 -c 1 -f -o gpurun_out/r2l_flow_conv0_L0 python tools/profile_step.py 0 > gpurun_out/r2l_ncu_flow_conv0_L0.log 2>&1
ls -la gpurun_out | grep r2l_flow
