mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 6 -o gpurun_out/prof_conv_v1 python tools/conv_cases.py l3_conv3,l3_down,l1_conv3 256 fp16x3 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
