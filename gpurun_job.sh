mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tunables.py -q -m gpu 2>&1 | tail -4
USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r01_conv_tc_l3_down python tools/conv_cases.py l3_down 256 fp16x3 > /dev/null 2>&1
USOT_DEBUG_SPLIT_OUT=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r01_conv_tc_l3_conv3 python tools/conv_cases.py l3_conv3 256 fp16x3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:groupdw_tma -s 1 -c 1 -o gpurun_out/r01_groupdw_tma python tools/groupdw_case.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 2 -c 1 -o gpurun_out/r01_stem_tc python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
