mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02b_stem_gemm_pool -f python tools/stem_case.py > gpurun_out/ncu_s1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_s2d_kernel --launch-skip 2 -c 1 -o gpurun_out/r02b_stem_s2d -f python tools/stem_case.py > gpurun_out/ncu_s2.log 2>&1
tail -2 gpurun_out/ncu_s1.log gpurun_out/ncu_s2.log; ls -la gpurun_out/r02b_*.ncu-rep
