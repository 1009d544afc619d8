cd $GRAFT_REPO_ROOT
O=gpurun_out/final_r2
mkdir -p $O gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k "eig or head_mds" > gpurun_out/r2/20_eig_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/20_eig_tests.log
timeout 300 python tools/time_eig.py 82 300 640 1024 > gpurun_out/r2/20_eig.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)2" -s 40 -c 2 -f -o $O/conv_f16f8 python tools/time_conv.py f16f8 > $O/ncu_conv_f16f8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)1" -s 60 -c 2 -f -o $O/conv_f16x3 python tools/time_conv.py f16x3 > $O/ncu_conv_f16x3.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"k_conv5_tc<\(int\)1" -c 10 -f -o $O/gemm_tc python tools/time_conv.py f16f8 > $O/ncu_gemm.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2/20_bench.json 2> gpurun_out/r2/20_bench.err
