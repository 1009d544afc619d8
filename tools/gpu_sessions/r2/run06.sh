cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 40 -c 2 -f -o gpurun_out/r2/06_prof_f16f8 python tools/time_conv.py f16f8 > gpurun_out/r2/06_ncu_f16f8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 40 -c 2 -f -o gpurun_out/r2/06_prof_f16x3 python tools/time_conv.py f16x3 > gpurun_out/r2/06_ncu_f16x3.log 2>&1
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k "vgru" > gpurun_out/r2/06_vgru.log 2>&1
timeout 300 python tools/diag_stages64.py 300 1000 0 > gpurun_out/r2/06_stages64.log 2>&1
