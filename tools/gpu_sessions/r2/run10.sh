cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/r2/10_stages.log 2>&1
echo "stages exit $?" >> gpurun_out/r2/10_stages.log
timeout 200 python tools/time_conv.py > gpurun_out/r2/10_time_conv.log 2>&1
DMP2_CONV_CLUSTER=2 timeout 200 python tools/time_conv.py >> gpurun_out/r2/10_time_conv.log 2>&1
DMP2_CONV_CHUNK=5 timeout 200 python tools/time_conv.py >> gpurun_out/r2/10_time_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 40 -c 1 -f -o gpurun_out/r2/10_prof_f16f8 python tools/time_conv.py f16f8 > gpurun_out/r2/10_ncu_f16f8.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg4" > gpurun_out/r2/10_cfg4.log 2>&1
