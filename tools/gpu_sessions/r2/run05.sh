cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/r2/05_stages.log 2>&1
echo "stages exit $?" >> gpurun_out/r2/05_stages.log
timeout 200 python tools/time_conv.py > gpurun_out/r2/05_time_conv.log 2>&1
DMP2_CONV_CLUSTER=2 timeout 200 python tools/time_conv.py >> gpurun_out/r2/05_time_conv.log 2>&1
DMP2_CONV_CHUNK=5 timeout 200 python tools/time_conv.py >> gpurun_out/r2/05_time_conv.log 2>&1
DMP2_CONV_SMS=132 timeout 200 python tools/time_conv.py f16x3 >> gpurun_out/r2/05_time_conv.log 2>&1
timeout 600 python tools/diag_fold_dump.py cfg2_s0_n10_m100 300 1000 0 10 100 structured f16x3,f16f8 > gpurun_out/r2/05_fold_cfg2.log 2>&1
