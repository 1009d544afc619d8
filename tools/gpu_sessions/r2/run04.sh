cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
for c in 1 5 25; do DMP2_CONV_CHUNK=$c timeout 200 python tools/time_conv.py >> gpurun_out/r2/04_time_conv.log 2>&1; done
DMP2_CONV_CLUSTER=1 timeout 200 python tools/time_conv.py >> gpurun_out/r2/04_time_conv.log 2>&1
DMP2_FUSE_STATS=0 timeout 200 python tools/time_conv.py f16x3 >> gpurun_out/r2/04_time_conv.log 2>&1
DMP2_CONV_SMS=132 timeout 200 python tools/time_conv.py f16x3 >> gpurun_out/r2/04_time_conv.log 2>&1
