cd $GRAFT_REPO_ROOT
O=gpurun_out/r2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k "conv" > $O/21_conv_tests.log 2>&1
echo "exit $?" >> $O/21_conv_tests.log
# single-stream cost of the dynamic schedule
timeout 200 python tools/time_conv.py f16f8 f16x3 > $O/21_time_conv.log 2>&1
DMP2_CONV_DYNAMIC=1 timeout 200 python tools/time_conv.py f16f8 f16x3 >> $O/21_time_conv.log 2>&1
# throughput: cfg2 (12 targets) and cfg3 (32 targets = one GPU's share of the 256), static vs dynamic
for dyn in 0 1; do
  timeout 300 python tools/throughput_cfg3.py --L 300 --N 1000 --targets 12 --streams 1,2,3,4 --dynamic $dyn >> $O/21_tp_cfg2.log 2>&1
  timeout 300 python tools/throughput_cfg3.py --targets 32 --streams 1,2,3,4,6 --dynamic $dyn >> $O/21_tp_cfg3.log 2>&1
done
timeout 300 python tools/throughput_cfg3.py --L 300 --N 1000 --targets 12 --streams 3,4 --dynamic 1 --conv-sms 128 >> $O/21_tp_cfg2.log 2>&1
timeout 300 python tools/throughput_cfg3.py --targets 32 --streams 3,4,6 --dynamic 1 --conv-sms 128 >> $O/21_tp_cfg3.log 2>&1
DMP2_EIG_CL=8 timeout 300 python tools/throughput_cfg3.py --L 300 --N 1000 --targets 12 --streams 3 --dynamic 1 >> $O/21_tp_cfg2.log 2>&1
