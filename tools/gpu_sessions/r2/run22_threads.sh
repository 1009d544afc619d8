cd $GRAFT_REPO_ROOT
O=gpurun_out/r2
mkdir -p $O
for ht in 0 1; do
  timeout 300 python tools/throughput_cfg3.py --L 300 --N 1000 --targets 12 --streams 2,3,4 --host-threads $ht >> $O/22_tp_cfg2.log 2>&1
  timeout 300 python tools/throughput_cfg3.py --targets 32 --streams 2,3,4,6 --host-threads $ht >> $O/22_tp_cfg3.log 2>&1
done
timeout 300 python tools/throughput_cfg3.py --targets 32 --streams 3,4 --host-threads 1 --dynamic 0 >> $O/22_tp_cfg3.log 2>&1
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x -k "batch or stream or pool" > $O/22_e2e_batch.log 2>&1
echo "exit $?" >> $O/22_e2e_batch.log
