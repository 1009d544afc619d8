cd $GRAFT_REPO_ROOT
O=gpurun_out/r2
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x -k "shared_vgru or batch" > $O/23_scan_tests.log 2>&1
echo "exit $?" >> $O/23_scan_tests.log
for sr in 0 384 768 1536; do
  timeout 300 python tools/throughput_cfg3.py --targets 32 --streams 2,3,4 --scan-rows $sr >> $O/23_tp_cfg3.log 2>&1
done
for sr in 0 600 1200; do
  timeout 300 python tools/throughput_cfg3.py --L 300 --N 1000 --targets 12 --streams 3 --scan-rows $sr >> $O/23_tp_cfg2.log 2>&1
done
