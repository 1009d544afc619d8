cd $GRAFT_REPO_ROOT
O=gpurun_out/r2
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k "eig or head_mds" > $O/28_eig_tests.log 2>&1
echo "exit $?" >> $O/28_eig_tests.log
timeout 300 python tools/time_eig.py 82 150 300 640 1024 2048 > $O/28_eig.log 2>&1
