cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -x -q -k "eig or head_mds or pf10963 or determin" > gpurun_out/r40_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r40_tests.log
timeout 200 python tools/time_eig.py 82 150 300 436 520 640 > gpurun_out/r40_eig.log 2>&1
