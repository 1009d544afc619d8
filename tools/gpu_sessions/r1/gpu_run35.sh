cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_strip.py tests/test_gpu_stages.py -x -q -k "strip or eig" > gpurun_out/r35_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r35_tests.log
timeout 300 python tools/time_eig.py 700 1024 1500 2048 > gpurun_out/r35_eig.log 2>&1
