cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_strip.py tests/test_gpu_stages.py -x -q -k "strip or dca or reweight" > gpurun_out/r37_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r37_tests.log
