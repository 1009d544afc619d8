cd $GRAFT_REPO_ROOT
DMP2_STRIP_TRACE=1 timeout 300 python -m pytest tests/test_gpu_strip.py -x -q -s -k "single_engine and 2" > gpurun_out/r31_strip.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r31_strip.log
