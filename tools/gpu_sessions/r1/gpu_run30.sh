cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_strip.py -x -q -s > gpurun_out/r30_strip.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r30_strip.log
