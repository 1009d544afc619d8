cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_stages.py -q -m gpu -x > gpurun_out/r2/14_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/14_tests.log
timeout 300 python tools/diag_stages64.py 300 1000 0 > gpurun_out/r2/14_stages64.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/14_bench.json 2> gpurun_out/r2/14_bench.err
timeout 300 python tools/time_eig.py 300 > gpurun_out/r2/14_eig.log 2>&1
