cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2/12_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/12_tests.log
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg2 or cfg3" > gpurun_out/r2/12_parity.log 2>&1
timeout 300 python tools/diag_stages64.py 300 1000 0 > gpurun_out/r2/12_stages64.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/12_bench.json 2> gpurun_out/r2/12_bench.err
