cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -q -m gpu -x > gpurun_out/r2/17_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/17_tests.log
timeout 300 python tools/time_eig.py 82 300 640 > gpurun_out/r2/17_eig.log 2>&1
DMP2_EIG_INVIT=2 timeout 300 python tools/time_eig.py 82 300 640 >> gpurun_out/r2/17_eig.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2/17_bench.json 2> gpurun_out/r2/17_bench.err
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg2 or cfg3" > gpurun_out/r2/17_parity.log 2>&1
timeout 600 python tools/torch_cuda_bar.py > gpurun_out/r2/17_cuda_bar.log 2>&1
