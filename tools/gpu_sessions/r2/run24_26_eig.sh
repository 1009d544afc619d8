cd $GRAFT_REPO_ROOT
O=gpurun_out/r2
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stages.py -q -m gpu -x -k "eig or head_mds" > $O/26_eig_tests.log 2>&1
echo "exit $?" >> $O/26_eig_tests.log
timeout 300 python tools/time_eig.py 82 150 300 640 1024 2048 > $O/26_eig.log 2>&1
DMP2_EIG_CL=8 timeout 300 python tools/time_eig.py 82 150 300 >> $O/26_eig.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_e2e.py -q -m gpu -x > $O/26_parity_e2e.log 2>&1
echo "exit $?" >> $O/26_parity_e2e.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/26_bench.json 2> $O/26_bench.err
