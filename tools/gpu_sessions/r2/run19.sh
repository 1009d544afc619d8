cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2/19_bench_lowprio.json 2> gpurun_out/r2/19_bench.err
DMP2_SIDE_PRIORITY=default timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2/19_bench_defprio.json 2>> gpurun_out/r2/19_bench.err
timeout 600 python -m pytest tests/test_gpu_e2e.py -q -m gpu -x > gpurun_out/r2/19_e2e.log 2>&1
