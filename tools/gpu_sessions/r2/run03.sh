cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_stages.py -x -q -m gpu > gpurun_out/r2/03_stages.log 2>&1
echo "stages exit $?" >> gpurun_out/r2/03_stages.log
timeout 600 python tools/diag_stages64.py 300 1000 0 > gpurun_out/r2/03_stages64.log 2>&1
timeout 600 python tools/diag_fold_dump.py cfg2_s0_n10_m100 300 1000 0 10 100 > gpurun_out/r2/03_fold_cfg2.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --conv-mode f16x3 > gpurun_out/r2/03_bench_f16x3.json 2> gpurun_out/r2/03_bench_f16x3.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --conv-mode f16f8 > gpurun_out/r2/03_bench_f16f8.json 2> gpurun_out/r2/03_bench_f16f8.err
