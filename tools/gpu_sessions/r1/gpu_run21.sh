cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r21_all.log 2>&1
echo "all exit $?" >> gpurun_out/r21_all.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err
DMP2_EIG_CL=16 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r21_bench_eig16.json 2> gpurun_out/r21_bench_eig16.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r21_launches.csv python tools/profile_fold.py 1 f16f8 > gpurun_out/r21_ncu_launches.log 2>&1
