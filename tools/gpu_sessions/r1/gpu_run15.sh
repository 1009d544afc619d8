cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r15_all.log 2>&1
echo "all exit $?" >> gpurun_out/r15_all.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r15_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r15_ncu_bench.log 2>&1
