cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final2_all.log 2>&1
echo "all exit $?" >> gpurun_out/final2_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final2_smoke.log
timeout 600 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final2_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final2_ncu_bench.log 2>&1
DMP2_EIG_CL=16 timeout 100 python tools/time_eig.py 300 > gpurun_out/final2_eig_cl16.log 2>&1
