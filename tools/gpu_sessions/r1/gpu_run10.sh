cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/r10_all.log 2>&1
echo "all exit $?" >> gpurun_out/r10_all.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 18 -c 1 -o gpurun_out/r10_conv -f python tools/profile_fold.py 1 f16f8 > gpurun_out/r10_ncu_conv.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r10_launches.csv python tools/profile_fold.py 1 f16f8 > gpurun_out/r10_ncu_launches.log 2>&1
