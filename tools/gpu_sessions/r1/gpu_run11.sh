cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 300 -k "hgru or coord_gru" > gpurun_out/r11_gru.log 2>&1
echo "gru exit $?" >> gpurun_out/r11_gru.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r11_all.log 2>&1
echo "all exit $?" >> gpurun_out/r11_all.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r11_launches.csv python tools/profile_fold.py 1 f16f8 > gpurun_out/r11_ncu_launches.log 2>&1
