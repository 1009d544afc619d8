cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r5_all.log 2>&1
echo "all exit $?" >> gpurun_out/r5_all.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_launches.csv python tools/profile_fold.py 1 > gpurun_out/r5_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vgru_step -s 500 -c 1 -o gpurun_out/r5_vgru -f python tools/profile_fold.py 0 > gpurun_out/r5_ncu_vgru.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_eig_top8 -c 1 -o gpurun_out/r5_eig -f python tools/profile_fold.py 0 > gpurun_out/r5_ncu_eig.log 2>&1
