cd $GRAFT_REPO_ROOT
TAG=${1:-r3}
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_fold.py 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv5_tc -s 18 -c 2 -o gpurun_out/${TAG}_conv -f python tools/profile_fold.py 1 > gpurun_out/${TAG}_ncu_conv.log 2>&1
