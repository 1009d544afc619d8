cd $GRAFT_REPO_ROOT
timeout 900 python tools/time_configs.py > gpurun_out/r23_configs.log 2>&1
