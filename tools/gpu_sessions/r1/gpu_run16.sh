cd $GRAFT_REPO_ROOT
timeout 1500 python tools/diag_cfg2.py 300 1000 > gpurun_out/r16_diag.log 2>&1
