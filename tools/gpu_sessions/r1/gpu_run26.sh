cd $GRAFT_REPO_ROOT
timeout 1200 python tools/diag_cfg2.py 164 512 > gpurun_out/r26_diag164.log 2>&1
