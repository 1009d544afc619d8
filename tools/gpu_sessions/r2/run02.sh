cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 120 ./build/tmem_bw > gpurun_out/r2/02_tmem_bw.log 2>&1
timeout 600 python tools/probe_tmem_acc.py > gpurun_out/r2/02_tmem_acc.log 2>&1
timeout 900 python tools/probe_conv_twolevel.py > gpurun_out/r2/02_conv_twolevel.log 2>&1
