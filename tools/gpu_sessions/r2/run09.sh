cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg2 or cfg3 or pf10963_long or cfg4" > gpurun_out/r2/09_parity_f16x3.log 2>&1
DMP2_CONV_MODE=f16f8 timeout 1500 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg2 or cfg3 or pf10963_long or cfg4" > gpurun_out/r2/09_parity_f16f8.log 2>&1
DMP2_CONV_MODE=ffma timeout 1500 python -m pytest tests/test_gpu_parity_r2.py -q -s -m gpu -k "cfg2 or cfg3 or pf10963_long" > gpurun_out/r2/09_parity_ffma.log 2>&1
