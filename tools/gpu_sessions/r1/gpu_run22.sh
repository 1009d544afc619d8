cd $GRAFT_REPO_ROOT
timeout 900 python tools/torch_cuda_bar.py > gpurun_out/r22_torchbar.log 2>&1
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 600 -k "batch_api" > gpurun_out/r22_batch.log 2>&1
