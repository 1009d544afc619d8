cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/25_strip_4gpu.log 2>&1
timeout 600 python -m pytest tests/test_gpu_strip.py -q -m gpu > gpurun_out/r2/25_strip_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/25_strip_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/time_strip.py 2048 3000 10 100 >> gpurun_out/r2/25_strip_4gpu.log 2>&1
echo "W=4 rc=$?" >> gpurun_out/r2/25_strip_4gpu.log
