cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/15_strip_4gpu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_strip.py -q -m gpu > gpurun_out/r2/15_strip_tests.log 2>&1
echo "exit $?" >> gpurun_out/r2/15_strip_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/time_strip.py 2048 3000 10 100 >> gpurun_out/r2/15_strip_4gpu.log 2>&1
echo "W=4 rc=$?" >> gpurun_out/r2/15_strip_4gpu.log
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/time_strip.py 2048 3000 10 100 >> gpurun_out/r2/15_strip_4gpu.log 2>&1
echo "W=2 rc=$?" >> gpurun_out/r2/15_strip_4gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 tools/time_strip.py 1024 2048 30 200 >> gpurun_out/r2/15_strip_4gpu.log 2>&1
echo "cfg4-size W=4 rc=$?" >> gpurun_out/r2/15_strip_4gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r2/15_bench_4gpu.json 2> gpurun_out/r2/15_bench_4gpu.err
