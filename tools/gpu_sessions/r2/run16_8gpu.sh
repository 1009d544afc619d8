cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2
nvidia-smi -L > gpurun_out/r2/16_cfg3_8gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/throughput_cfg3.py --targets 256 --streams 1,4 >> gpurun_out/r2/16_cfg3_8gpu.log 2>&1
echo "rc=$?" >> gpurun_out/r2/16_cfg3_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2/16_bench_8gpu.json 2> gpurun_out/r2/16_bench_8gpu.err
