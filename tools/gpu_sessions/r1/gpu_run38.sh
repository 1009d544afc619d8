cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r38_strip4.log 2>&1
for W in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 tools/time_strip.py 2048 3000 10 100 >> gpurun_out/r38_strip4.log 2>&1
echo "W=$W rc=$?" >> gpurun_out/r38_strip4.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 tools/time_strip.py 1024 2048 30 200 >> gpurun_out/r38_strip4.log 2>&1
echo "cfg4-size W=4 rc=$?" >> gpurun_out/r38_strip4.log
