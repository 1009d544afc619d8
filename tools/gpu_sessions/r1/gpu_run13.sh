cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r13_bench2.json 2> gpurun_out/r13_bench2.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r13_bench1.json 2> gpurun_out/r13_bench1.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r13_ref.json 2> gpurun_out/r13_ref.err
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 300 -k "eig or head_mds or resblock" > gpurun_out/r13_eig.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1
