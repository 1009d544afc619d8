cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 200 -k "conv5 or resblock or resnet_pass or gemm_core" > gpurun_out/r7_conv.log 2>&1
echo "conv exit $?" >> gpurun_out/r7_conv.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 > gpurun_out/r7_e2e.log 2>&1
echo "e2e exit $?" >> gpurun_out/r7_e2e.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
DMP2_CONV_CLUSTER=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r7_bench_nocluster.json 2> gpurun_out/r7_bench_nocluster.err
