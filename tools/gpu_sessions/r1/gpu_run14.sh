cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 300 -k "conv_cluster_variants" > gpurun_out/r14_pair.log 2>&1
echo "pair exit $?" >> gpurun_out/r14_pair.log
DMP2_CONV_CLUSTER=pair timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 600 -k "pf10963 or structured" > gpurun_out/r14_e2e_pair.log 2>&1
echo "e2e exit $?" >> gpurun_out/r14_e2e_pair.log
DMP2_CONV_CLUSTER=pair timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r14_bench_pair.json 2> gpurun_out/r14_bench_pair.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r14_bench_default.json 2> gpurun_out/r14_bench_default.err
DMP2_CONV_CLUSTER=pair timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --conv-mode f16x3 > gpurun_out/r14_bench_pair_x3.json 2> gpurun_out/r14_bench_pair_x3.err
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 800 -k "cfg5" -s > gpurun_out/r14_cfg5.log 2>&1
echo "cfg5 exit $?" >> gpurun_out/r14_cfg5.log
