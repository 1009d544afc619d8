cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 120 -k "vgru" > gpurun_out/r4_vgru.log 2>&1
echo "vgru exit $?" >> gpurun_out/r4_vgru.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r4_all.log 2>&1
echo "all exit $?" >> gpurun_out/r4_all.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
