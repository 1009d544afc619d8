cd $GRAFT_REPO_ROOT
timeout 600 python tools/conv_accuracy.py > gpurun_out/r17_convacc.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r17_all.log 2>&1
echo "all exit $?" >> gpurun_out/r17_all.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
