cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final_all.log 2>&1
echo "all exit $?" >> gpurun_out/final_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
