cd $GRAFT_REPO_ROOT
timeout 300 python tools/time_vgru.py > gpurun_out/r29_time_vgru.log 2>&1
echo "time_vgru rc=$?" >> gpurun_out/r29_time_vgru.log
timeout 600 python -m pytest tests -m gpu -x -q -k "vgru or pf10963 or structured" > gpurun_out/r29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r29_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err
